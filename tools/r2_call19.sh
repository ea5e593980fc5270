#!/bin/bash
# N GPUs: where does an exchange spend its time? (CUDA events between its phases)
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c19_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
}
run prof "--steps 30 --no-e2e --no-cpu" OA_EXCHANGE_PROFILE=1
run nohost "--steps 30 --no-e2e --no-cpu" OA_BENCH_TO_HOST=0 OA_EXCHANGE_PROFILE=1
run reserve40 "--steps 30 --no-e2e --no-cpu" OA_SM_RESERVE=40 OA_EXCHANGE_PROFILE=1
run mainstream "--steps 30 --no-e2e --no-cpu" OA_EXCHANGE_STREAM=main OA_SM_RESERVE=0 OA_EXCHANGE_PROFILE=1
python - $O/c19_${N}_prof.log $O/c19_${N}_nohost.log $O/c19_${N}_reserve40.log $O/c19_${N}_mainstream.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
              'host phases', d.get('host_phases_ms_per_step'), 'exchange phases', d.get('exchange_phases_ms'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
