#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c23_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
}
run prepack "--steps 40 --no-e2e" OA_EXCHANGE_PREPACK=1 OA_EXCHANGE_PROFILE=1
run plain "--steps 40 --no-e2e --no-cpu" OA_EXCHANGE_PREPACK=0 OA_EXCHANGE_PROFILE=1
python - $O/c23_${N}_prepack.log $O/c23_${N}_plain.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
              'events/step %.1f'%d['events_per_step'], 'parity', d.get('parity'), d.get('parity_multi_gpu'),
              'host phases', d.get('host_phases_ms_per_step'), 'exchange', d.get('exchange_phases_ms'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
