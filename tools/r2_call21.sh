#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c21_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
}
run k1 "--steps 30 --no-e2e --no-cpu" OA_EXCHANGE_BATCH=1
run k16 "--steps 48 --no-e2e --no-cpu" OA_EXCHANGE_BATCH=16
python - $O/c21_${N}_k1.log $O/c21_${N}_k16.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
              'events/step %.1f'%d['events_per_step'], 'host phases', d.get('host_phases_ms_per_step'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
