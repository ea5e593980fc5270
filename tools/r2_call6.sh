#!/bin/bash
# Round 2, multi-GPU: batched exchange after the reserve / fresh-splitter fixes.
#   gpurun --gpus N --timeout 900 -- bash tools/r2_call6.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c6_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
  python - "$O/c6_${N}_$tag.log" "$tag" <<'PY'
import json,sys
f,tag=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
    print(tag, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
          'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
          'events/step %.1f'%d['events_per_step'], 'parity', d.get('parity'), d.get('parity_multi_gpu'),
          'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('events_equal_device_run'),
          'host phases', d.get('host_phases_ms_per_step'))
except Exception as e:
    print(tag,'FAILED',e); print(open(f).read()[-1500:])
PY
}
run k1 "--steps 32 --no-e2e" OA_EXCHANGE_BATCH=1
run k8 "--steps 32 --no-e2e --no-cpu" OA_EXCHANGE_BATCH=8
run k16 "--steps 32 --no-e2e --no-cpu" OA_EXCHANGE_BATCH=16
run k16_e2e "--steps 32 --no-cpu" OA_EXCHANGE_BATCH=16
