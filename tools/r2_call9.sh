#!/bin/bash
# dynamic chunk tickets in the hash kernel: GPU suite, 1-GPU bench, N-GPU bench (K=1)
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c9_tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/c9_tests_gpu.log
timeout 300 python bench.py --no-e2e --no-cpu > $O/c9_bench1.log 2>&1; echo "bench1 rc=$?"
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c9_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
}
run k1 "--steps 32 --no-e2e --no-cpu" OA_DUMMY=1
run k1_r0 "--steps 32 --no-e2e --no-cpu" OA_SM_RESERVE=0
run k1_full "--steps 32" OA_DUMMY=1
python - $O/c9_bench1.log $O/c9_${N}_k1.log $O/c9_${N}_k1_r0.log $O/c9_${N}_k1_full.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
              'events/step %.1f'%d['events_per_step'], 'parity', d.get('parity'), d.get('parity_multi_gpu'),
              'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('h2d_gb_per_s_per_gpu'),
              'host phases', d.get('host_phases_ms_per_step'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
