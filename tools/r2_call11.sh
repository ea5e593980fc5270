#!/bin/bash
# 8-GPU (or N-GPU) scaling point with the current defaults; every run under its own timeout
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
echo "host cores: $(nproc)   memory: $(free -g | awk 'NR==2{print $2" GB"}')"
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c11_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
}
run plain "--steps 30 --no-e2e --no-cpu" OA_DUMMY=1
run full "--steps 30" OA_DUMMY=1
run plain2 "--steps 30 --no-e2e --no-cpu --profile" OA_DUMMY=1
python - $O/c11_${N}_plain.log $O/c11_${N}_full.log $O/c11_${N}_plain2.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
              'events/step %.1f'%d['events_per_step'], 'parity', d.get('parity'), d.get('parity_multi_gpu'),
              'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('h2d_gb_per_s_per_gpu'),
              (d.get('e2e') or {}).get('events_equal_device_run'),
              'host phases', d.get('host_phases_ms_per_step'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
grep -A22 "Ordered by: internal time" $O/c11_${N}_plain2.log | head -30
