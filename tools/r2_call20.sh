#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c20_tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/c20_tests_gpu.log
timeout 300 python bench.py --no-e2e > $O/c20_bench1.log 2>&1; echo "bench1 rc=$?"
python - $O/c20_bench1.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d['roofline']
print('value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'], 'kernel %.3f'%r['kernel_ms'], 'parity', d.get('parity'), d['cpu_baseline'].get('angle_ulp_histogram'), d['host_phases_ms_per_step'])
PY
