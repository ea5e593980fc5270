#!/bin/bash
# 1 GPU: whole suite, the full headline line (e2e, entry-point e2e, CPU arm), the other configs
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/c12_tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/c12_tests_gpu.log
timeout 600 python bench.py > $O/c12_bench_full.log 2>&1; echo "full bench rc=$?"
for c in 0 3 4; do
  timeout 600 python bench.py --config $c > $O/c12_config$c.log 2>&1; echo "config $c rc=$?"
  tail -c 1800 $O/c12_config$c.log
done
timeout 400 python bench.py --config 2 --no-e2e > $O/c12_config2.log 2>&1; echo "config 2 rc=$?"
python - $O/c12_bench_full.log $O/c12_config2.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f'%r['kernel_ms'], 'frac %.3f'%r['frac'], 'parity', d.get('parity'),
              'e2e', d.get('e2e'), 'entry', d.get('e2e_entry_point'),
              'cpu', d.get('cpu_baseline'), 'host phases', d.get('host_phases_ms_per_step'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
