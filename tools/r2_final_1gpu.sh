#!/bin/bash
# Round 2 final evidence on one GPU:  gpurun --timeout 2400 -- bash tools/r2_final_1gpu.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/f_tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/f_tests_gpu.log
timeout 600 python bench.py > $O/f_bench_full.log 2>&1; echo "full bench rc=$?"
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > $O/f_bench_reference.log 2>&1; echo "reference arm rc=$?"
tail -c 600 $O/f_bench_reference.log
for c in 0 3 4; do
  timeout 600 python bench.py --config $c > $O/f_config$c.log 2>&1; echo "config $c rc=$?"
done
timeout 400 python bench.py --config 2 --no-e2e > $O/f_config2.log 2>&1; echo "config 2 rc=$?"
# ncu: launch list of whole steps, then one full capture of the tracking kernel
timeout 300 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/f_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $O/f_launches.csv python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 \
    > $O/f_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:oa_track_kernel -s 4 -c 1 -o $O/f_track_full -f \
    python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/f_ncu_full.log 2>&1
echo "full capture rc=$?"
python - $O/f_bench_full.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d['roofline']
print('value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'], 'kernel %.3f'%r['kernel_ms'], 'frac %.3f'%r['frac'],
      'parity', d.get('parity'), 'e2e %.3g'%d['e2e']['value'], 'entry %.3g'%d['e2e_entry_point']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
PY
