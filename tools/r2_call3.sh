#!/bin/bash
# Round 2: second-generation partitioned join (oa_pj2) on hardware for the first time.
#   gpurun --timeout 900 -- bash tools/r2_call3.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zzz_pjoin.py -x -q \
    -k "pj2 and plain and pericentric" > $O/c3_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -n 12 $O/c3_sanitizer.log
timeout 600 python -m pytest tests/test_gpu_zzz_pjoin.py -q -k pj2 > $O/c3_tests_pj2.log 2>&1; echo "pj2 tests rc=$?"
tail -n 15 $O/c3_tests_pj2.log
summ() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    r = d['roofline']
    print(f, 'value %.2f G' % (d['value'] / 1e9), 'ms/step %.3f' % d['ms_per_step'],
          'kernel %.3f ms' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'events', d['events_per_step'],
          'host', d.get('host_phases_ms_per_step'))
    if 'pj2_stage_profile' in d: print('   ', d['pj2_stage_profile'])
except Exception as e:
    print(f, 'FAILED', e); print(open(f).read()[-1200:])
PY
}
for impl in pj2 hash; do
  OA_TRACK_IMPL=$impl timeout 300 python bench.py --no-e2e --no-cpu > $O/c3_bench_$impl.log 2>&1
  echo "bench $impl rc=$?"; summ $O/c3_bench_$impl.log
done
OA_LIB_PATH=$PWD/variants/liborbit_b200_pj2stats.so OA_TRACK_IMPL=pj2 timeout 300 python bench.py --no-e2e --no-cpu \
    > $O/c3_bench_pj2stats.log 2>&1; echo "stats rc=$?"; summ $O/c3_bench_pj2stats.log
for grp in 18 20 21; do
  OA_PJ2_GROUP=$((1<<grp)) OA_TRACK_IMPL=pj2 timeout 300 python bench.py --no-e2e --no-cpu > $O/c3_bench_grp$grp.log 2>&1
  echo "group 2^$grp rc=$?"; summ $O/c3_bench_grp$grp.log
done
OA_TRACK_IMPL=pj2 timeout 300 python bench.py --no-e2e > $O/c3_bench_pj2_cpu.log 2>&1; echo "pj2 + cpu parity rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/c3_bench_pj2_cpu.log') if l.startswith('{')][-1])
print('parity', d.get('parity'), d.get('cpu_baseline', {}).get('parity_vs_gpu_on_sample'))
PY
OA_TRACK_IMPL=pj2 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:oa_pj2_kernel -s 4 -c 1 -o $O/c3_pj2_full -f \
    python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/c3_ncu_full_pj2.log 2>&1
echo "full capture rc=$?"
