#!/bin/bash
# Round 2, first GPU call (1 GPU):  gpurun --timeout 1700 -- bash tools/r2_call1.sh
# whole GPU suite (no -x), the partitioned join on hardware for the first time,
# bench of both implementations + tuning builds, launch list + one full capture.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/c1_tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -n 4 $O/c1_tests_gpu.log
OA_TEST_PJOIN=1 timeout 600 python -m pytest tests/test_gpu_zzz_pjoin.py -q \
    > $O/c1_tests_pjoin.log 2>&1; echo "pjoin tests rc=$?"
tail -n 4 $O/c1_tests_pjoin.log
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
    if 'pjoin_stage_profile' in d: print('   ', d['pjoin_stage_profile'])
except Exception as e:
    print(f, 'FAILED', e)
PY
}
for impl in hash pjoin; do
  OA_TRACK_IMPL=$impl timeout 300 python bench.py --no-e2e --no-cpu > $O/c1_bench_$impl.log 2>&1
  echo "bench $impl rc=$?"; summ $O/c1_bench_$impl.log
done
for lib in variants/liborbit_b200_*.so; do
  [ -f "$lib" ] || continue
  tag=$(basename "$lib" .so)
  OA_LIB_PATH=$PWD/$lib OA_TEST_PJOIN=1 OA_TRACK_IMPL=pjoin timeout 300 python -m pytest \
      tests/test_gpu_zzz_pjoin.py -x -q -k "at_scale" > $O/c1_tests_$tag.log 2>&1
  echo "$tag parity rc=$?"
  OA_LIB_PATH=$PWD/$lib OA_TRACK_IMPL=pjoin timeout 300 python bench.py --no-e2e --no-cpu \
      > $O/c1_bench_$tag.log 2>&1
  echo "$tag bench rc=$?"; summ $O/c1_bench_$tag.log
done
for lag in 17 21; do
  OA_PJOIN_LAG=$((1<<lag)) OA_TRACK_IMPL=pjoin timeout 300 python bench.py --no-e2e --no-cpu > $O/c1_bench_lag$lag.log 2>&1
  echo "lag 2^$lag rc=$?"; summ $O/c1_bench_lag$lag.log
done
# full bench line (e2e + cpu baseline) for the default implementation
timeout 600 python bench.py > $O/c1_bench_full.log 2>&1; echo "full bench rc=$?"; tail -c 1500 $O/c1_bench_full.log
# ncu: launch lists, then one full capture of the partitioned-join kernel
for impl in pjoin hash; do
  OA_TRACK_IMPL=$impl timeout 300 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/c1_plain_$impl.log 2>&1 &&
  OA_TRACK_IMPL=$impl timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $O/c1_launches_$impl.csv python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 \
      > $O/c1_ncu_launches_$impl.log 2>&1
  echo "launch list $impl rc=$?"
done
OA_TRACK_IMPL=pjoin timeout 300 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/c1_plain2.log 2>&1 &&
OA_TRACK_IMPL=pjoin timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:oa_pjoin_kernel -s 4 -c 1 -o $O/c1_pjoin_full -f \
    python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/c1_ncu_full_pjoin.log 2>&1
echo "full capture rc=$?"
echo done
