#!/bin/bash
# First GPU call of round 2 (1 GPU):  gpurun --timeout 1500 -- tools/round2_first_call.sh
# 1. the regular GPU suite (must be green before anything else is looked at),
# 2. the partitioned-join kernel on hardware for the first time: opt-in parity
#    tests under a hard timeout (a dependency-wait bug traps, it must not hang),
# 3. bench of both implementations, launch list + one full ncu capture each.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu suite rc=$?"
OA_TEST_PJOIN=1 timeout 600 python -m pytest tests/test_gpu_zzz_pjoin.py -x -q \
    > gpurun_out/r2_tests_pjoin.log 2>&1; echo "pjoin tests rc=$?"
tail -n 5 gpurun_out/r2_tests_gpu.log gpurun_out/r2_tests_pjoin.log
for impl in hash pjoin; do
  OA_TRACK_IMPL=$impl timeout 600 python bench.py --no-e2e --no-cpu \
      > gpurun_out/r2_bench_$impl.log 2>&1; echo "bench $impl rc=$?"
  python - "$impl" <<'PY'
import json, sys
impl = sys.argv[1]
try:
    d = json.loads([l for l in open('gpurun_out/r2_bench_%s.log' % impl) if l.startswith('{')][-1])
    r = d['roofline']
    print(impl, 'value %.2f G' % (d['value'] / 1e9), 'ms/step %.3f' % d['ms_per_step'],
          'kernel %.3f ms' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'events', d['events_per_step'])
except Exception as e:
    print(impl, 'FAILED', e)
PY
done
# tuning builds of the same kernel, if present (build them here, before the call):
#   python -m nbody_orbit_analysis_b200._build --out variants/liborbit_b200_small.so \
#     -DOA_PJOIN_THREADS=256 -DOA_PJOIN_MIN_CTAS=4 -DOA_PJOIN_TILE=1024 \
#     -DOA_PJOIN_REC_CAP=1408 -DOA_PJOIN_TARGET=1152
#   python -m nbody_orbit_analysis_b200._build --out variants/liborbit_b200_tma.so -DOA_PJOIN_TMA=1
for lib in variants/liborbit_b200_*.so; do
  [ -f "$lib" ] || continue
  tag=$(basename "$lib" .so)
  OA_LIB_PATH=$PWD/$lib OA_TEST_PJOIN=1 OA_TRACK_IMPL=pjoin timeout 300 python -m pytest \
      tests/test_gpu_zzz_pjoin.py -x -q -k "at_scale" > gpurun_out/r2_tests_$tag.log 2>&1
  echo "$tag parity rc=$?"
  OA_LIB_PATH=$PWD/$lib OA_TRACK_IMPL=pjoin timeout 600 python bench.py --no-e2e --no-cpu \
      > gpurun_out/r2_bench_$tag.log 2>&1
  echo "$tag bench rc=$?"
  grep -o '"value": [0-9.e+]*\|"kernel_ms": [0-9.]*' gpurun_out/r2_bench_$tag.log | head -2
done
for impl in hash pjoin; do
  OA_TRACK_IMPL=$impl timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
      --log-file gpurun_out/r2_launches_$impl.csv python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 \
      > gpurun_out/r2_ncu_launches_$impl.log 2>&1
done
OA_TRACK_IMPL=pjoin timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:oa_pjoin_kernel -s 4 -c 1 -o gpurun_out/r2_pjoin_full -f \
    python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/r2_ncu_full_pjoin.log 2>&1
echo done
