#!/bin/bash
# pj2 tuning call: parity + bench of the default build and of every variants/liborbit_b200_pj2*.so
#   gpurun --timeout 900 -- bash tools/r2_pj2_variants.sh TAG [full]
set -u
TAG=${1:-x}
mkdir -p gpurun_out
O=gpurun_out
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
timeout 600 python -m pytest tests/test_gpu_zzz_pjoin.py -q -k pj2 > $O/${TAG}_tests_pj2.log 2>&1; echo "pj2 tests rc=$?"
tail -n 3 $O/${TAG}_tests_pj2.log
OA_TRACK_IMPL=pj2 timeout 300 python bench.py --no-e2e --no-cpu > $O/${TAG}_bench_pj2.log 2>&1
echo "bench pj2 rc=$?"; summ $O/${TAG}_bench_pj2.log
for lib in variants/liborbit_b200_pj2*.so; do
  [ -f "$lib" ] || continue
  tag=$(basename "$lib" .so | sed 's/liborbit_b200_//')
  OA_LIB_PATH=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_zzz_pjoin.py -x -q -k "pj2 and at_scale" \
      > $O/${TAG}_tests_$tag.log 2>&1
  echo "$tag parity rc=$?"
  OA_LIB_PATH=$PWD/$lib OA_TRACK_IMPL=pj2 timeout 300 python bench.py --no-e2e --no-cpu \
      > $O/${TAG}_bench_$tag.log 2>&1
  echo "$tag bench rc=$?"; summ $O/${TAG}_bench_$tag.log
done
if [ "${2:-}" = "full" ]; then
  OA_TRACK_IMPL=pj2 timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:oa_pj2_kernel -s 4 -c 1 -o $O/${TAG}_pj2_full -f \
      python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > $O/${TAG}_ncu_full_pj2.log 2>&1
  echo "full capture rc=$?"
fi
