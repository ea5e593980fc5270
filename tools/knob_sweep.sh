#!/bin/bash
# Usage: tools/knob_sweep.sh "0 1 2 4 8" [ncu]
# Runs the device-resident bench once per OA_TRACK_KNOBS value and prints the
# fused kernel's time; with "ncu" also the DRAM bytes of one launch.
for k in $1; do
  OA_TRACK_KNOBS=$k python bench.py --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/knob_$k.log 2>&1
  python - "$k" <<'PY'
import json,sys
k=sys.argv[1]
try:
    line=[l for l in open('gpurun_out/knob_%s.log'%k) if l.startswith('{')][-1]
    d=json.loads(line); r=d['roofline']
    print('knobs=%s kernel_ms=%.4f frac=%.3f step_ms=%.3f'%(k,r['kernel_ms'],r['frac'],d['ms_per_step']))
except Exception as e:
    print('knobs=%s FAILED %s'%(k,e))
PY
done
if [ "$2" == "ncu" ]; then
  for k in $1; do
    OA_TRACK_KNOBS=$k ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:oa_track_kernel -s 2 -c 1 --csv --log-file gpurun_out/knob_ncu_$k.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
    echo "knobs=$k $(grep -E 'dram__bytes|gpu__time' gpurun_out/knob_ncu_$k.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ' ')"
  done
fi
