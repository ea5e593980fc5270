#!/bin/bash
# 8-GPU diagnosis: where does the step go at N ranks?  every run under its own timeout
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; shift; extra=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $extra \
      > $O/c13_${N}_$tag.log 2>&1
  echo "$tag rc=$?"
}
run default "--steps 30 --no-e2e --no-cpu" OA_BENCH_RANK_CLOCKS=1
run noexchange "--steps 30 --no-e2e --no-cpu" OA_BENCH_NO_EXCHANGE=1 OA_BENCH_RANK_CLOCKS=1
run mainstream "--steps 30 --no-e2e --no-cpu" OA_EXCHANGE_STREAM=main OA_SM_RESERVE=0
run fewctas "--steps 30 --no-e2e --no-cpu" NCCL_MAX_CTAS=4 OA_SM_RESERVE=4
python - $O/c13_${N}_default.log $O/c13_${N}_noexchange.log $O/c13_${N}_mainstream.log $O/c13_${N}_fewctas.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
              'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
              'events/step %.1f'%d['events_per_step'],
              'host phases', d.get('host_phases_ms_per_step'))
        for pr in d.get('per_rank') or []:
            c = pr.get('clocks') or {}
            print('    rank', pr['rank'], 'kernel %.3f'%(pr['kernel_ms'] or 0), 'step_ms %.1f'%pr['step_ms'],
                  'sm_mhz', c.get('sm_mhz'), c.get('reasons'), 'samples', c.get('samples'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f).read()[-1500:])
PY
