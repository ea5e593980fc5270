#!/bin/bash
# usage: tools/scale_variants.sh N  -- exchange configurations at N GPUs, with the
# host-phase breakdown of rank 0 (where does a step go when it is not the kernel?)
N=$1
mkdir -p gpurun_out
echo "host cores: $(nproc)   memory: $(free -g | awk 'NR==2{print $2" GB"}')"
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-e2e --no-cpu --steps 20 \
      > gpurun_out/scale_${N}_$tag.log 2>&1
  python - "$tag" "$N" <<'PY'
import json,sys
tag,N=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open('gpurun_out/scale_%s_%s.log'%(N,tag)) if l.startswith('{')][-1]); r=d['roofline']
    print(tag, 'value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'],
          'kernel %.3f [%.3f..%.3f]'%(r['kernel_ms'],r['kernel_ms_min'],r['kernel_ms_max']),
          'events/step %.1f'%d['events_per_step'], 'host phases (ms/step)', d.get('host_phases_ms_per_step'))
except Exception as e:
    print(tag,'FAILED',e)
PY
}
run default
run default_again                       # events/step must equal the first run's
run batch8 OA_EXCHANGE_BATCH=8
run mainstream OA_EXCHANGE_STREAM=main OA_SM_RESERVE=0
run fewctas NCCL_MAX_CTAS=4 OA_SM_RESERVE=8
run nccl_info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH
grep -h -m 12 -E "NVLS|P2P|via|Channel 00" gpurun_out/scale_${N}_nccl_info.log | cut -c1-200
