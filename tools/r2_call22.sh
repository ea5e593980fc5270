#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,P2P OA_EXCHANGE_PROFILE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --no-e2e --no-cpu \
      > $O/c22_${N}_info.log 2>&1
echo "rc=$?"
grep -h -E "via|NVLS|P2P|SHM|Channel 00|nChannels|Connected" $O/c22_${N}_info.log | cut -c1-180 | sort | uniq -c | sort -rn | head -25
python - $O/c22_${N}_info.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d['roofline']
print('value %.2f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'], 'kernel %.3f'%r['kernel_ms'], d.get('host_phases_ms_per_step'), 'exchange phases', d.get('exchange_phases_ms'))
PY
nvidia-smi topo -m | head -14
