"""Per-phase host timing of the sharded step (run under torchrun, 2+ GPUs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nbody_orbit_analysis_b200.synth import DeviceSynth
from nbody_orbit_analysis_b200.tracker import OrbitTracker
from nbody_orbit_analysis_b200 import sharded
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
gen = DeviceSynth(256**3, 1000, rank=rank, world=world)
K = 16
snaps = [gen.snapshot(t) for t in range(K)]
cats = [gen.regions(t) for t in range(K)]
exists = np.arange(1000)
comm = sharded.Comm(world, rank)
trk = OrbitTracker(); trk.events_on_device = True
torch.cuda.synchronize(); dist.barrier()
pend = None
hprev = None
for t in range(K):
    t0 = time.perf_counter()
    hb = nxt_b if t > 0 else comm.start_broadcast(*cats[t])
    if t + 1 < K: nxt_b = comm.start_broadcast(*cats[t + 1])
    pos, rad, bulk = comm.finish_broadcast(hb)
    t1 = time.perf_counter()
    dev, n, off = snaps[t]
    p = trk.submit_device(dev, n, np.float32, np.int64, off, exists, pos, bulk, 0.0, box_size=100.0, gpos=dev['gpos'])
    t2 = time.perf_counter()
    t3 = t4 = t5 = t2
    if pend is not None:
        res = trk.collect(pend)
        t3 = time.perf_counter()
        if res.apsis_offsets is not None:
            h = comm.start_merge(trk, res, to_host='slice')
            t4 = time.perf_counter()
            if hprev is not None:
                comm.finish_merge(hprev)
            hprev = h
        t5 = time.perf_counter()
    pend = p
    if rank == 0:
        print('t=%d bcast %.2f submit %.2f collect %.2f start %.2f finish %.2f ms' % (t, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3), flush=True)
dist.destroy_process_group()
