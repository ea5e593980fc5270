"""Host/GPU timeline of the pipelined tracking step (1 GPU).
usage: python tools/prof_step.py [depth] [steps]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import deque
from nbody_orbit_analysis_b200.synth import DeviceSynth
from nbody_orbit_analysis_b200.tracker import OrbitTracker
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
gen = DeviceSynth(256**3, 1000)
snaps = [gen.snapshot(t) for t in range(K)]
cats = [gen.regions(t) for t in range(K)]
exists = np.arange(1000)
torch.cuda.synchronize()

def run(verbose):
    trk = OrbitTracker()
    q = deque(); rows = []
    T0 = time.perf_counter()
    for t in range(K):
        dev, n, off = snaps[t]; pos, rad, bulk = cats[t]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        p = trk.submit_device(dev, n, np.float32, np.int64, off, exists, pos, bulk, 0.0, box_size=100.0)
        e1.record()
        t1 = time.perf_counter()
        q.append((p, e0, e1))
        tc = 0.0
        if len(q) > depth:
            pp, a, b = q.popleft()
            c0 = time.perf_counter(); trk.collect(pp); tc = time.perf_counter() - c0
            rows.append((t, (t1 - t0) * 1e3, tc * 1e3, a.elapsed_time(b)))
    while q:
        pp, a, b = q.popleft(); trk.collect(pp)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - T0) * 1e3
    if verbose:
        for r in rows[-10:]:
            print('t=%d submit %.2f ms collect %.2f ms gpu(step) %.2f ms' % r)
        print('depth %d: %.3f ms/step wall over %d steps' % (depth, wall / K, K))
run(False)
run(True)
