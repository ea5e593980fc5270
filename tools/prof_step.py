import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nbody_orbit_analysis_b200.synth import DeviceSynth
from nbody_orbit_analysis_b200.tracker import OrbitTracker
gen = DeviceSynth(256**3, 1000)
K=8
snaps=[gen.snapshot(t) for t in range(K)]
cats=[gen.regions(t) for t in range(K)]
exists=np.arange(1000)
trk=OrbitTracker()
torch.cuda.synchronize()
import cProfile, pstats
def run():
    pend=None
    for t in range(K):
        t0=time.perf_counter()
        dev,n,off=snaps[t]; pos,rad,bulk=cats[t]
        p=trk.submit_device(dev,n,np.float32,np.int64,off,exists,pos,bulk,0.0,box_size=100.0)
        t1=time.perf_counter()
        if pend is not None: trk.collect(pend)
        t2=time.perf_counter()
        pend=p
        print('t=%d submit %.2f ms collect %.2f ms'%(t,(t1-t0)*1e3,(t2-t1)*1e3))
    trk.collect(pend)
run()
trk=OrbitTracker()
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
print(torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_stats()['num_device_alloc'], torch.cuda.memory_stats()['num_device_free'])
