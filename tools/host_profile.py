#!/usr/bin/env python
"""Host cost of one tracking step (Python + ctypes + torch bookkeeping) in the
CPU container: OrbitTracker.submit_device / collect with every kernel a no-op
(tests/fake_cuda.py) on the region table of BASELINE config[1].  What it does
NOT contain: the CUDA launch / memcpy / event calls themselves.

    python tools/host_profile.py [hash|pjoin] [--profile]
"""
import cProfile
import os
import pstats
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np          # noqa: E402
import torch                # noqa: E402
import fake_cuda            # noqa: E402
from nbody_orbit_analysis_b200.synth import SynthSim     # noqa: E402


class NoopLib(fake_cuda.FakeLib):
    def oa_track_fused(self, args, stream):
        return 0        # (the stand-in of the tests fills the mark array: 5 ms)

    def oa_pjoin_step(self, args, stream):
        return 0

    def oa_select_count(self, *a):
        return 0

    def oa_select_gather_events_ids(self, *a):
        return 0

    def oa_select_gather_events(self, *a):
        return 0

    def oa_segment_offsets(self, *a):
        return 0


def main():
    torch.set_num_threads(1)      # (CPU stand-in of cudaMemcpyAsync: no OpenMP team)
    impl = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in ('hash', 'pjoin') else 'hash'
    n_h, steps = 1000, 40
    sim = SynthSim(256 ** 3, n_h, 2, dtype=np.float32, catalogue_dtype=np.float32)
    lens = (sim.sizes * 0.81).astype(np.int64)
    offsets = np.concatenate(([0], np.cumsum(lens)))
    n = int(offsets[-1])
    exists = np.arange(n_h)
    pos = sim.halo_centre(0).astype(np.float32)
    bulk = sim.vh.astype(np.float32)
    dev = {'pos': torch.zeros(8), 'vel': torch.zeros(8),
           'ids': torch.zeros(8, dtype=torch.int64), 'mass': None}
    with fake_cuda.install(None) as fake:
        from nbody_orbit_analysis_b200 import tracker
        tracker.lib = NoopLib(fake._real, None)
        trk = tracker.OrbitTracker(device='cpu', impl=impl)

        def step():
            p = trk.submit_device(dev, n, np.float32, np.int64, offsets, exists,
                                  pos, bulk, 0.0, box_size=100.0)
            # (collect reads the small read-back: make it say "no events")
            if p.h_small is not None:
                p.h_small.zero_()
            return trk.collect(p)
        for _ in range(5):
            step()
        if '--profile' in sys.argv:
            prof = cProfile.Profile()
            prof.enable()
        dt = 1e9
        for _ in range(8):                 # noisy container: best of 8 batches
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            dt = min(dt, (time.perf_counter() - t0) / steps)
        if '--profile' in sys.argv:
            prof.disable()
            pstats.Stats(prof).sort_stats('tottime').print_stats(22)
    print('%s: %.3f ms of host work per step (%d regions, %d particles)'
          % (impl, dt * 1e3, n_h, n))


if __name__ == '__main__':
    main()
