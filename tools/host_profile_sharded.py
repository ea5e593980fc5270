#!/usr/bin/env python
"""Host cost of one SHARDED tracking step in the CPU container: what a rank's
Python thread does per snapshot at N > 1 -- ``submit_device`` + ``collect`` +
``Comm.start_merge`` + ``Comm.finish_merge`` (all-to-all path, results left in
HBM) -- with every kernel and every collective a no-op.  At 8 GPUs the step is
bound by exactly this (profiles/r02_scaling.md: host phases 0.78 ms of a
0.81 ms step).  Not contained: the CUDA / NCCL launch calls themselves.

    python tools/host_profile_sharded.py [--world 8] [--profile]
"""
import argparse
import cProfile
import ctypes as C
import os
import pstats
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np          # noqa: E402
import torch                # noqa: E402
import torch.distributed as dist   # noqa: E402
import fake_cuda            # noqa: E402
from nbody_orbit_analysis_b200.synth import SynthSim     # noqa: E402

EVENTS = 1460000


class NoopLib(fake_cuda.FakeLib):
    def oa_track_fused(self, args, stream):
        return 0

    def oa_select_count(self, marks, n, op, value, ws, ws_bytes, total_dev, st):
        fake_cuda._arr(total_dev, 1, C.c_int64)[0] = min(EVENTS, n)
        return 0

    def oa_select_gather_events_ids(self, *a):
        return 0

    def oa_select_gather_events(self, *a):
        return 0

    def oa_segment_offsets(self, sel, n_sel, n_dev, seg_begin, n_seg, out, st):
        return 0

    # exchange kernels: only the numbers the host reads back
    def oa_split_quantiles(self, *a):
        return 0

    def oa_pack_split(self, gpos, sel, ids, ang, small, n_seg, prop, world, cap,
                      bnd, out, counts, st):
        fake_cuda._arr(counts, n_seg, C.c_int64)[:] = EVENTS // max(n_seg, 1)
        return 0

    def oa_merge_blocks(self, recv, world, cap, ids_out, ang_out, info, st):
        i = fake_cuda._arr(info, 2, C.c_int64)
        i[0], i[1] = EVENTS, EVENTS // world
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--profile', action='store_true')
    a = ap.parse_args()
    W = a.world
    torch.set_num_threads(1)
    n_h, steps = 1000, 40
    sim = SynthSim(256 ** 3, n_h, 2, dtype=np.float32, catalogue_dtype=np.float32)
    lens = (sim.sizes * 0.81).astype(np.int64)
    offsets = np.concatenate(([0], np.cumsum(lens)))
    n = int(offsets[-1])
    exists = np.arange(n_h)
    pos = sim.halo_centre(0).astype(np.float32)
    bulk = sim.vh.astype(np.float32)
    dev = {'pos': torch.zeros(8), 'vel': torch.zeros(8),
           'ids': torch.zeros(8, dtype=torch.int64), 'mass': None,
           'gpos': torch.zeros(8, dtype=torch.int64)}

    # collectives: what one rank's call costs in Python is NOT what is measured
    # here -- they are replaced by the cheapest thing that keeps the numbers the
    # host reads consistent (every rank sent the same meta block)
    def all_gather(out, inp, **k):
        out.view(W, -1)[:] = inp
    dist.all_gather_into_tensor = all_gather
    dist.all_to_all_single = lambda out, inp, **k: None

    def all_reduce(t, **k):
        return None
    dist.all_reduce = all_reduce
    with fake_cuda.install(None) as fake:
        from nbody_orbit_analysis_b200 import sharded, tracker
        noop = NoopLib(fake._real, None)
        tracker.lib = noop
        sharded.lib = noop
        trk = tracker.OrbitTracker(device='cpu')
        trk.events_on_device = True
        comm = sharded.Comm(W, 0, device=torch.device('cpu'))
        phases = {}

        def timed(name, fn, *args, **kw):
            t0 = time.perf_counter()
            out = fn(*args, **kw)
            phases[name] = phases.get(name, 0.0) + time.perf_counter() - t0
            return out
        state = {'pending': None, 'inflight': None}

        def step():
            p = timed('submit', trk.submit_device, dev, n, np.float32, np.int64,
                      offsets, exists, pos, bulk, 0.0, box_size=100.0,
                      gpos=dev['gpos'])
            prev, state['pending'] = state['pending'], p
            if prev is None:
                return
            res = timed('collect', trk.collect, prev)
            if res.apsis_offsets is None:
                return
            h = timed('start_merge', comm.start_merge, trk, res, to_host=False)
            old, state['inflight'] = state['inflight'], h
            if old is not None:
                timed('finish_merge', comm.finish_merge, old)
        for _ in range(6):
            step()
        prof = None
        if a.profile:
            prof = cProfile.Profile()
            prof.enable()
        best = None
        for _ in range(8):                 # noisy container: best of 8 batches
            phases.clear()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            dt = (time.perf_counter() - t0) / steps
            if best is None or dt < best[0]:
                best = (dt, {k: v / steps for k, v in phases.items()})
        if prof is not None:
            prof.disable()
            pstats.Stats(prof).sort_stats('tottime').print_stats(30)
    print('world %d: %.3f ms of host work per step: %s' % (
        W, best[0] * 1e3,
        ', '.join('%s %.3f' % (k, v * 1e3) for k, v in best[1].items())))


if __name__ == '__main__':
    main()
