#!/usr/bin/env python
"""Host cost of one snapshot through the drop-in ``track_orbits()`` entry point
in the CPU container: the BASELINE config[1] shape (13.5 M region-particles,
1000 halos, 1.46 M events per snapshot), loader callbacks returning PAGEABLE
numpy arrays, every kernel a no-op (tests/fake_cuda.py), "device" copies =
host memcpys.  What it shows: the staging copy into the pinned ring, the driver's
Python, and the result-file write -- the host side of ``e2e_entry_point`` in
bench.py.  (The container's memory bandwidth is not the GPU box's: compare
variants, not absolute numbers.)

    python tools/host_profile_entry.py [--particles N] [--snapshots S] [--profile]
"""
import argparse
import cProfile
import ctypes as C
import os
import pstats
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np          # noqa: E402
import fake_cuda            # noqa: E402

EVENTS = 1460000
FIRST = 6       # timed from this snapshot on (the staging ring has been touched)


class NoopLib(fake_cuda.FakeLib):
    """No kernels; the selection reports EVENTS events spread over the halos."""

    def oa_pjoin_step(self, args, stream):
        return 0

    def oa_select_count(self, marks, n, op, value, ws, ws_bytes, total_dev, st):
        fake_cuda._arr(total_dev, 1, C.c_int64)[0] = min(EVENTS, n)
        return 0

    def oa_select_gather_events_ids(self, *a):
        return 0

    def oa_select_gather_events(self, *a):
        return 0

    def oa_segment_offsets(self, sel, n_sel, n_dev, seg_begin, n_seg, out, st):
        total = int(fake_cuda._arr(n_dev, 1, C.c_int64)[0])
        fake_cuda._arr(out, n_seg, C.c_int64)[:] = np.linspace(
            0, total, n_seg, endpoint=False).astype(np.int64)
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--particles', type=int, default=13560000)
    ap.add_argument('--halos', type=int, default=1000)
    ap.add_argument('--snapshots', type=int, default=14)
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--torch-copy', action='store_true',
                    help='stage with Tensor.copy_ (what the driver did before oa_host_copy)')
    a = ap.parse_args()
    n, n_h, n_s = a.particles, a.halos, a.snapshots
    rng = np.random.default_rng(1)
    offsets = np.linspace(0, n, n_h, endpoint=False).astype(np.int64)
    pos = rng.random((n_h, 3), dtype=np.float32) * 100
    rad = np.ones(n_h, dtype=np.float32)
    bulk = np.zeros((n_h, 3), dtype=np.float32)
    # two distinct sets of arrays, alternating (distinct source pages per step)
    host = []
    for k in range(2):
        host.append({'ids': rng.permutation(n).astype(np.int64),
                     'coordinates': rng.random((n, 3), dtype=np.float32),
                     'velocities': rng.random((n, 3), dtype=np.float32),
                     'masses': 1.0, 'region_offsets': offsets, 'box_size': 100.0,
                     'redshift': 0.0, 'H0': 0.0, 'Omega_m': 0.3, 'Omega_L': 0.7})
    marks = {}

    def regions(sn, halo_ids):
        return pos, rad, bulk

    def loader(sn, p, r):
        marks.setdefault(int(sn), time.perf_counter())
        return host[int(sn) % 2]
    tmp = tempfile.mkdtemp(prefix='oa_entry_prof_')
    mb = np.tile(np.arange(n_h, dtype=np.int64), (n_s, 1))
    spent = {'stage': 0.0, 'file': 0.0}
    with fake_cuda.install(None) as fake:
        from nbody_orbit_analysis_b200 import storage, tracker, track_orbits
        tracker.lib = NoopLib(fake._real, None)
        if a.torch_copy:
            tracker.STAGE_MIN_BYTES = 1 << 62
        # explicit timers around the two host costs of interest (the "device"
        # copies of the stand-in are host memcpys too and must not be counted)
        real_stage, real_file = tracker._stage_copy, storage.File

        def stage(dst, src):
            t0 = time.perf_counter()
            real_stage(dst, src)
            if marks.get(FIRST) is not None:
                spent['stage'] += time.perf_counter() - t0

        class TimedFile:
            def __init__(self, *a, **k):
                self.t0 = time.perf_counter()
                self.f = real_file(*a, **k)

            def __enter__(self):
                return self.f.__enter__()

            def __exit__(self, *exc):
                out = self.f.__exit__(*exc)
                if marks.get(FIRST) is not None:
                    spent['file'] += time.perf_counter() - self.t0
                return out
        tracker._stage_copy = stage
        storage.File = TimedFile
        prof = cProfile.Profile() if a.profile else None
        if prof:
            prof.enable()
        track_orbits.track_orbits(np.arange(n_s), mb, regions, loader,
                                  os.path.join(tmp, 'e.h5'), verbose=False,
                                  device='cpu')
        t1 = time.perf_counter()
        if prof:
            prof.disable()
            pstats.Stats(prof).sort_stats('tottime').print_stats(25)
    size = os.path.getsize(os.path.join(tmp, 'e.h5'))
    shutil.rmtree(tmp, ignore_errors=True)
    first = FIRST
    per = (t1 - marks[first]) / (n_s - first)
    tracker._stage_copy, storage.File = real_stage, real_file
    k = n_s - first
    print('%.2f ms per snapshot in the stand-in run, of which staging copy %.2f ms '
          '(%.1f GB/s, %d threads), result file %.2f ms  (%d particles = %.0f MB '
          'staged, %.1f MB written per snapshot)' % (
              per * 1e3, spent['stage'] / k * 1e3,
              n * 32 / (spent['stage'] / k) / 1e9, tracker.stage_threads(),
              spent['file'] / k * 1e3, n, n * 32 / 1e6, size / (n_s - 1) / 1e6))


if __name__ == '__main__':
    main()
