"""Run the HOST code of ``OrbitTracker`` in the CPU container.

Test infrastructure only.  ``install`` makes ``torch.cuda``'s streams / events
no-ops on CPU tensors and replaces the compute entry points of the C ABI inside
``tracker`` by stand-ins that work on host pointers:

* ``oa_pjoin_step``          -> the g++ build of the very same stage code
  (tests/pjoin_emul), so ``OrbitTracker(impl='pjoin', device='cpu')`` produces
  real results that the tests compare with the oracle;
* ordered selection / offsets -> numpy restatements;
* ``oa_track_fused`` & co     -> no-ops (the hash-table kernel has no host twin:
  that branch is only checked to run through without Python errors).

Size queries (``*_workspace_bytes``, ``oa_index_bits`` ...) go to the real
library, which answers them without a device.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

NO_EVENT = 0x8000


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, enable_timing=False, **k):
        self.t = None

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t)


def _arr(p, n, ctype):
    addr = p.value if isinstance(p, C.c_void_p) else p
    if n <= 0 or not addr:
        return np.zeros(0, dtype=np.dtype(ctype))
    return np.ctypeslib.as_array((ctype * int(n)).from_address(addr))


class FakeLib:
    """Real library for host-only queries, stand-ins for the kernels."""

    def __init__(self, real, emul):
        self._real, self._emul = real, emul
        self.calls = []

    def __getattr__(self, name):
        return getattr(self._real, name)

    def oa_copy_async(self, dst, src, nbytes, stream):
        C.memmove(dst, src, int(nbytes))
        return 0

    oa_copy_small = oa_copy_async

    # ---- no-op kernels of the hash-table branch --------------------------------
    def oa_table_clear(self, *a):
        self.calls.append('oa_table_clear')
        return 0

    def oa_track_fused(self, args, stream):
        """No tracking; only what the next snapshot's selection relies on: the
        marks of this snapshot say "no event"."""
        self.calls.append('oa_track_fused')
        a = args._obj
        _arr(a.mark_cur, a.n_cur, C.c_uint16)[:] = NO_EVENT
        return 0

    def oa_bulk_velocity(self, vel, vel_dtype, mass, mass_dtype, cur_off, n_h, n,
                         round_f32, rows, bulk_out, ws, ws_bytes, st):
        """csrc/oa_bulk.cu reproduces numpy's summation order: here numpy
        itself, written to the rows and to bulk_out."""
        self.calls.append('oa_bulk_velocity')
        v = _arr(vel, 3 * n, C.c_double if vel_dtype else C.c_float).reshape(-1, 3)
        m = _arr(mass, n, C.c_double if mass_dtype else C.c_float) if mass else None
        off = _arr(cur_off, n_h + 1, C.c_int64)
        from nbody_orbit_analysis_b200._lib import REGION_DTYPE
        rec = _arr(rows, n_h * 128, C.c_uint8).view(REGION_DTYPE)
        out = _arr(bulk_out, 3 * n_h, C.c_double) if bulk_out else None
        with np.errstate(all='ignore'):
            for j in range(n_h):
                lo, hi = off[j], off[j + 1]
                # the reference's own expressions (track_orbits.py:270-280)
                if m is not None:
                    b = np.sum(m[lo:hi][:, np.newaxis] * v[lo:hi], axis=0) / \
                        np.sum(m[lo:hi])
                else:
                    b = np.mean(v[lo:hi], axis=0)
                b = b.astype(np.float64)
                rec['bulk'][j] = b
                rec['bulk_f'][j] = b
                if out is not None:
                    out[3 * j:3 * j + 3] = b
        return 0

    # ---- partitioned join: the emulated kernel --------------------------------
    def oa_pjoin_step(self, args, stream):
        self.calls.append('oa_pjoin_step')
        import os
        return self._emul.pj_emul_step(args, int(os.environ.get('OA_FAKE_CTAS', '3')))

    # ---- ordered selection (numpy) ---------------------------------------------
    def _sel(self, marks, n):
        m = _arr(marks, n, C.c_uint16)
        return np.flatnonzero(m != NO_EVENT), m

    def oa_select_count(self, marks, n, op, value, ws, ws_bytes, total_dev, st):
        assert op == 0 and value == NO_EVENT
        sel, _ = self._sel(marks, n)
        _arr(total_dev, 1, C.c_int64)[0] = len(sel)
        return 0

    def oa_select_gather_events_ids(self, marks, n, ws, ids, sel_out, ids_out,
                                    ang_out, st):
        sel, m = self._sel(marks, n)
        k = len(sel)
        _arr(sel_out, k, C.c_int64)[:] = sel
        _arr(ids_out, k, C.c_int64)[:] = _arr(ids, n, C.c_int64)[sel]
        _arr(ang_out, k, C.c_uint16)[:] = m[sel]
        return 0

    def oa_select_gather_events(self, marks, n, ws, rec, frame, sel_out, ids_out,
                                ang_out, st):
        sel, m = self._sel(marks, n)
        k = len(sel)
        stride = 8 if frame else 4            # record size in int64 words
        _arr(sel_out, k, C.c_int64)[:] = sel
        _arr(ids_out, k, C.c_int64)[:] = _arr(rec, n * stride, C.c_int64)[sel * stride]
        _arr(ang_out, k, C.c_uint16)[:] = m[sel]
        return 0

    def oa_segment_offsets(self, sel, n_sel, n_dev, seg_begin, n_seg, out, st):
        total = int(_arr(n_dev, 1, C.c_int64)[0]) if n_dev else int(n_sel)
        s = _arr(sel, min(total, int(n_sel)), C.c_int64)
        _arr(out, n_seg, C.c_int64)[:] = np.searchsorted(
            s, _arr(seg_begin, n_seg, C.c_int64))
        return 0


@contextlib.contextmanager
def install(emul, lib_class=None):
    """Patch torch.cuda and the tracker's library handle; restore on exit.
    ``lib_class``: a ``FakeLib`` subclass (tests/hash_twin.py: the default
    implementation tracked by a numpy twin instead of a no-op)."""
    from nbody_orbit_analysis_b200 import tracker, _lib
    saved = {}
    cuda = torch.cuda
    for name, val in (('is_available', lambda: True),
                      ('current_device', lambda: 0),
                      ('Stream', FakeStream), ('Event', FakeEvent),
                      ('stream', lambda s: contextlib.nullcontext()),
                      ('current_stream', lambda *a, **k: FakeStream()),
                      ('synchronize', lambda *a, **k: None)):
        saved[name] = getattr(cuda, name)
        setattr(cuda, name, val)
    saved['set_device'] = cuda.set_device
    cuda.set_device = lambda *a, **k: None
    real_empty, real_zeros, real_tensor = torch.empty, torch.zeros, torch.tensor

    def host(k):            # no pinned memory, no CUDA device in this container
        k.pop('pin_memory', None)
        if str(k.get('device', '')).startswith('cuda'):
            k['device'] = 'cpu'
        return k

    def empty(*a, **k):
        return real_empty(*a, **host(k))

    def zeros(*a, **k):
        return real_zeros(*a, **host(k))

    def tensor(*a, **k):
        return real_tensor(*a, **host(k))
    torch.empty, torch.zeros, torch.tensor = empty, zeros, tensor
    fake = (lib_class or FakeLib)(_lib.lib, emul)
    real_lib = tracker.lib
    tracker.lib = fake
    real_init = tracker.OrbitTracker.__init__

    def init(self, mode='pericentric', device=None, onthefly=False, impl=None):
        real_init(self, mode, 'cpu', onthefly, impl)
    tracker.OrbitTracker.__init__ = init
    try:
        yield fake
    finally:
        tracker.OrbitTracker.__init__ = real_init
        tracker.lib = real_lib
        torch.empty, torch.zeros, torch.tensor = real_empty, real_zeros, real_tensor
        for name, val in saved.items():
            setattr(cuda, name, val)
