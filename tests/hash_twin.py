"""numpy twin of ``oa_track_fused`` (csrc/oa_track.cu) for the CPU container.

TEST INFRASTRUCTURE ONLY.  The default implementation of the tracking step is a
CUDA kernel without a host build; ``tests/fake_cuda.py`` therefore ran its
branch of ``OrbitTracker`` with a no-op.  This module restates what the kernel
computes -- on the very buffers the host code hands it (``oa_track_args``,
``oa_region`` rows, 32 / 64-byte records, event marks) -- with the ORACLE's
per-region functions (reference ``track_orbits.py:247-351``), so that the host
side of the default path (region table, launch arguments, ring buffers, ordered
selection, result assembly, checkpoint / resume / device-state checkpoints) is
compared with the oracle on the CPU as well.  It says nothing about the CUDA
kernel itself: that is what the ``-m gpu`` tests are for.

Not covered: the on-the-fly arithmetic (``onthefly = 1``), the hash table (the
twin matches by ID directly; ``tab`` stays untouched).
"""
import ctypes as C

import numpy as np

import fake_cuda
from fake_cuda import NO_EVENT, _arr
from nbody_orbit_analysis_b200._lib import REGION_DTYPE
from oracle import orbit_oracle as oracle

REC32 = np.dtype([('id', np.int64), ('rhat', np.float32, (3,)),
                  ('vr', np.float32), ('r', np.float32),
                  ('angle', np.float16), ('flags', np.uint16)])
REC64 = np.dtype([('id', np.int64), ('rhat', np.float64, (3,)),
                  ('vr', np.float64), ('r', np.float64),
                  ('angle', np.float16), ('flags', np.uint16),
                  ('pad', np.uint32), ('tail', np.uint64)])
assert REC32.itemsize == 32 and REC64.itemsize == 64


class _OnePlus:
    """``1 + z`` exactly as the launch arguments carry it (the oracle writes
    ``1 + snapshot['redshift']``)."""

    def __init__(self, one_plus_z):
        self.v = one_plus_z

    def __radd__(self, other):
        assert other == 1
        return self.v


def _records(p, n, frame_f64):
    dt = REC64 if frame_f64 else REC32
    raw = _arr(p, n * dt.itemsize, C.c_uint8)
    return raw.view(dt)


def sign_faithful_f32(v):
    """float32 copy of v_r whose sign survives underflow (only the sign of the
    previous radial velocity is ever used, ``track_orbits.py:311-314``)."""
    out = v.astype(np.float32)
    lost = (out == 0) & (v != 0)
    out[lost] = np.copysign(np.float32(1e-38), v[lost]).astype(np.float32)
    return out


class TwinLib(fake_cuda.FakeLib):
    """``FakeLib`` whose ``oa_track_fused`` tracks for real (numpy)."""

    def oa_track_fused(self, args, stream):
        self.calls.append('oa_track_fused')
        a = args._obj
        if a.onthefly:
            raise NotImplementedError('the twin has no on-the-fly arithmetic')
        n, n_h = int(a.n_cur), int(a.n_regions)
        x64, f64 = bool(a.data_dtype), bool(a.frame_dtype)
        fl = C.c_double if x64 else C.c_float
        snap = {'coordinates': _arr(a.pos, 3 * n, fl).reshape(-1, 3),
                'velocities': _arr(a.vel, 3 * n, fl).reshape(-1, 3),
                'masses': 1.0, 'redshift': _OnePlus(a.one_plus_z)}
        if a.periodic:
            snap['box_size'] = np.array([a.box[0], a.box[1], a.box[2]])
        ids = _arr(a.ids, n, C.c_int64)
        rows = _arr(a.regions, n_h * 128, C.c_uint8).view(REGION_DTYPE)
        off = _arr(a.cur_off, n_h + 1, C.c_int64)
        rec = _records(a.rec_cur, n, f64)
        mark_cur = _arr(a.mark_cur, n, C.c_uint16)
        mark_cur[:] = NO_EVENT
        n_prev = int(a.n_prev) if a.rec_prev else 0
        rec_prev = _records(a.rec_prev, n_prev, f64) if n_prev else None
        mark_prev = _arr(a.mark_prev, n_prev, C.c_uint16) if n_prev else None
        mode = {0: 'pericentric', 1: 'apocentric'}[int(a.mode)]
        H = np.float64(a.hubble)      # hubble_parameter returns a numpy float64
        out_angle = _arr(a.out_angle, n, C.c_uint16) if a.out_angle else None
        diag = None
        if a.out_match:
            fdt = C.c_double if f64 else C.c_float
            diag = (_arr(a.out_rhat, 3 * n, fdt).reshape(-1, 3),
                    _arr(a.out_vr, n, C.c_double), _arr(a.out_r, n, fdt),
                    _arr(a.out_match, n, C.c_int64))
            diag[3][:] = -1
        with np.errstate(all='ignore'):
            for j in range(n_h):
                row = rows[j]
                lo, hi = int(off[j]), int(off[j + 1])
                assert lo == row['cur_begin'] and hi - lo == row['cur_count']
                centre = row['centre_f'] if a.centre_f32 else row['centre']
                bulk = row['bulk_f'] if a.bulk_f32 else row['bulk']
                rh, vr, _ = oracle.region_frame(snap, (lo, hi), centre, bulk, H)
                delta = snap['coordinates'][lo:hi] - centre
                if a.periodic:
                    delta = oracle.minimum_image(delta, snap['box_size'])
                rads = np.sqrt(np.einsum('...i,...i', delta, delta))
                assert rh.dtype == (np.float64 if f64 else np.float32), \
                    'frame dtype of the launch arguments'
                ang = np.zeros(hi - lo, dtype=np.float16)
                plo, cnt = int(row['prev_begin']), int(row['prev_count'])
                if plo >= 0 and rec_prev is not None:
                    pr = rec_prev[plo:plo + cnt]
                    d = oracle.compare_radial_velocities(
                        ids[lo:hi], pr['id'], vr, pr['vr'], rh, pr['rhat'], mode)
                    ang, eang = oracle.calc_angles(hi - lo, pr['angle'], d)
                    surviving = np.delete(np.arange(cnt), d['inds_departed'])
                    mark_prev[plo + surviving[d['apsis_inds']]] = \
                        eang.view(np.uint16)
                    if diag is not None:
                        diag[3][lo + d['inds_match']] = plo + surviving
                rec['id'][lo:hi] = ids[lo:hi]
                rec['rhat'][lo:hi] = rh
                rec['vr'][lo:hi] = vr if f64 else sign_faithful_f32(
                    np.asarray(vr, dtype=np.float64))
                rec['r'][lo:hi] = rads
                rec['angle'][lo:hi] = ang
                rec['flags'][lo:hi] = 0
                if out_angle is not None:
                    out_angle[lo:hi] = ang.view(np.uint16)
                if diag is not None:
                    diag[0][lo:hi] = rh
                    diag[1][lo:hi] = vr
                    diag[2][lo:hi] = rads
        return 0

    def oa_set_record_angles(self, rec, frame_dtype, angles, n, stream):
        self.calls.append('oa_set_record_angles')
        _records(rec, n, bool(frame_dtype))['angle'] = \
            _arr(angles, n, C.c_uint16).view(np.float16)
        return 0

    def oa_fill_u16(self, dst, n, value, stream):
        _arr(dst, n, C.c_uint16)[:] = value
        return 0
