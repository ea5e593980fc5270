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

The on-the-fly variant (``onthefly = 1``: frame and v_r in the data dtype, no
Hubble flow, raw angle change and 0 / 1 marks per matched previous particle,
reference ``track_orbits_onthefly.py:71-205``) and the small kernels the
on-the-fly driver calls around it (general ordered selection, gathers,
per-segment sort keys, radix sort) are restated as well.  Not covered: the hash
table (the twin matches by ID directly; ``tab`` stays untouched).
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
                    _arr(a.out_vr, n, fdt if a.onthefly else C.c_double),
                    _arr(a.out_r, n, fdt), _arr(a.out_match, n, C.c_int64))
            diag[3][:] = -1
        with np.errstate(all='ignore'):
            for j in range(n_h):
                row = rows[j]
                lo, hi = int(off[j]), int(off[j + 1])
                assert lo == row['cur_begin'] and hi - lo == row['cur_count']
                centre = row['centre_f'] if a.centre_f32 else row['centre']
                bulk = row['bulk_f'] if a.bulk_f32 else row['bulk']
                delta = snap['coordinates'][lo:hi] - centre
                if a.periodic:
                    delta = oracle.minimum_image(delta, snap['box_size'])
                if a.onthefly:
                    # track_orbits_onthefly.py:82-110: the work arrays have the
                    # snapshot's dtypes (a float64 centre / bulk velocity is
                    # rounded back on assignment), no Hubble term
                    delta = delta.astype(snap['coordinates'].dtype)
                    rv = (snap['velocities'][lo:hi] - bulk).astype(
                        snap['velocities'].dtype)
                    rads = np.sqrt(np.einsum('...i,...i', delta, delta))
                    rh = delta / rads[:, np.newaxis]
                    vr = np.einsum('...i,...i', rv, rh)
                else:
                    rh, vr, _ = oracle.region_frame(snap, (lo, hi), centre, bulk,
                                                    H)
                    rads = np.sqrt(np.einsum('...i,...i', delta, delta))
                assert rh.dtype == (np.float64 if f64 else np.float32), \
                    'frame dtype of the launch arguments'
                ang = np.zeros(hi - lo, dtype=np.float16)
                plo, cnt = int(row['prev_begin']), int(row['prev_count'])
                if plo >= 0 and rec_prev is not None:
                    pr = rec_prev[plo:plo + cnt]
                    d = oracle.compare_radial_velocities(
                        ids[lo:hi], pr['id'], vr, pr['vr'], rh, pr['rhat'], mode)
                    surviving = np.delete(np.arange(cnt), d['inds_departed'])
                    if a.onthefly:
                        mark_prev[plo + surviving] = 0
                        mark_prev[plo + surviving[d['apsis_inds']]] = 1
                        if a.dangle_prev:
                            _arr(a.dangle_prev, n_prev, fl)[plo + surviving] = \
                                d['angle_changes']
                    else:
                        ang, eang = oracle.calc_angles(hi - lo, pr['angle'], d)
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

    # ---- the small kernels of the on-the-fly driver --------------------------------
    @staticmethod
    def _hits(marks, n, op, value):
        m = _arr(marks, n, C.c_uint16)
        return np.flatnonzero((m == value) if op == 1 else (m != value))

    def oa_select_count(self, marks, n, op, value, ws, ws_bytes, total_dev, st):
        _arr(total_dev, 1, C.c_int64)[0] = len(self._hits(marks, n, op, value))
        return 0

    def oa_select_gather(self, marks, n, op, value, ws, sel_out, st):
        sel = self._hits(marks, n, op, value)
        _arr(sel_out, len(sel), C.c_int64)[:] = sel
        return 0

    @staticmethod
    def _count(n_sel, n_dev):
        return int(_arr(n_dev, 1, C.c_int64)[0]) if n_dev else int(n_sel)

    def oa_gather_record_ids(self, rec, frame_dtype, sel, n_sel, n_dev, out, st):
        k = self._count(n_sel, n_dev)
        s_ = _arr(sel, k, C.c_int64)
        stride = 8 if frame_dtype else 4
        n_rec = int(s_.max()) + 1 if k else 0
        _arr(out, k, C.c_int64)[:] = _arr(rec, n_rec * stride, C.c_int64)[s_ * stride]
        return 0

    def oa_gather_i64(self, src, sel, n_sel, n_dev, out, st):
        k = self._count(n_sel, n_dev)
        s_ = _arr(sel, k, C.c_int64)
        n_src = int(s_.max()) + 1 if k else 0
        _arr(out, k, C.c_int64)[:] = _arr(src, n_src, C.c_int64)[s_]
        return 0

    def oa_gather_f(self, src, dtype, sel, n_sel, n_dev, out, st):
        k = self._count(n_sel, n_dev)
        s_ = _arr(sel, k, C.c_int64)
        ct = C.c_double if dtype else C.c_float
        n_src = int(s_.max()) + 1 if k else 0
        _arr(out, k, ct)[:] = _arr(src, n_src, ct)[s_]
        return 0

    def oa_mark_unmatched(self, match, n, marks, st):
        _arr(marks, n, C.c_uint16)[:] = _arr(match, n, C.c_int64) < 0
        return 0

    def oa_minmax_i64(self, x, n, out, st):
        v = _arr(x, n, C.c_int64)
        o = _arr(out, 2, C.c_int64)
        o[0], o[1] = v.min(), v.max()
        return 0

    def oa_segment_sort_keys(self, ids, n, seg_off, n_seg, flag, minmax, key_lo,
                             key_hi, index, st):
        """csrc/oa_segment.cu: key_hi = segment, key_lo = id - min for segments
        to be sorted, else the element's own position."""
        off = _arr(seg_off, n_seg + 1, C.c_int64)
        i = np.arange(n, dtype=np.int64)
        seg = np.searchsorted(off[1:], i, side='right')
        sort = np.ones(n, dtype=bool) if not flag else \
            _arr(flag, n_seg, C.c_uint8)[seg] != 0
        lo = _arr(minmax, 2, C.c_int64)[0]
        _arr(key_lo, n, C.c_uint64)[:] = np.where(
            sort, _arr(ids, n, C.c_int64) - lo, i).astype(np.uint64)
        _arr(key_hi, n, C.c_uint64)[:] = seg.astype(np.uint64)
        _arr(index, n, C.c_uint64)[:] = i.astype(np.uint64)
        return 0

    def oa_sort_pairs_u64(self, keys_in, vals_in, keys_out, vals_out, n, begin_bit,
                          end_bit, ws, ws_bytes, st):
        """Stable LSD radix sort on bits [begin_bit, end_bit)."""
        k = _arr(keys_in, n, C.c_uint64)
        mask = np.uint64(((1 << (end_bit - begin_bit)) - 1) if end_bit - begin_bit < 64
                         else 0xFFFFFFFFFFFFFFFF)
        order = np.argsort((k >> np.uint64(begin_bit)) & mask, kind='stable')
        _arr(keys_out, n, C.c_uint64)[:] = k[order]
        _arr(vals_out, n, C.c_uint64)[:] = _arr(vals_in, n, C.c_uint64)[order]
        return 0
