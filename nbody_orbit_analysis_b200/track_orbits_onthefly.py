"""Drop-in ``track_orbits_onthefly.track_orbits`` on B200.

Same signature, callback protocol and one-file-per-snapshot layout as the
reference ``orbitanalysis/track_orbits_onthefly.py:8-9, 208-252`` (SURVEY.md
section 8(a-12)): two snapshots ``[s, s-1]`` are loaded per call, no Hubble
flow, bulk velocities always derived, frame and v_r in the data dtype, and the
outputs are the apsis IDs in previous-block order, the raw angle change of every
matched particle, and sorted entered / departed ID lists per halo.

The per-halo numpy loop (``:123-205``) is replaced by two launches of the fused
tracking kernel (previous snapshot, then current) plus ordered selections and a
segmented radix sort, all through the C ABI.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib, storage
from ._lib import lib, check, ptr
from .tracker import OrbitTracker, require_cuda


def repack(arr, length, inds):
    """Scatter ``arr`` into a -1-filled array of leading length ``length``
    (reference ``track_orbits_onthefly.py:61-68``; host-side catalogue rows)."""
    arr = np.asarray(arr)
    out = -np.ones((length,) + arr.shape[1:], dtype=arr.dtype)
    out[inds] = arr
    return out


def _full_begins(offsets, exists, n_halo):
    """Block start of every halo column; a missing column gets the start of the
    next existing block, so that offsets stay monotone and its segment is empty.
    """
    total = int(offsets[-1])
    begin = np.full(n_halo + 1, total, dtype=np.int64)
    begin[exists] = offsets[:-1]
    begin = np.minimum.accumulate(begin[::-1])[::-1]
    return np.ascontiguousarray(begin)


class _Lists:
    """Device-side helper: per-column ID lists out of a selection."""

    def __init__(self, trk):
        self.trk = trk

    def seg_offsets(self, sel, total, begin_full):
        """offsets (n_halo+1,) of a selection into the full column layout."""
        off = self.trk.segment_offsets(sel, total, begin_full[:-1])
        return np.concatenate((off, [total])).astype(np.int64)

    def sort_segments(self, d_ids, total, seg_off, sort_flag):
        """Sort ``d_ids[:total]`` ascending inside every segment whose flag is
        set; other segments keep their order.  Two stable radix passes."""
        trk = self.trk
        st = trk._stream()
        if total == 0:
            return d_ids
        n_seg = len(seg_off) - 1
        d_seg = trk._to_device(seg_off)
        d_flag = trk._to_device(np.ascontiguousarray(sort_flag, dtype=np.uint8))
        mm = trk._empty(2, torch.int64)
        check(lib.oa_minmax_i64(ptr(d_ids), total, ptr(mm), st))
        lo, hi = (int(v) for v in mm.cpu().tolist())
        lo_bits = max((max(hi - lo, total)).bit_length(), 1)
        hi_bits = max(int(n_seg).bit_length(), 1)
        k_lo, k_hi, idx = (trk._empty(total, torch.int64) for _ in range(3))
        check(lib.oa_segment_sort_keys(
            ptr(d_ids), total, ptr(d_seg), n_seg, ptr(d_flag), ptr(mm),
            ptr(k_lo), ptr(k_hi), ptr(idx), st))
        ws_bytes = lib.oa_sort_workspace_bytes(total)
        ws = trk._empty(ws_bytes, torch.uint8)
        k1, v1 = trk._empty(total, torch.int64), trk._empty(total, torch.int64)
        check(lib.oa_sort_pairs_u64(ptr(k_lo), ptr(idx), ptr(k1), ptr(v1),
                                    total, 0, lo_bits, ptr(ws), ws_bytes, st))
        k2 = trk._empty(total, torch.int64)
        check(lib.oa_gather_i64(ptr(k_hi), ptr(v1), total, None, ptr(k2), st))
        k3, v3 = trk._empty(total, torch.int64), trk._empty(total, torch.int64)
        check(lib.oa_sort_pairs_u64(ptr(k2), ptr(v1), ptr(k3), ptr(v3), total,
                                    0, hi_bits, ptr(ws), ws_bytes, st))
        out = trk._empty(total, torch.int64)
        check(lib.oa_gather_i64(ptr(d_ids), ptr(v3), total, None, ptr(out), st))
        trk.launches += 5 + 4 * (-(-lo_bits // 8) + -(-hi_bits // 8))
        return out


def track_orbits(snapshot_number, progenitor_links, regions,
                 load_snapshot_data, savefile, mode='pericentric', verbose=True,
                 device=None):
    """On-the-fly apsis detection between snapshots ``s-1`` and ``s``.

    Parameters are those of the reference (``track_orbits_onthefly.py:8-9``):
    ``progenitor_links`` is a ``(2, n_halo)`` integer array ``[ids at s, ids at
    s-1]`` with -1 where a halo does not exist; ``regions(s, halo_ids)`` returns
    ``(positions, radii)``; ``savefile`` contains a ``{}`` placeholder for the
    zero-padded snapshot number.
    """
    if mode not in ('pericentric', 'apocentric'):
        raise ValueError(
            "Orbit detection mode not recognized. Please specify either "
            "'pericentric' or 'apocentric'.")
    require_cuda()
    t_start = time.time()
    progenitor_links = np.asarray(progenitor_links)
    n_halo = progenitor_links.shape[1]

    # ---- callbacks, in the reference's order: s first, then s-1 --------------
    snaps, exists, pos_full, rad_full, pos_exist = [], [], [], [], []
    box_size = None
    for s, links in zip([snapshot_number, snapshot_number - 1],
                        progenitor_links):
        ex = np.flatnonzero(links != -1)
        pos, rad = regions(s, links[ex])
        pos, rad = np.asarray(pos), np.asarray(rad)
        pos_full.append(repack(pos, n_halo, ex))
        rad_full.append(repack(rad, n_halo, ex))
        snap = load_snapshot_data(s, pos, rad)
        snaps.append(snap)
        exists.append(ex)
        pos_exist.append(pos)
        box_size = snap['box_size'] if 'box_size' in snap else None

    ids_dtype = np.asarray(snaps[0]['ids']).dtype
    trk = OrbitTracker(mode=mode, device=device, onthefly=True)
    lists = _Lists(trk)
    st = trk._stream()

    # ---- previous snapshot, then the current one ------------------------------
    res_prev = trk.step(snaps[1], exists[1], pos_exist[1].reshape(-1, 3), None)
    gen_prev = trk.prev
    pend = trk.submit(snaps[0], exists[0], pos_exist[0].reshape(-1, 3), None,
                      diagnostics=True)
    res_cur = trk.collect_keep(pend)
    gen_cur = trk.prev
    n_prev, n_cur = gen_prev.n, gen_cur.n
    fdt = np.float64 if gen_cur.frame_f64 else np.float32

    begin_prev = _full_begins(gen_prev.offsets, exists[1], n_halo)
    begin_cur = _full_begins(gen_cur.offsets, exists[0], n_halo)
    prev_len = np.diff(begin_prev)
    has_prev = prev_len > 0            # reference: `if np.diff(sl_prev) > 0`

    def prev_list(op, value, gather_angles=False):
        sel, total = trk.select(gen_prev.mark, n_prev, op, value)
        off = lists.seg_offsets(sel, total, begin_prev)
        d_ids = trk._empty(max(total, 1), torch.int64)
        check(lib.oa_gather_record_ids(
            ptr(gen_prev.rec), int(gen_prev.frame_f64), ptr(sel), total, None,
            ptr(d_ids), st))
        ang = None
        if gather_angles:
            d_ang = trk._empty(max(total, 1), torch.float64 if
                               gen_cur.frame_f64 else torch.float32)
            check(lib.oa_gather_f(ptr(pend.dangle), int(gen_cur.frame_f64),
                                  ptr(sel), total, None, ptr(d_ang), st))
            ang = d_ang[:total].cpu().numpy()
        trk.launches += 1 + int(gather_angles)
        return d_ids, total, off, ang

    if n_prev > 0:
        ev_ids, n_ev, ev_off, _ = prev_list(_lib.OA_SEL_EQ, 1)
        ev_ids = ev_ids[:n_ev].cpu().numpy()
        _, n_m, _, angles = prev_list(_lib.OA_SEL_NE, _lib.OA_NO_EVENT, True)
        dep_ids, n_dep, dep_off, _ = prev_list(_lib.OA_SEL_EQ,
                                               _lib.OA_NO_EVENT)
        dep_ids = lists.sort_segments(dep_ids, n_dep, dep_off,
                                      np.ones(n_halo, dtype=np.uint8))
        dep_ids = dep_ids[:n_dep].cpu().numpy()
    else:
        ev_ids = dep_ids = np.zeros(0, dtype=np.int64)
        ev_off = dep_off = np.zeros(n_halo + 1, dtype=np.int64)
        angles = np.zeros(0, dtype=fdt)

    # entered: current particles without a match; sorted (setdiff1d) where the
    # halo had a previous block, in block order otherwise (:176-177)
    if n_cur > 0:
        marks = trk._empty(n_cur + 8, torch.int16)
        check(lib.oa_mark_unmatched(ptr(pend.diag['match']), n_cur,
                                    ptr(marks), st))
        sel, n_ent = trk.select(marks, n_cur, _lib.OA_SEL_EQ, 1)
        ent_off = lists.seg_offsets(sel, n_ent, begin_cur)
        d_ids = trk._empty(max(n_ent, 1), torch.int64)
        check(lib.oa_gather_i64(ptr(pend.keep[0]['ids']), ptr(sel), n_ent,
                                None, ptr(d_ids), st))
        trk.launches += 2
        ent_ids = lists.sort_segments(d_ids, n_ent, ent_off,
                                      has_prev.astype(np.uint8))
        ent_ids = ent_ids[:n_ent].cpu().numpy()
    else:
        ent_ids = np.zeros(0, dtype=np.int64)
        ent_off = np.zeros(n_halo + 1, dtype=np.int64)
    pend.keep = None

    # dtypes follow the reference's concatenations: a halo
    # without a previous block contributes empty arrays of the ID dtype
    parts = [np.dtype(fdt) if h else ids_dtype for h in has_prev]
    angles = angles.astype(np.result_type(*parts) if parts else fdt)

    def bulk_full(res, ex, dtype):
        out = np.full((n_halo, 3), np.nan, dtype=dtype)
        out[ex] = res.bulk_velocities
        return out
    vdt = res_cur.bulk_velocities.dtype
    bulk = [bulk_full(res_cur, exists[0], vdt),
            bulk_full(res_prev, exists[1], vdt)]

    tag = mode[:8] + 'er'
    t0 = time.time()
    with storage.File(savefile.format('%0.3d' % snapshot_number), 'w') as hf:
        hf.create_dataset(tag + '_offsets', data=ev_off)
        hf.create_dataset(tag + '_IDs', data=ev_ids.astype(ids_dtype))
        hf.create_dataset('angles', data=angles)
        hf.create_dataset('entered_offsets', data=ent_off)
        hf.create_dataset('entered_IDs', data=ent_ids.astype(ids_dtype))
        hf.create_dataset('departed_offsets', data=dep_off)
        hf.create_dataset('departed_IDs', data=dep_ids.astype(ids_dtype))
        hf.create_dataset('progenitor_links', data=progenitor_links)
        hf.create_dataset('region_radii', data=np.array(rad_full))
        hf.create_dataset('region_positions', data=np.array(pos_full))
        hf.create_dataset('bulk_velocities', data=np.array(bulk))
        if box_size is not None:
            hf.attrs['box_size'] = box_size
    if verbose:
        print('Saved to file in {} s\n'.format(time.time() - t0))
        print('Identified {}s in {} s\n'.format(tag, time.time() - t_start))
