"""Drop-in ``orbitanalysis.progenitors`` on B200.

Same two functions, arguments and return values as the reference
``progenitors.py:5-56`` (``get_central_particle_ids``) and ``:59-117``
(``find_main_progenitors``).  The numpy sorts / ``in1d(kind='table')`` /
per-descendant ``np.unique`` loops are replaced by device radix sorts, binary
search joins and one segmented arg-max through the C ABI (``csrc/oa_join.cu``,
``csrc/oa_sort.cu``); PyTorch owns the buffers only.
"""
import numpy as np
import torch

from . import _lib
from ._device import DeviceContext
from ._lib import lib, check, ptr


def get_central_particle_ids(snapshot, halo_positions, n=100, device=None):
    """IDs of the ``n`` particles closest to every halo centre, ordered by
    radius (reference ``progenitors.py:5-56``).

    ``snapshot``: dict with ``ids`` (N,), ``coordinates`` (N,3),
    ``region_offsets`` (n_halos,), optional ``box_size``.  Returns
    ``(central_ids, offsets)`` like the reference: blocks of
    ``min(n, block length)`` IDs and the start of every block.
    """
    ctx = DeviceContext(device)
    st = ctx.stream()
    ids = np.asarray(snapshot['ids'])
    coords = np.asarray(snapshot['coordinates'])
    N = len(ids)
    offsets = np.concatenate((np.asarray(snapshot['region_offsets'],
                                         dtype=np.int64), [N]))
    n_h = len(offsets) - 1
    lens = np.diff(offsets)
    take = np.minimum(lens, int(n))
    out_off = np.concatenate(([0], np.cumsum(take))).astype(np.int64)
    n_out = int(out_off[-1])
    if N == 0 or n_out == 0:
        return np.zeros(0, dtype=ids.dtype), out_off[:-1]

    centres = np.asarray(halo_positions)
    if centres.dtype not in (np.float32, np.float64):
        centres = centres.astype(np.float64)
    if coords.dtype not in (np.float32, np.float64):
        coords = coords.astype(np.float64)
    centre_f32 = centres.dtype == np.float32
    d_pos = ctx.upload(coords)
    d_ids = ctx.upload(ids, np.int64)
    d_off = ctx.upload(offsets)
    d_cen = ctx.upload(centres.astype(np.float64))
    periodic = 'box_size' in snapshot
    box = (_lib.C.c_double * 3)(0, 0, 0)
    if periodic:
        b = np.broadcast_to(np.asarray(snapshot['box_size'],
                                       dtype=np.float64), (3,))
        box[0], box[1], box[2] = float(b[0]), float(b[1]), float(b[2])
    r = ctx.empty(N, torch.float64)
    check(lib.oa_central_radii(
        ptr(d_pos), _lib.dtype_code(coords.dtype), ptr(d_off), n_h, ptr(d_cen),
        int(centre_f32), int(periodic), box, N, ptr(r), st))
    # radii are >= 0: their bit patterns order like the values
    order = ctx.segment_order(r.view(torch.int64), N, offsets)
    d_seg = d_off
    d_out_off = ctx.upload(out_off)
    out = ctx.empty(n_out, torch.int64)
    check(lib.oa_segment_heads(ptr(d_ids), ptr(order), ptr(d_seg),
                               ptr(d_out_off), n_h, n_out, ptr(out), st))
    ctx.launches += 2
    central = out[:n_out].cpu().numpy().astype(ids.dtype, copy=False)
    return central, out_off[:-1]


def find_main_progenitors(halo_pids, halo_offsets, tracked_pids,
                          tracked_offsets, device=None):
    """Main progenitor of every descendant = the halo holding the plurality of
    its tracked particles (reference ``progenitors.py:59-117``).

    A tracked ID that occurs more than once counts only at its first
    occurrence (``:82-84``); ties go to the smallest halo index and a
    descendant without any tracked particle in a halo gets -1 (``:107-115``).
    Returns a list of length ``len(tracked_offsets)``.
    """
    ctx = DeviceContext(device)
    st = ctx.stream()
    halo_pids = np.asarray(halo_pids)
    tracked_pids = np.asarray(tracked_pids)
    halo_offsets = np.ascontiguousarray(halo_offsets, dtype=np.int64)
    tracked_offsets = np.ascontiguousarray(tracked_offsets, dtype=np.int64)
    N, M = len(halo_pids), len(tracked_pids)
    n_halos, n_desc = len(halo_offsets), len(tracked_offsets)
    if n_desc == 0:
        return []
    if M == 0 or N == 0 or n_halos == 0:
        return [-1] * n_desc

    d_tr = ctx.upload(tracked_pids, np.int64)
    d_hp = ctx.upload(halo_pids, np.int64)

    # first occurrences of the tracked IDs (np.unique(return_index=True))
    t_keys, t_order, _ = ctx.argsort_values(d_tr, M)
    head = ctx.empty(M + 8, torch.int16)
    check(lib.oa_run_heads(ptr(t_keys), ptr(t_keys), M, ptr(head), st))
    first = ctx.empty(M + 8, torch.int16)
    check(lib.oa_scatter_flags(ptr(head), ptr(t_order), M, ptr(first), st))

    # membership + position in halo_pids (np.in1d + myin1d, kind='table')
    h_keys, h_order, h_min = ctx.argsort_values(d_hp, N)
    where = ctx.empty(M, torch.int64)
    check(lib.oa_lookup_sorted(ptr(h_keys), ptr(h_order), N, ptr(d_tr),
                               ptr(first), h_min, None, None, M, ptr(where),
                               st))

    # plurality vote per descendant
    d_hoff = ctx.upload(halo_offsets)
    d_toff = ctx.upload(tracked_offsets)
    keys = ctx.empty(M, torch.int64)
    check(lib.oa_vote_keys(ptr(where), ptr(d_hoff), n_halos, ptr(d_toff),
                           n_desc, M, ptr(keys), st))
    dummy = ctx.empty(M, torch.int64)
    # valid keys use 32 + bits(n_desc) bits; the "no halo" key ~0 must still
    # sort last, so the sort covers all 64 bits
    keys_sorted, _ = ctx.sort_pairs(keys, dummy, M, 64)
    best = ctx.empty(n_desc, torch.int64)
    out = ctx.empty(n_desc, torch.int64)
    check(lib.oa_vote_reduce(ptr(keys_sorted), M, n_desc, ptr(best), ptr(out),
                             st))
    ctx.launches += 6
    return list(out[:n_desc].cpu().numpy())
