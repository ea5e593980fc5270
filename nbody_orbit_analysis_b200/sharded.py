"""Multi-GPU execution of the tracking path: one process per GPU, particles
sharded by ID.

A particle's record in a halo's block at snapshot s and at s-1 land on the same
GPU (``gpu = id mod G``), so matching and carried state are purely local -- no
particle ever crosses NVLink.  Per snapshot the ranks exchange only

* the halo catalogue rows (broadcast from rank 0),
* per-halo event counts (all-reduce, gives the global ``region_offsets``),
* the event records ``(order key, ID, float16 angle)`` (all-gather, then a
  radix sort on the order key = position in the *unsharded* previous block,
  which reproduces the reference's event order, ``track_orbits.py:315-316``).

``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests of the exchange
logic) is the transport; the ordering step runs on the GPU through the C ABI.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import lib, check, ptr


def shard_snapshot(snapshot, rank, world):
    """Host-side helper for loaders that return the full snapshot: keep the
    particles with ``id mod world == rank`` (block order preserved) and return
    ``(local_snapshot, gpos)`` where ``gpos`` is each kept particle's index in
    the unsharded arrays (its global block position)."""
    ids = np.asarray(snapshot['ids'])
    keep = np.flatnonzero((ids % world) == rank)
    offs = np.append(np.asarray(snapshot['region_offsets'], dtype=np.int64),
                     len(ids))
    local = dict(snapshot)
    for k in ('ids', 'coordinates', 'velocities'):
        local[k] = np.ascontiguousarray(np.asarray(snapshot[k])[keep])
    if isinstance(snapshot.get('masses'), np.ndarray):
        local['masses'] = np.ascontiguousarray(snapshot['masses'][keep])
    local['region_offsets'] = np.searchsorted(keep, offs[:-1]).astype(np.int64)
    return local, keep.astype(np.int64)


class Comm:
    """Collectives of one tracking step."""

    def __init__(self, world=None, rank=None, device=None):
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) \
                if dist.get_backend() == 'nccl' else torch.device('cpu')
        self.device = device

    # -- catalogue ---------------------------------------------------------------
    def broadcast_catalogue(self, pos, rad, bulk):
        """Rank 0's (centres, radii, bulk velocities) to every rank."""
        pos, rad = np.asarray(pos), np.asarray(rad)
        has_bulk = bulk is not None
        n_h = len(rad)
        buf = np.zeros(7 * n_h, dtype=np.float64)
        if self.rank == 0:
            buf[:3 * n_h] = pos.reshape(-1)
            buf[3 * n_h:4 * n_h] = rad
            if has_bulk:
                buf[4 * n_h:] = np.asarray(bulk).reshape(-1)
        t = torch.from_numpy(buf).to(self.device)
        dist.broadcast(t, src=0)
        out = t.cpu().numpy()
        pos_o = out[:3 * n_h].reshape(n_h, 3).astype(pos.dtype)
        rad_o = out[3 * n_h:4 * n_h].astype(rad.dtype)
        bulk_o = out[4 * n_h:].reshape(n_h, 3).astype(
            np.asarray(bulk).dtype) if has_bulk else None
        return pos_o, rad_o, bulk_o

    # -- events --------------------------------------------------------------------
    def exchange_events(self, keys, ids, angles, local_counts):
        """All-gather variable-length event records and all-reduce the per-halo
        counts.  Device-agnostic (tensors live on ``self.device``).

        Returns ``(keys, ids, angles, global_counts)`` with the records of all
        ranks concatenated in rank order."""
        n_loc = torch.tensor([keys.numel()], dtype=torch.int64,
                             device=self.device)
        sizes = [torch.zeros_like(n_loc) for _ in range(self.world)]
        dist.all_gather(sizes, n_loc)
        sizes = [int(s.item()) for s in sizes]
        cap = max(max(sizes), 1)
        counts = local_counts.clone()
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)

        # one all-gather of packed (key, id, angle bits) records
        rec = torch.zeros((cap, 3), dtype=torch.int64, device=self.device)
        m = keys.numel()
        rec[:m, 0] = keys
        rec[:m, 1] = ids
        rec[:m, 2] = angles.to(torch.int64)
        out = [torch.empty_like(rec) for _ in range(self.world)]
        dist.all_gather(out, rec)
        rec = torch.cat([o[:s] for o, s in zip(out, sizes)])
        return (rec[:, 0].contiguous(), rec[:, 1].contiguous(),
                rec[:, 2].to(angles.dtype).contiguous(), counts)

    def merge_events(self, tracker, res):
        """Turn a rank-local ``StepResult`` into the global event lists (same
        on every rank), ordered like the unsharded reference run."""
        gen = res.prev_gen
        if gen is None or gen.gpos is None:
            raise _lib.OrbitB200Error(
                "sharded tracking needs the global block position of every "
                "particle (pass gpos= to step_device)")
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        E = res.n_events
        sel = res.apsis_prev_index
        keys = torch.empty(max(E, 1), dtype=torch.int64, device=self.device)
        check(lib.oa_gather_i64(ptr(gen.gpos), ptr(sel), E, None, ptr(keys), st))
        ids = torch.from_numpy(
            res.apsis_ids.astype(np.int64, copy=False)).to(self.device)
        ang = torch.from_numpy(
            res.apsis_angles.view(np.int16)).to(self.device)
        local_counts = torch.from_numpy(
            np.diff(res.apsis_offsets)).to(self.device)
        def order(keys_all):
            _, perm = self.sort_keys(keys_all, st)
            return perm

        def take(src, perm):
            out = torch.empty(max(perm.numel(), 1), dtype=src.dtype,
                              device=self.device)
            fn = lib.oa_gather_i64 if src.dtype == torch.int64 else \
                lib.oa_gather_u16
            check(fn(ptr(src), ptr(perm), perm.numel(), None, ptr(out), st))
            return out[:perm.numel()]
        ids_o, ang_o, offsets = self.merge(keys[:E], ids, ang, local_counts,
                                           order, take)
        tracker.launches += 3
        res.apsis_ids = ids_o.cpu().numpy().astype(gen.ids_dtype, copy=False)
        res.apsis_angles = ang_o.cpu().numpy().view(np.float16)
        res.apsis_offsets = offsets
        res.n_events = int(ids_o.numel())
        return res

    def merge(self, keys, ids, angles, local_counts, order, take):
        """Exchange + ordering step shared by the GPU path and the gloo tests.

        ``keys`` are the positions of the rank's event particles in the
        UNSHARDED previous snapshot; sorting the gathered records by key
        reproduces the reference order (``track_orbits.py:315-316``) because the
        blocks of the halos are contiguous there.  ``order(keys) -> perm`` and
        ``take(src, perm)`` are the device primitives (radix sort / gather
        through the C ABI on the GPU)."""
        keys, ids, angles, counts = self.exchange_events(
            keys, ids, angles, local_counts)
        if keys.numel():
            perm = order(keys)
            ids, angles = take(ids, perm), take(angles, perm)
        offsets = np.concatenate(
            ([0], np.cumsum(counts.cpu().numpy()))).astype(np.int64)
        return ids, angles, offsets

    def sort_keys(self, keys, st):
        """Radix sort of int64 order keys; returns (sorted keys, permutation)."""
        n = keys.numel()
        if n == 0:
            return keys, torch.empty(0, dtype=torch.int64, device=self.device)
        idx = torch.arange(n, dtype=torch.int64, device=self.device)
        k_out, v_out = torch.empty_like(keys), torch.empty_like(idx)
        ws_bytes = lib.oa_sort_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        hi = int(keys.max().item())
        bits = max(hi.bit_length(), 1)
        check(lib.oa_sort_pairs_u64(ptr(keys), ptr(idx), ptr(k_out),
                                    ptr(v_out), n, 0, bits, ptr(ws), ws_bytes,
                                    st))
        return k_out, v_out
