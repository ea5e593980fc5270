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
        self.stream = None           # CUDA stream of the exchange (NCCL path)

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
        if self.device.type == 'cuda':
            # on the exchange stream: independent of the kernels in flight
            if self.stream is None:
                self.stream = torch.cuda.Stream(self.device)
            with torch.cuda.stream(self.stream):
                t = torch.from_numpy(buf).to(self.device)
                dist.broadcast(t, src=0)
                out = t.cpu().numpy()
        else:
            t = torch.from_numpy(buf)
            dist.broadcast(t, src=0)
            out = t.numpy()
        pos_o = out[:3 * n_h].reshape(n_h, 3).astype(pos.dtype)
        rad_o = out[3 * n_h:4 * n_h].astype(rad.dtype)
        bulk_o = out[4 * n_h:].reshape(n_h, 3).astype(
            np.asarray(bulk).dtype) if has_bulk else None
        return pos_o, rad_o, bulk_o

    # -- events --------------------------------------------------------------------
    def exchange_events(self, keys, ids, angles, local_counts,
                        return_sizes=False):
        """Two collectives per snapshot: an all-gather of every rank's
        ``[per-halo event counts | number of events]`` (summed on the host it
        is the all-reduce that gives the global ``region_offsets``) and an
        all-gather of the packed, padded event records.  Device-agnostic
        (tensors live on ``self.device``).

        Returns ``(keys, ids, angles, global_counts)`` with the records of all
        ranks concatenated in rank order; ``global_counts`` is a host array."""
        m = keys.numel()
        meta = torch.cat((local_counts.to(torch.int64).reshape(-1),
                          torch.tensor([m], dtype=torch.int64,
                                       device=self.device)))
        meta_all = torch.empty(self.world * meta.numel(), dtype=torch.int64,
                               device=self.device)
        dist.all_gather_into_tensor(meta_all, meta)
        meta_all = meta_all.cpu().numpy().reshape(self.world, -1)  # one sync
        sizes = [int(v) for v in meta_all[:, -1]]
        counts = meta_all[:, :-1].sum(axis=0)
        cap = max(max(sizes), 1)

        rec = torch.empty((cap, 3), dtype=torch.int64, device=self.device)
        rec[:m, 0] = keys
        rec[:m, 1] = ids
        rec[:m, 2] = angles.to(torch.int64)
        rec_all = torch.empty(self.world * cap * 3, dtype=torch.int64,
                              device=self.device)
        dist.all_gather_into_tensor(rec_all, rec.reshape(-1))
        rec_all = rec_all.reshape(self.world * cap, 3)
        if any(sz != cap for sz in sizes):          # drop the padding
            rec = torch.cat([rec_all[r * cap:r * cap + sz]
                             for r, sz in enumerate(sizes)])
        else:
            rec = rec_all
        out = (rec[:, 0].contiguous(), rec[:, 1].contiguous(),
               rec[:, 2].to(angles.dtype).contiguous(), counts)
        return out + (sizes,) if return_sizes else out

    def merge_events(self, tracker, res, to_host=True):
        """Turn a rank-local ``StepResult`` (events left in HBM, see
        ``OrbitTracker.events_on_device``) into the global event lists, ordered
        like the unsharded reference run.  Every rank's list is already
        ascending in the order key, so the ordering step is a multi-way merge
        (``oa_merge_event_lists``), not a sort.  With ``to_host=False`` the
        merged lists stay on the device (``res.d_ids`` / ``res.d_ang``): only
        the rank that writes the result file needs them on the host."""
        gen = res.prev_gen
        if gen is None or gen.gpos is None:
            raise _lib.OrbitB200Error(
                "sharded tracking needs the global block position of every "
                "particle (pass gpos= to step_device)")
        # The exchange runs on its own stream, after this snapshot's compaction
        # only: it overlaps the kernels of the NEXT snapshot, which the caller
        # has already submitted on the main stream.
        if self.stream is None:
            self.stream = torch.cuda.Stream(self.device)
        self.stream.wait_event(res.compacted)
        with torch.cuda.stream(self.stream):
            st = C.c_void_p(self.stream.cuda_stream)
            E = res.n_events
            if res.d_ids is None:
                res.d_ids = torch.from_numpy(
                    res.apsis_ids.astype(np.int64, copy=False)).to(self.device)
                res.d_ang = torch.from_numpy(
                    res.apsis_angles.view(np.int16)).to(self.device)
            keys = torch.empty(max(E, 1), dtype=torch.int64,
                               device=self.device)
            check(lib.oa_gather_i64(ptr(gen.gpos), ptr(res.apsis_prev_index),
                                    E, None, ptr(keys), st))
            local_counts = torch.from_numpy(
                np.diff(res.apsis_offsets)).to(self.device, non_blocking=True)
            k_all, i_all, a_all, counts, sizes = self.exchange_events(
                keys[:E], res.d_ids[:E], res.d_ang[:E], local_counts,
                return_sizes=True)
            total = int(k_all.numel())
            ids_o = torch.empty(max(total, 1), dtype=torch.int64,
                                device=self.device)
            ang_o = torch.empty(max(total, 1), dtype=torch.int16,
                                device=self.device)
            list_off = torch.from_numpy(np.concatenate(
                ([0], np.cumsum(sizes))).astype(np.int64)).to(
                    self.device, non_blocking=True)
            check(lib.oa_merge_event_lists(
                ptr(k_all), ptr(i_all), ptr(a_all), total, ptr(list_off),
                self.world, ptr(ids_o), ptr(ang_o), st))
            tracker.launches += 2
            res.d_ids, res.d_ang = ids_o[:total], ang_o[:total]
            res.n_events = total
            res.apsis_offsets = np.concatenate(
                ([0], np.cumsum(counts))).astype(np.int64)
            # the ring buffers read above may be rewritten only after this
            done = torch.cuda.Event()
            done.record(self.stream)
            tracker.wait_before_submit = done
        if to_host:
            # asynchronous: res.wait_host() before reading apsis_ids / angles
            h_ids, h_ang, ready = tracker.to_host_async(
                res.d_ids, res.d_ang, stream=self.stream)
            res.apsis_ids = h_ids.numpy().astype(gen.ids_dtype, copy=False)
            res.apsis_angles = h_ang.numpy().view(np.float16)
            res.host_ready = ready
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
        offsets = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
        return ids, angles, offsets
