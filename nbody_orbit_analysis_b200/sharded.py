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


class _Exchange:
    """Handle of an exchange in flight (``Comm.start_merge``)."""
    pass


class _HostStream:
    """Stand-in for a CUDA stream / event when the communicator lives on the CPU
    (gloo): the CPU tests drive the very same exchange code with host tensors
    and numpy stand-ins for the kernels."""
    cuda_stream = 0

    def wait_event(self, event):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


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
        self.cat_stream = None       # stream / communicator of the catalogue
        self.cat_group = None
        self._cat_pinned, self._cat_turn = {}, {}
        self._cap = None             # records per rank in the send buffers
        self._last_total = 0
        self._results = 0            # finished exchanges: pinned slot of their lists
        self._splitters = None       # all ranks' quantile keys of the last exchange
        import os
        self.batch_size = max(1, min(31, int(os.environ.get('OA_EXCHANGE_BATCH', '1'))))
        self._open, self._launched, self._batches = None, [], 0
        self._stage_sets = {}
        self._repeats = 0            # consecutive repeats after an overflow

    # -- catalogue ---------------------------------------------------------------
    def start_broadcast(self, pos, rad, bulk):
        """Begin broadcasting rank 0's (centres, radii, bulk velocities).  On the
        NCCL path the broadcast has its own communicator and stream, so it does
        not queue behind the event exchange; call it one snapshot ahead and
        ``finish_broadcast`` costs nothing."""
        pos, rad = np.asarray(pos), np.asarray(rad)
        has_bulk = bulk is not None
        n_h = len(rad)
        h = _Exchange()
        h.meta = (pos.dtype, rad.dtype,
                  np.asarray(bulk).dtype if has_bulk else None, n_h)
        if self.device.type == 'cuda':
            if self.cat_stream is None:
                self.cat_stream = torch.cuda.Stream(self.device)
                self.cat_group = dist.new_group(backend='nccl')
            host = self._cat_buffer(n_h)
            buf = host.numpy()
            buf[:] = 0
        else:
            host = None
            buf = np.zeros(7 * n_h, dtype=np.float64)
        if self.rank == 0:
            buf[:3 * n_h] = pos.reshape(-1)
            buf[3 * n_h:4 * n_h] = rad
            if has_bulk:
                buf[4 * n_h:] = np.asarray(bulk).reshape(-1)
        if host is not None:
            with torch.cuda.stream(self.cat_stream):
                t = host.to(self.device, non_blocking=True)
                dist.broadcast(t, src=0, group=self.cat_group)
                host.copy_(t, non_blocking=True)
                h.ready = torch.cuda.Event()
                h.ready.record(self.cat_stream)
            h.host, h.keep = host, t
        else:
            t = torch.from_numpy(buf)
            dist.broadcast(t, src=0)
            h.host, h.ready = t, None
        return h

    CAT_RING = 4

    def _cat_buffer(self, n_h, make=None):
        """Pinned staging buffer of the next catalogue broadcast: a ring of
        CAT_RING buffers per catalogue length, used strictly in turn (pinning
        costs ~0.1 ms, so nothing is pinned in steady state).  A buffer comes
        round again CAT_RING broadcasts later; at most two are ever in flight."""
        ring = self._cat_pinned.setdefault(n_h, [])
        turn = self._cat_turn.get(n_h, 0)
        self._cat_turn[n_h] = turn + 1
        if len(ring) < self.CAT_RING:
            make = make or (lambda: torch.zeros(
                7 * n_h, dtype=torch.float64, pin_memory=True))
            ring.append(make())
        return ring[turn % self.CAT_RING]

    def finish_broadcast(self, h):
        if h.ready is not None:
            h.ready.synchronize()
        out = h.host.numpy()
        pdt, rdt, bdt, n_h = h.meta
        pos_o = out[:3 * n_h].reshape(n_h, 3).astype(pdt)
        rad_o = out[3 * n_h:4 * n_h].astype(rdt)
        bulk_o = out[4 * n_h:].reshape(n_h, 3).astype(bdt) \
            if bdt is not None else None
        return pos_o, rad_o, bulk_o

    def broadcast_catalogue(self, pos, rad, bulk):
        """Rank 0's (centres, radii, bulk velocities) to every rank."""
        return self.finish_broadcast(self.start_broadcast(pos, rad, bulk))

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

    # -- events: asynchronous exchange (NCCL path) ---------------------------------
    HEADROOM = 2.0       # send-buffer capacity / largest event list last snapshot

    def start_merge(self, tracker, res, to_host=True):
        """Enqueue the exchange of one snapshot's events and return a handle for
        ``finish_merge``.  Nothing here waits for the GPU: the send buffer is
        packed by a kernel from the tracker's device arrays (event count and
        per-halo offsets included), ONE all-gather moves every rank's buffer,
        and a kernel merges the gathered (ascending) lists by order key and
        derives the global ``region_offsets``.  The capacity of the send buffers
        comes from the previous snapshot's sizes; a snapshot whose event list
        outgrows it is detected in ``finish_merge`` and exchanged again."""
        gen = res.prev_gen
        if gen is None or gen.gpos is None:
            raise _lib.OrbitB200Error(
                "sharded tracking needs the global block position of every "
                "particle (pass gpos= to step_device)")
        if res.d_small is None:
            raise _lib.OrbitB200Error(
                "start_merge needs OrbitTracker.events_on_device = True")
        if self.stream is None:
            # OA_EXCHANGE_STREAM=main: enqueue the exchange behind the kernels
            # already submitted instead of beside them (no SM sharing between
            # NCCL and the persistent tracking kernel)
            import os
            if self.device.type != 'cuda':
                self.stream = _HostStream()
            elif os.environ.get('OA_EXCHANGE_STREAM') == 'main':
                self.stream = torch.cuda.current_stream(self.device)
            else:
                self.stream = torch.cuda.Stream(self.device)
        if self._cap is None:
            # first exchange: agree on a capacity (the only blocking collective)
            t = torch.tensor([res.n_events], dtype=torch.int64,
                             device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self._cap = self._round_cap(int(t.item()))
        if to_host is True:
            # a single writer needs the whole merged list: all-gather path
            return self._launch_merge(tracker, res, self._cap, to_host)
        # every rank keeps (and writes) its own key range: all-to-all path, the
        # volume per rank does not grow with the number of GPUs
        return self._launch_split(tracker, res, self._block_cap(), to_host)

    def _on_stream(self):
        return torch.cuda.stream(self.stream) if self.device.type == 'cuda' \
            else self.stream

    def _event(self):
        return torch.cuda.Event() if self.device.type == 'cuda' \
            else _HostStream()

    def _block_cap(self):
        """Records per (source, destination) block of the all-to-all path."""
        per_block = -(-self._cap // self.world)             # ceil
        return (per_block // 1024 + 1) * 1024

    MAX_REPEATS = 4

    def _check_inputs_alive(self, h):
        """An exchange is repeated from the tracker's ring buffers of its
        snapshot; they are recycled RING submits later.  (Every rank takes the
        same branch: the sizes that trigger a repeat are all-gathered.)"""
        # buffers of snapshot k are rewritten while snapshot k + RING is being
        # submitted, i.e. once tracker._step (snapshots submitted) > k + RING
        # (consecutive repeats: a finished exchange resets the count)
        self._repeats += 1
        if self._repeats > self.MAX_REPEATS:
            raise _lib.OrbitB200Error(
                "the event exchange overflowed its send buffers %d times; the "
                "event lists are not a uniform sample per rank (is the sharding "
                "by particle ID?)" % self._repeats)
        if getattr(h.res, 'persistent', False):
            # a batch is exchanged from its own staging set, which is reused by
            # the third batch after it
            if self._batches - h.res.batch_seq >= 3:
                raise _lib.OrbitB200Error(
                    "the exchange of a batch overflowed after its staging "
                    "buffers had been reused; call finish_batch() earlier")
            return
        if h.tracker._step - h.step0 > h.tracker.RING:
            raise _lib.OrbitB200Error(
                "the event exchange of a snapshot overflowed its send buffers "
                "after the snapshot's device buffers had been recycled; call "
                "finish_merge() no later than one snapshot after start_merge() "
                "or raise Comm.HEADROOM")

    def prepack(self, tracker, p):
        """Call right after ``tracker.submit*()`` returned the pending snapshot
        ``p``: enqueue the quantile and pack kernels of its events on the MAIN
        stream, directly behind its selection kernels -- i.e. before the next
        snapshot's persistent tracking kernel occupies the SMs (launched later,
        from ``start_merge`` on the exchange stream, the same two kernels took
        0.76 ms instead of 0.13 at 8 GPUs, profiles/r02_scaling.md).  The
        splitters are those of the last FINISHED exchange (complete on the host,
        so no stream has to wait for them).  ``start_merge`` then only runs the
        collectives and the merge.  No-op whenever the plain path applies (first
        exchanges, batches, no previous snapshot).  OPT-IN (OA_EXCHANGE_PREPACK=1):
        on 2 GPUs it is correct (oracle parity on every path) but slower (1.08 vs
        0.91 ms per step: the pack kernels lengthen the main stream's critical
        path and the contention moves to the merge kernel, 0.09 -> 0.41 ms); it
        has not been measured on 8 GPUs."""
        import os
        if os.environ.get('OA_EXCHANGE_PREPACK', '0') != '1':
            return
        prev = getattr(p, 'prev', None)
        if self.batch_size > 1 or self._cap is None or prev is None or \
                getattr(prev, 'gpos', None) is None or p.d_small is None or \
                getattr(self, '_splitters_ready', None) is None or \
                not tracker.events_on_device:
            return
        W = self.world
        n_seg = p.n_m
        cap = self._block_cap()
        st = tracker._stream()
        i64 = dict(dtype=torch.int64, device=self.device)
        n_prop, n_cnt = max(W - 1, 1), max(n_seg, 1)
        meta = torch.empty(2 + n_cnt + n_prop, **i64)
        counts, prop = meta[2:2 + n_cnt], meta[2 + n_cnt:]
        check(lib.oa_split_quantiles(ptr(prev.gpos), ptr(p.sel), ptr(p.d_small),
                                     n_seg, W, ptr(prop), st))
        prop_all = self._splitters_ready
        send = torch.empty(W * lib.oa_exchange_bytes(0, cap), dtype=torch.uint8,
                           device=self.device)
        bnd = torch.empty(W + 1, **i64)
        check(lib.oa_pack_split(
            ptr(prev.gpos), ptr(p.sel), ptr(p.d_ids), ptr(p.d_ang),
            ptr(p.d_small), n_seg, ptr(prop_all), W, cap, ptr(bnd), ptr(send),
            ptr(counts), st))
        packed = self._event()
        packed.record(tracker._main())
        tracker.launches += 2
        p.prepack = dict(meta=meta, send=send, bnd=bnd, prop_all=prop_all, cap=cap,
                         n_seg=n_seg, packed=packed)
        self.n_prepacked = getattr(self, 'n_prepacked', 0) + 1

    def _launch_split(self, tracker, res, cap, to_host, splitters=None):
        """``splitters``: the all-gathered quantile proposals to split by (a
        repeat uses those of the attempt it repeats).  Default: the proposals
        that travelled with the PREVIOUS exchange's sizes -- the key ranges lag
        one snapshot behind (the event density over the key space changes
        slowly; an imbalance beyond the headroom ends in a repeat), and an
        exchange costs two collectives: the all-to-all and one all-gather."""
        gen = res.prev_gen
        W = self.world
        n_seg = len(res.apsis_offsets) - 1
        h = _Exchange()
        h.res, h.cap, h.n_seg, h.to_host, h.tracker = res, cap, n_seg, to_host, tracker
        h.split = True
        h.step0 = getattr(res, 'step', tracker._step)
        pre = getattr(res, 'prepack', None) if splitters is None else None
        if pre is not None and (pre['cap'] != cap or pre['n_seg'] != n_seg):
            pre = None
        if pre is not None:
            res.prepack = None                 # a repeat packs again, the plain way
            self.stream.wait_event(pre['packed'])
        elif res.compacted is not None:
            self.stream.wait_event(res.compacted)
        prof = self._profile_events()
        with self._on_stream():
            st = C.c_void_p(self.stream.cuda_stream)
            self._mark(prof, 'start')
            i64 = dict(dtype=torch.int64, device=self.device)
            u8 = dict(dtype=torch.uint8, device=self.device)
            n_prop = max(W - 1, 1)
            # [slice size | largest block | per-halo counts of this rank | this
            # rank's quantile proposals]: one buffer, so that ONE all-gather
            # carries the sizes, the counts and the next exchange's splitters
            n_cnt = max(n_seg, 1)
            meta = pre['meta'] if pre is not None else \
                torch.empty(2 + n_cnt + n_prop, **i64)
            info, counts, prop = meta[:2], meta[2:2 + n_cnt], meta[2 + n_cnt:]
            if pre is None:
                check(lib.oa_split_quantiles(ptr(gen.gpos), ptr(res.d_sel),
                                             ptr(res.d_small), n_seg, W,
                                             ptr(prop), st))
            persistent = getattr(res, 'persistent', False)
            prop_all = splitters if splitters is not None else (
                None if persistent else self._splitters)
            if pre is not None:
                prop_all = pre['prop_all']
            if prop_all is None:
                # first exchange: nothing to lag behind.  A batch always splits by
                # its own quantiles: batches differ in length (the last one, the
                # one flushed at a checkpoint), and one more small all-gather per
                # K snapshots costs nothing
                prop_all = torch.empty(W * n_prop, **i64)
                dist.all_gather_into_tensor(prop_all, prop.contiguous())
            h.splitters = prop_all
            blk = lib.oa_exchange_bytes(0, cap)
            recv = torch.empty(W * blk, **u8)
            if pre is not None:
                send, bnd = pre['send'], pre['bnd']
            else:
                send = torch.empty(W * blk, **u8)
                bnd = torch.empty(W + 1, **i64)
                check(lib.oa_pack_split(
                    ptr(gen.gpos), ptr(res.d_sel), ptr(res.d_ids_buf),
                    ptr(res.d_ang_buf), ptr(res.d_small), n_seg, ptr(prop_all), W,
                    cap, ptr(bnd), ptr(send), ptr(counts), st))
            # The tracker's ring buffers (gpos, selection, event lists) have now
            # been read: the NEXT snapshot may be submitted once the pack kernel
            # is done.  (Round 1 recorded this event after the collectives, so
            # every submit waited for the all-to-all of the previous snapshot --
            # i.e. for the slowest rank -- before its tracking kernel could start.)
            # A batch reads staging buffers of its own: nothing to wait for.
            if not persistent and pre is None:
                packed = self._event()
                packed.record(self.stream)
                tracker.wait_before_submit = packed
            self._mark(prof, 'packed')
            dist.all_to_all_single(recv, send)
            self._mark(prof, 'all_to_all')
            h.ids = torch.empty(W * cap, **i64)
            h.ang = torch.empty(W * cap, dtype=torch.int16, device=self.device)
            check(lib.oa_merge_blocks(ptr(recv), W, cap, ptr(h.ids), ptr(h.ang),
                                      ptr(info), st))
            self._mark(prof, 'merged')
            meta_all = torch.empty(W * meta.numel(), **i64)
            dist.all_gather_into_tensor(meta_all, meta)
            self._mark(prof, 'all_gather')
            h.prof = prof
            # the proposals of all ranks, [W][W - 1], for the next exchange
            if not persistent:
                self._splitters = meta_all.view(W, -1)[:, 2 + n_cnt:].contiguous()
                h.next_splitters = self._splitters
            tracker.launches += 5
            # (meta_all is read by the copy stream below: it must stay allocated
            # until the handle is finished, or the caching allocator hands its
            # block to the next exchange while the copy is still queued)
            h.keep = (send, recv, prop_all, bnd, meta, meta_all)
        # small read-back into pinned buffers OWNED by the handle (torch's host
        # allocator caches them): any number of exchanges -- repeats included --
        # may be launched before this one is finished
        h.h_meta, h.ready = tracker.to_host_async(meta_all, stream=self.stream)
        return h

    BATCH_HEADROOM = 1.3  # batches split by their own quantiles: blocks are balanced

    def _round_cap(self, largest, headroom=None):
        headroom = self.HEADROOM if headroom is None else headroom
        return -(-int(headroom * max(largest, 1024)) // 4096) * 4096

    def _launch_merge(self, tracker, res, cap, to_host):
        gen = res.prev_gen
        n_seg = len(res.apsis_offsets) - 1
        h = _Exchange()
        h.res, h.cap, h.n_seg, h.to_host, h.tracker = res, cap, n_seg, to_host, tracker
        h.step0 = getattr(res, 'step', tracker._step)
        self.stream.wait_event(res.compacted)
        with self._on_stream():
            st = C.c_void_p(self.stream.cuda_stream)
            nbytes = lib.oa_exchange_bytes(n_seg, cap)
            send = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            check(lib.oa_pack_events(
                ptr(gen.gpos), ptr(res.d_sel), ptr(res.d_ids_buf),
                ptr(res.d_ang_buf), ptr(res.d_small), n_seg, cap, ptr(send),
                st))
            # the tracker's ring buffers have been read (see _launch_split)
            packed = self._event()
            packed.record(self.stream)
            tracker.wait_before_submit = packed
            recv = torch.empty(self.world * nbytes, dtype=torch.uint8,
                               device=self.device)
            dist.all_gather_into_tensor(recv, send)
            h.ids = torch.empty(self.world * cap, dtype=torch.int64,
                                device=self.device)
            h.ang = torch.empty(self.world * cap, dtype=torch.int16,
                                device=self.device)
            info = torch.empty(n_seg + 3 + self.world, dtype=torch.int64,
                               device=self.device)
            check(lib.oa_merge_gathered(ptr(recv), self.world, n_seg, cap,
                                        ptr(h.ids), ptr(h.ang), ptr(info), st))
            tracker.launches += 3
            h.keep = (send, recv, info)
        # small read-back (total, global offsets, sizes, overflow flag) into a
        # pinned buffer owned by the handle (see _launch_split)
        h.h_info, h.ready = tracker.to_host_async(info, stream=self.stream)
        return h

    def finish_merge(self, h):
        """Wait for an exchange and fill the global event lists into its
        ``StepResult``: ``n_events`` and ``apsis_offsets`` are global;
        ``d_ids`` / ``d_ang`` (and, after ``host_ready``, ``apsis_ids`` /
        ``apsis_angles``) hold the records ``host_slice`` of the global lists --
        everything on the all-gather path, this rank's key range on the
        all-to-all path."""
        h.ready.synchronize()
        if getattr(h, 'split', False):
            return self._finish_split(h)
        info = h.h_info.numpy()
        total = int(info[0])
        sizes = info[2 + h.n_seg:2 + h.n_seg + self.world]
        if int(info[2 + h.n_seg + self.world]):
            # some rank had more events than the send buffers hold: repeat this
            # snapshot's exchange with room for the largest list (all ranks see
            # the same sizes and take the same branch)
            self._check_inputs_alive(h)
            self._cap = max(self._cap, self._round_cap(int(sizes.max())))
            return self.finish_merge(self._launch_merge(
                h.tracker, h.res, self._cap, h.to_host))
        self._repeats = 0
        self._cap = max(self._cap, self._round_cap(int(sizes.max())))
        self._last_total = total
        res = h.res
        res.n_events = total
        res.apsis_offsets = info[1:2 + h.n_seg].copy()
        res.d_ids, res.d_ang = h.ids[:total], h.ang[:total]
        if h.to_host:
            # the merged lists live on every rank; hand the host either all of
            # them (single writer) or this rank's 1/world share (parallel write:
            # the per-rank copy stays constant under weak scaling).  The copy is
            # asynchronous: wait for res.host_ready before reading.
            lo, hi = 0, total
            if h.to_host == 'slice':
                lo = total * self.rank // self.world
                hi = total * (self.rank + 1) // self.world
            gen = res.prev_gen
            self._results += 1
            h_ids, h_ang, ready = h.tracker.to_host_async(
                res.d_ids[lo:hi], res.d_ang[lo:hi], stream=self.stream,
                names=('x_ids', 'x_ang'), reserve=max(hi - lo, self._cap),
                step=self._results, bulk=True)
            res.apsis_ids = h_ids.numpy().astype(gen.ids_dtype, copy=False)
            res.apsis_angles = h_ang.numpy().view(np.float16)
            res.host_slice, res.host_ready = (lo, hi), ready
        h.keep = None
        return res

    # -- optional profile of the exchange phases (OA_EXCHANGE_PROFILE=1) -------------
    def _profile_events(self):
        import os
        if self.device.type != 'cuda' or os.environ.get('OA_EXCHANGE_PROFILE') != '1':
            return None
        return []

    def _mark(self, prof, name):
        if prof is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(self.stream)
            prof.append((name, ev))

    def _account(self, h):
        prof = getattr(h, 'prof', None)
        if not prof:
            return
        seen = self.__dict__.get('_profiled', 0)
        self._profiled = seen + 1
        if seen < 3:                  # communicator set-up, first allocations
            return
        acc = self.__dict__.setdefault('phase_ms', {})
        for (_, a), (name, b) in zip(prof[:-1], prof[1:]):
            acc[name] = acc.get(name, 0.0) + a.elapsed_time(b)
        acc['exchanges'] = acc.get('exchanges', 0) + 1

    def _finish_split(self, h):
        W = self.world
        self._account(h)
        meta = h.h_meta.numpy().reshape(W, -1)
        info = meta[:, :2]
        sizes = info[:, 0]
        # info[:, 1] = largest (source, destination) block, in records, that
        # each rank was sent; every rank reads the same all-gathered numbers
        largest = int(info[:, 1].max())
        if largest > h.cap:
            # a block outgrew the send buffers: repeat with room for it (the
            # key ranges of the repeat are the same, so once is enough)
            self._check_inputs_alive(h)
            self._cap = max(self._cap, self._round_cap(largest * W))
            return self.finish_merge(self._launch_split(
                h.tracker, h.res, self._block_cap(), h.to_host,
                splitters=h.splitters))
        self._repeats = 0
        if getattr(h, 'next_splitters', None) is not None:
            self._splitters_ready = h.next_splitters   # complete: safe on any stream
        batch = getattr(h.res, 'persistent', False)
        self._cap = max(self._cap, self._round_cap(
            int(sizes.max()), self.BATCH_HEADROOM if batch else None))
        res = h.res
        total = int(sizes.sum())
        lo = int(sizes[:self.rank].sum())
        hi = lo + int(sizes[self.rank])
        res.n_events = total
        # global per-halo counts = sum over the ranks (the "all-reduce" of
        # SURVEY 8(e), done on the gathered rows)
        res.apsis_offsets = np.concatenate(
            ([0], np.cumsum(meta[:, 2:2 + h.n_seg].sum(axis=0)))).astype(np.int64)
        # (meta rows: [size | largest | counts | proposals for the next exchange])
        res.d_ids, res.d_ang = h.ids[:hi - lo], h.ang[:hi - lo]
        res.host_slice = (lo, hi)
        if h.to_host:
            gen = res.prev_gen
            self._results += 1
            # (a batch: K snapshots' events in buffers of their own, reserved once
            # by _reserve_batch; a single snapshot: room for the send capacity)
            batch = getattr(res, 'persistent', False)
            h_ids, h_ang, ready = h.tracker.to_host_async(
                res.d_ids, res.d_ang, stream=self.stream,
                names=('xb_ids', 'xb_ang') if batch else ('x_ids', 'x_ang'),
                reserve=int(1.25 * (hi - lo)) if batch else max(hi - lo, self._cap),
                step=self._results, bulk=True)
            res.apsis_ids = h_ids.numpy().astype(gen.ids_dtype, copy=False)
            res.apsis_angles = h_ang.numpy().view(np.float16)
            res.host_ready = ready
        h.keep = None
        return res

    def _reserve_batch(self, tracker, n_local, n_seg):
        """Called with the first snapshot staged on `tracker`: allocate what a
        full batch will need -- staging sets, send / receive / merge buffers (left
        in torch's caching allocator) and the pinned result buffers -- so that no
        exchange pays for cudaMalloc / page pinning later (pinning ~0.5 GB costs
        tenths of a second).  Every rank calls it at the same snapshot."""
        K, W = self.batch_size, self.world
        t = torch.tensor([n_local], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_snapshot = max(int(t.item()), 1024)
        self._cap = max(self._cap or 0, self._round_cap(K * per_snapshot,
                                                        self.BATCH_HEADROOM))
        n_ev = int(1.3 * K * per_snapshot)
        keep_open = self._open
        for slot in range(3):
            self._open = _Exchange()
            self._open.n_local = self._open.n_seg = 0
            self._staging(slot, n_ev, K * (n_seg + 1) + 1)
        self._open = keep_open
        if self.device.type == 'cuda':
            cap = self._block_cap()
            blk = lib.oa_exchange_bytes(0, cap)
            with self._on_stream():
                warm = [torch.empty(W * blk, dtype=torch.uint8, device=self.device)
                        for _ in range(2)]
                warm.append(torch.empty(W * cap, dtype=torch.int64, device=self.device))
                warm.append(torch.empty(W * cap, dtype=torch.int16, device=self.device))
            del warm
        if hasattr(tracker, '_hbuf'):         # (the CPU stand-in has no pinned rings)
            for k in range(tracker.HOST_RING):
                tracker._hbuf('xb_ids', n_ev, torch.int64, k)
                tracker._hbuf('xb_ang', n_ev, torch.int16, k)

    # -- events: several snapshots per exchange -----------------------------------
    # The tracking of later snapshots does not need the merged events, so the
    # exchange can lag: the local events of up to BATCH snapshots are staged in
    # HBM (one kernel per snapshot, no collective) and exchanged like ONE snapshot
    # whose halos are the concatenated halos of all of them -- the snapshot's slot
    # in the high bits of the order key keeps the snapshots apart.  The host work
    # and the collectives of an exchange are then paid once per batch.
    TAG_SHIFT = 58           # order keys stay below 2^58; at most 31 slots

    def stage_merge(self, tracker, res):
        """Append one snapshot's local events to the open batch (launching its
        exchange when ``batch_size`` snapshots are staged).  Returns the batch;
        ``finish_batch`` yields the snapshots' global results."""
        gen = res.prev_gen
        if gen is None or gen.gpos is None or res.d_small is None:
            raise _lib.OrbitB200Error(
                "stage_merge needs gpos= and OrbitTracker.events_on_device")
        if self.stream is None:
            self.stream = _HostStream() if self.device.type != 'cuda' \
                else torch.cuda.Stream(self.device)
        if getattr(self, 'stage_stream', None) is None:
            # staging has a stream of its own: on the exchange stream a snapshot's
            # stage kernel would queue behind the previous batch's all-to-all, and
            # the next submit (which waits for the stage kernel) with it
            self.stage_stream = _HostStream() if self.device.type != 'cuda' \
                else torch.cuda.Stream(self.device)
        b = self._open
        n_local, n_seg = int(res.n_events), len(res.apsis_offsets) - 1
        if getattr(self, '_reserved_for', None) is not tracker:
            self._reserved_for = tracker
            self._reserve_batch(tracker, n_local, n_seg)
        if b is None:
            b = self._open = _Exchange()
            b.items, b.n_local, b.n_seg, b.h = [], 0, 0, None
            b.slot = self._batches % 3          # staging set (2 batches alive)
            self._batches += 1
        st_set = self._staging(b.slot, b.n_local + n_local, b.n_seg + n_seg + 1)
        if res.compacted is not None:
            self.stage_stream.wait_event(res.compacted)
        # (the kernel takes its stream as an argument: no stream context)
        check(lib.oa_stage_events(
            ptr(gen.gpos), ptr(res.d_sel), ptr(res.d_ids_buf),
            ptr(res.d_ang_buf), ptr(res.d_small), n_seg, n_local,
            len(b.items) << self.TAG_SHIFT, b.n_local, ptr(st_set['keys']),
            ptr(st_set['ids']), ptr(st_set['ang']),
            ptr(st_set['small'][b.n_seg:]),
            C.c_void_p(self.stage_stream.cuda_stream)))
        tracker.launches += 1
        done = self._event()
        done.record(self.stage_stream)
        tracker.wait_before_submit = done       # the ring buffers were read
        b.staged = done                         # the exchange waits for the last one
        b.items.append((res, n_seg, b.n_seg, n_local))
        b.n_local += n_local
        b.n_seg += n_seg
        b.ids_dtype = gen.ids_dtype
        if len(b.items) >= self.batch_size:
            self.launch_batch(tracker)
        return b

    def _staging(self, slot, n_events, n_small):
        """Staging arrays of a batch (kept and grown geometrically).  Everything
        here -- allocation, the grow-copy, the identity selection -- is issued on
        the STAGING stream, which also runs the stage kernels; the exchange
        stream waits for the batch's last stage kernel before it reads them.  A
        replaced array stays referenced by the open batch until the batch is
        finished, so its block cannot be recycled under a pending copy."""
        st_set = self._stage_sets.setdefault(slot, {})
        b = self._open
        keys = st_set.get('keys')
        if keys is not None and keys.numel() >= n_events and \
                st_set['small'].numel() >= n_small and \
                st_set['iota'].numel() >= keys.numel():
            return st_set                   # steady state: nothing to allocate
        retired = b.__dict__.setdefault('retired', [])

        def grow(name, n, dtype, keep):
            old = st_set.get(name)
            if old is None or old.numel() < n:
                new = torch.empty(max(int(1.5 * n), 4096), dtype=dtype,
                                  device=self.device)
                if old is not None:
                    if keep:
                        new[:keep].copy_(old[:keep])
                    retired.append(old)
                st_set[name] = new

        with (torch.cuda.stream(self.stage_stream) if self.device.type == 'cuda'
              else self.stage_stream):
            grow('keys', n_events, torch.int64, b.n_local)
            grow('ids', n_events, torch.int64, b.n_local)
            grow('ang', n_events, torch.int16, b.n_local)
            grow('small', n_small, torch.int64, b.n_seg + 1)
            iota = st_set.get('iota')
            if iota is None or iota.numel() < st_set['keys'].numel():
                if iota is not None:
                    retired.append(iota)
                st_set['iota'] = torch.arange(st_set['keys'].numel(),
                                              dtype=torch.int64,
                                              device=self.device)
        return st_set

    def launch_batch(self, tracker):
        """Exchange the open batch (full or not); no-op without one."""
        b, self._open = self._open, None
        if b is None or not b.items:
            return None
        st_set = self._stage_sets[b.slot]
        self.stream.wait_event(b.staged)        # every snapshot of the batch is staged
        if self._cap is None:
            t = torch.tensor([b.n_local], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self._cap = self._round_cap(int(t.item()))
        # the batch as ONE snapshot: keys, identity selection, concatenated halos
        pseudo = _Exchange()
        pseudo.prev_gen = _Exchange()
        pseudo.prev_gen.gpos, pseudo.prev_gen.ids_dtype = st_set['keys'], b.ids_dtype
        pseudo.d_sel, pseudo.d_ids_buf = st_set['iota'], st_set['ids']
        pseudo.d_ang_buf, pseudo.d_small = st_set['ang'], st_set['small']
        pseudo.apsis_offsets = np.zeros(b.n_seg + 1, dtype=np.int64)
        pseudo.n_events, pseudo.compacted = b.n_local, None
        pseudo.persistent, pseudo.step = True, tracker._step
        pseudo.batch_seq = self._batches
        b.h = self._launch_split(tracker, pseudo, self._block_cap(), 'slice')
        self._launched.append(b)
        return b

    def finish_batch(self, b):
        """Global results of the snapshots of a launched batch, in order: like
        ``finish_merge(to_host='slice')`` per snapshot (``host_slice`` is this
        rank's part of that snapshot's list, possibly empty; wait for
        ``host_ready`` before reading ``apsis_ids`` / ``apsis_angles`` -- they
        stay valid until HOST_RING further batches have been finished)."""
        big = self.finish_merge(b.h)
        # (the copy of the batch's lists to the host is in flight: the results
        # carry its event, nothing waits here)
        lo, hi = big.host_slice
        off = big.apsis_offsets
        out = []
        for res, n_seg, seg_base, _ in b.items:
            g0, g1 = int(off[seg_base]), int(off[seg_base + n_seg])
            a, z = min(max(lo, g0), g1), min(max(hi, g0), g1)
            res.n_events = g1 - g0
            res.apsis_offsets = off[seg_base:seg_base + n_seg + 1] - g0
            res.host_slice = (a - g0, z - g0)
            res.d_ids, res.d_ang = big.d_ids[a - lo:z - lo], big.d_ang[a - lo:z - lo]
            res.apsis_ids = big.apsis_ids[a - lo:z - lo]
            res.apsis_angles = big.apsis_angles[a - lo:z - lo]
            res.host_ready = big.host_ready
            out.append(res)
        if b in self._launched:
            self._launched.remove(b)
        return out

    def merge_events(self, tracker, res, to_host=True):
        """``start_merge`` + ``finish_merge`` + wait for the host copy."""
        res = self.finish_merge(self.start_merge(tracker, res, to_host))
        if res.host_ready is not None:
            res.host_ready.synchronize()
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
