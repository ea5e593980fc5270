"""numpy stand-ins for the kernels of the multi-GPU event exchange.

Test infrastructure only.  Each function restates, element by element, what the
kernel of the same name in ``csrc/oa_segment.cu`` computes (byte layout of the
send buffers included), on host tensors.  Two uses:

* ``tests/test_sharded_gloo.py`` plugs ``EmulLib`` / ``EmulTracker`` into
  ``sharded.Comm`` so that the HOST side of the asynchronous exchange (capacity
  agreement, overflow + repeat, read-back slots, slices per rank) runs under
  gloo with world_size > 1 in the CPU container -- the same Python code that
  drives NCCL on the GPU box;
* ``tests/test_gpu_zz_exchange.py`` checks the CUDA kernels against these
  functions on one GPU.
"""
import numpy as np
import torch

INT64_MAX = np.iinfo(np.int64).max


def layout(n_seg, cap):
    """Byte offsets of one send buffer (``exchange_layout`` in oa_segment.cu):
    [int64 size | int64 counts[n_seg] | pad16 | keys[cap] | ids[cap] |
    uint16 angles[cap] | pad16]."""
    o = 8 * (1 + int(n_seg))
    o = (o + 15) & ~15
    keys = o
    o += 8 * cap
    ids = o
    o += 8 * cap
    ang = o
    o += 2 * cap
    return keys, ids, ang, (o + 15) & ~15


def _np(t):
    return t.numpy() if isinstance(t, torch.Tensor) else t


def _i64(buf, off, n):
    return buf[off:off + 8 * n].view(np.int64)


def _u16(buf, off, n):
    return buf[off:off + 2 * n].view(np.uint16)


class EmulLib:
    """Drop-in for the ctypes ``lib`` object inside ``sharded`` (with
    ``sharded.ptr`` replaced by the identity and ``check`` by a no-op)."""

    @staticmethod
    def oa_exchange_bytes(n_seg, cap):
        return layout(n_seg, cap)[3]

    # -- all-gather path -----------------------------------------------------
    @staticmethod
    def oa_pack_events(gpos, sel, ids, ang, small, n_seg, cap, out, st=None):
        gpos, sel, ids, ang, small, out = map(_np, (gpos, sel, ids, ang, small, out))
        k_off, i_off, a_off, _ = layout(n_seg, cap)
        total = int(small[n_seg])
        m = min(total, cap)
        hdr = _i64(out, 0, 1 + n_seg)
        hdr[0] = total
        ends = np.append(small[1:n_seg], total) if n_seg else np.zeros(0, np.int64)
        hdr[1:] = ends - small[:n_seg]
        _i64(out, k_off, cap)[:m] = gpos[sel[:m]]
        _i64(out, i_off, cap)[:m] = ids[:m]
        _u16(out, a_off, cap)[:m] = ang[:m].view(np.uint16)
        return 0

    @staticmethod
    def oa_merge_gathered(recv, world, n_seg, cap, ids_out, ang_out, info, st=None):
        recv, ids_out, ang_out, info = map(_np, (recv, ids_out, ang_out, info))
        k_off, i_off, a_off, nb = layout(n_seg, cap)
        sizes = np.array([_i64(recv, r * nb, 1)[0] for r in range(world)])
        counts = sum(_i64(recv, r * nb + 8, n_seg) for r in range(world)) \
            if n_seg else np.zeros(0, np.int64)
        kept = np.minimum(sizes, cap)
        info[0] = kept.sum()
        info[1:2 + n_seg] = np.concatenate(([0], np.cumsum(counts)))
        info[2 + n_seg:2 + n_seg + world] = sizes
        info[2 + n_seg + world] = int((sizes > cap).any())
        keys = np.concatenate([_i64(recv, r * nb + k_off, cap)[:kept[r]]
                               for r in range(world)])
        ids = np.concatenate([_i64(recv, r * nb + i_off, cap)[:kept[r]]
                              for r in range(world)])
        ang = np.concatenate([_u16(recv, r * nb + a_off, cap)[:kept[r]]
                              for r in range(world)])
        order = np.argsort(keys, kind='stable')
        ids_out[:len(order)] = ids[order]
        ang_out.view(np.uint16)[:len(order)] = ang[order]
        return 0

    # -- all-to-all path -----------------------------------------------------
    @staticmethod
    def oa_split_quantiles(gpos, sel, small, n_seg, world, q_out, st=None):
        gpos, sel, small, q_out = map(_np, (gpos, sel, small, q_out))
        total = int(small[n_seg])
        for j in range(world - 1):
            q_out[j] = gpos[sel[total * (j + 1) // world]] if total > 0 \
                else INT64_MAX
        return 0

    @staticmethod
    def split_bounds(proposals, world, keys):
        """bnd[q] = first local event with key >= median proposal q-1."""
        bnd = np.zeros(world + 1, dtype=np.int64)
        bnd[world] = len(keys)
        prop = np.asarray(proposals).reshape(world, max(world - 1, 1))
        for q in range(1, world):
            split = np.sort(prop[:, q - 1])[(world - 1) // 2]
            bnd[q] = np.searchsorted(keys, split, side='left')
        return bnd

    @staticmethod
    def oa_pack_split(gpos, sel, ids, ang, small, n_seg, proposals, world, cap,
                      bnd_ws, out, counts, st=None):
        gpos, sel, ids, ang, small, out, counts, bnd_ws = map(
            _np, (gpos, sel, ids, ang, small, out, counts, bnd_ws))
        total = int(small[n_seg])
        keys = gpos[sel[:total]]
        bnd = EmulLib.split_bounds(_np(proposals), world, keys) if world > 1 \
            else np.array([0, total], dtype=np.int64)
        bnd_ws[:world + 1] = bnd
        ends = np.append(small[1:n_seg], total) if n_seg else np.zeros(0, np.int64)
        counts[:n_seg] = ends - small[:n_seg]
        k_off, i_off, a_off, nb = layout(0, cap)
        for q in range(world):
            lo, hi = int(bnd[q]), int(bnd[q + 1])
            _i64(out, q * nb, 1)[0] = hi - lo
            m = min(hi - lo, cap)
            _i64(out, q * nb + k_off, cap)[:m] = keys[lo:lo + m]
            _i64(out, q * nb + i_off, cap)[:m] = ids[lo:lo + m]
            _u16(out, q * nb + a_off, cap)[:m] = ang[lo:lo + m].view(np.uint16)
        return 0

    @staticmethod
    def oa_stage_events(gpos, sel, ids, ang, small, n_seg, n_local, tag, ev_base,
                        out_keys, out_ids, out_ang, out_small, st=None):
        gpos, sel, ids, ang, small, out_keys, out_ids, out_ang, out_small = map(
            _np, (gpos, sel, ids, ang, small, out_keys, out_ids, out_ang, out_small))
        out_small[:n_seg] = small[:n_seg] + ev_base
        out_small[n_seg] = ev_base + n_local
        out_keys[ev_base:ev_base + n_local] = gpos[sel[:n_local]] | tag
        out_ids[ev_base:ev_base + n_local] = ids[:n_local]
        out_ang.view(np.uint16)[ev_base:ev_base + n_local] = \
            ang[:n_local].view(np.uint16)
        return 0

    @staticmethod
    def oa_merge_blocks(recv, world, cap, ids_out, ang_out, info, st=None):
        recv, ids_out, ang_out, info = map(_np, (recv, ids_out, ang_out, info))
        k_off, i_off, a_off, nb = layout(0, cap)
        sizes = np.array([_i64(recv, r * nb, 1)[0] for r in range(world)])
        kept = np.minimum(sizes, cap)
        info[0] = kept.sum()
        info[1] = sizes.max()
        keys = np.concatenate([_i64(recv, r * nb + k_off, cap)[:kept[r]]
                               for r in range(world)])
        ids = np.concatenate([_i64(recv, r * nb + i_off, cap)[:kept[r]]
                              for r in range(world)])
        ang = np.concatenate([_u16(recv, r * nb + a_off, cap)[:kept[r]]
                              for r in range(world)])
        order = np.argsort(keys, kind='stable')
        ids_out[:len(order)] = ids[order]
        ang_out.view(np.uint16)[:len(order)] = ang[order]
        return 0


class _Done:
    def synchronize(self):
        pass


class EmulTracker:
    """The part of ``OrbitTracker`` the exchange touches, on the host.  Read-back
    buffers are REAL ring slots (name, step mod HOST_RING) that later copies
    overwrite, like the tracker's pinned rings: an exchange that reads a slot
    another exchange has reused sees the wrong numbers here too."""
    RING = 3
    HOST_RING = 4

    def __init__(self):
        self._step = 0
        self.launches = 0
        self.wait_before_submit = None
        self._ring = {}

    def to_host_async(self, *tensors, stream=None, names=None, reserve=0,
                      step=None, bulk=False):
        step = self._step if step is None else step
        if names is None:                  # buffers owned by the caller
            return [t.clone() for t in tensors] + [_Done()]
        out = []
        for t, name in zip(tensors, names):
            key = (name, step % self.HOST_RING)
            buf = self._ring.get(key)
            if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                buf = torch.empty(max(t.numel(), reserve, 1), dtype=t.dtype)
                self._ring[key] = buf
            buf[:t.numel()].copy_(t)
            out.append(buf[:t.numel()])
        return out + [_Done()]


class EmulGen:
    def __init__(self, gpos, ids_dtype=np.int64):
        self.gpos = gpos
        self.ids_dtype = np.dtype(ids_dtype)


class EmulResult:
    """Fields of ``StepResult`` read / written by ``Comm.start_merge`` and
    ``finish_merge``."""

    def __init__(self, step, gpos_prev, sel, ids, ang, seg_begin):
        total = len(sel)
        cap = max(total, 1) + 7                  # buffers longer than the list
        self.step = step
        self.prev_gen = EmulGen(torch.from_numpy(gpos_prev))
        self.n_events = total
        pad = np.full(cap - total, -7, dtype=np.int64)
        self.d_sel = torch.from_numpy(np.concatenate((sel, pad * 0)))
        self.d_ids_buf = torch.from_numpy(np.concatenate((ids, pad)))
        self.d_ang_buf = torch.from_numpy(np.concatenate(
            (ang.view(np.int16), pad.astype(np.int16))))
        offs = np.searchsorted(sel, seg_begin, side='left').astype(np.int64)
        self.d_small = torch.from_numpy(np.append(offs, total))
        self.apsis_offsets = np.append(offs, total)
        self.compacted = None
        self.d_ids = self.d_ang = None
        self.apsis_ids = self.apsis_angles = None
        self.host_ready = self.host_slice = None


def install(sharded):
    """Route ``sharded``'s kernel calls to the numpy stand-ins (CPU tests)."""
    sharded.lib = EmulLib()
    sharded.ptr = lambda t: t
    sharded.check = lambda rc: None
