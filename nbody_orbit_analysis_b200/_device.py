"""Small device-side helpers shared by the consumers of the tracking path
(``progenitors.py``, ``postprocessing.py``): buffers, uploads and the segmented
radix sort built from the C-ABI primitives.  PyTorch only owns the memory."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.OrbitB200Error(
            "orbit-b200 needs a CUDA device (B200, sm_100a); there is no CPU "
            "fallback.")


class DeviceContext:
    """Buffers + stream of one CUDA device."""

    def __init__(self, device=None):
        require_cuda()
        self.device = torch.device(
            device if device is not None else
            'cuda:%d' % torch.cuda.current_device())
        self.launches = 0

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)

    def upload(self, arr, dtype=None):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        if arr.size == 0:
            return torch.empty(1, dtype=torch.from_numpy(arr).dtype,
                               device=self.device)[:0]
        return torch.from_numpy(arr.reshape(-1)).to(self.device)

    # -- sorting ---------------------------------------------------------------
    def sort_pairs(self, keys, vals, n, bits):
        """Stable LSD radix sort of (keys, vals) on the low ``bits`` bits."""
        if n == 0:
            return keys, vals
        ws_bytes = lib.oa_sort_workspace_bytes(n)
        ws = self.empty(ws_bytes, torch.uint8)
        k_out, v_out = self.empty(n, torch.int64), self.empty(n, torch.int64)
        bits = max(int(bits), 1)
        check(lib.oa_sort_pairs_u64(ptr(keys), ptr(vals), ptr(k_out),
                                    ptr(v_out), n, 0, bits, ptr(ws), ws_bytes,
                                    self.stream()))
        self.launches += 4 * (-(-bits // 8))
        return k_out, v_out

    def argsort_values(self, values, n):
        """``(sorted keys = value - min, order, min)`` of an int64 device
        array: np.argsort(kind='stable') / the sort inside np.unique."""
        st = self.stream()
        mm = self.empty(2, torch.int64)
        check(lib.oa_minmax_i64(ptr(values), n, ptr(mm), st))
        lo, hi = (int(v) for v in mm.cpu().tolist())
        seg = self.upload(np.array([0, n], dtype=np.int64))
        k_lo, k_hi, idx = (self.empty(n, torch.int64) for _ in range(3))
        check(lib.oa_segment_sort_keys(ptr(values), n, ptr(seg), 1, None,
                                       ptr(mm), ptr(k_lo), ptr(k_hi), ptr(idx),
                                       st))
        self.launches += 2
        keys, order = self.sort_pairs(k_lo, idx, n, (hi - lo).bit_length())
        return keys, order, lo

    def segment_order(self, values, n, seg_off):
        """Order that sorts ``values`` (int64, device) ascending inside every
        segment ``[seg_off[s], seg_off[s+1])``: two stable radix sorts (value,
        then segment)."""
        st = self.stream()
        seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        n_seg = len(seg_off) - 1
        mm = self.empty(2, torch.int64)
        check(lib.oa_minmax_i64(ptr(values), n, ptr(mm), st))
        lo, hi = (int(v) for v in mm.cpu().tolist())
        d_seg = self.upload(seg_off)
        k_lo, k_hi, idx = (self.empty(n, torch.int64) for _ in range(3))
        check(lib.oa_segment_sort_keys(ptr(values), n, ptr(d_seg), n_seg, None,
                                       ptr(mm), ptr(k_lo), ptr(k_hi), ptr(idx),
                                       st))
        _, order = self.sort_pairs(k_lo, idx, n, (hi - lo).bit_length())
        seg_sorted = self.gather_i64(k_hi, order, n)
        _, order = self.sort_pairs(seg_sorted, order, n,
                                   max(n_seg - 1, 1).bit_length())
        self.launches += 2
        return order

    def gather_i64(self, src, sel, n):
        out = self.empty(n, torch.int64)
        if n:
            check(lib.oa_gather_i64(ptr(src), ptr(sel), n, None, ptr(out),
                                    self.stream()))
            self.launches += 1
        return out

    def select(self, marks, n, op, value):
        """Ascending positions i < n with ``marks[i] (op) value``."""
        st = self.stream()
        ws_bytes = lib.oa_select_workspace_bytes(n)
        ws = self.empty(ws_bytes, torch.uint8)
        d_total = self.empty(1, torch.int64)
        check(lib.oa_select_count(ptr(marks), n, op, value, ptr(ws), ws_bytes,
                                  ptr(d_total), st))
        total = int(d_total.item())
        sel = self.empty(total, torch.int64)
        if total:
            check(lib.oa_select_gather(ptr(marks), n, op, value, ptr(ws),
                                       ptr(sel), st))
        self.launches += 3
        return sel, total
