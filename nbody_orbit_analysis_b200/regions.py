"""Region extraction on the GPU -- the loader step in front of ``track_orbits``
(SURVEY.md section 8(f)-1).

The reference's example loader (``example_script.py:36-67``) selects, for every
halo, the particles within the region radius of the halo centre by testing ALL
particles -- O(N x n_halo) numpy work that dominates the wall clock once the
tracking itself is fast.  ``extract_regions`` returns exactly what that loader
returns (``ids``, ``coordinates``, ``velocities``, ``masses``,
``region_offsets``: regions in catalogue order, particle indices ascending
inside a region) from one pass over the particles: a uniform grid holds, per
cell, the regions whose sphere touches it; a particle is tested against the
regions of its cell only, with the reference's arithmetic (``coordinates -
position`` in numpy's promoted dtype, ``utils.recenter_coordinates``,
``utils.vector_norm``, ``r < radius``), so the selection is bit-identical.

    snapshot = extract_regions(coordinates, positions, radii, box_size=L,
                               ids=ids, velocities=vel, masses=m)

``to_host=False`` leaves the gathered arrays in HBM (flat torch tensors ``pos``,
``vel``, ``ids``) in the form ``OrbitTracker.submit_device`` takes.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._device import DeviceContext
from ._lib import lib, check, ptr

MAX_CELLS_PER_AXIS = 256


def _grid(centres, radii, box, n_axis=None):
    """Grid geometry: (lo[3], inv_cell[3], dim[3]) on the host."""
    if box is not None:
        lo = np.zeros(3)
        span = box.astype(np.float64)
    else:
        lo = (centres - radii[:, None]).min(axis=0)
        span = np.maximum((centres + radii[:, None]).max(axis=0) - lo, 1e-30)
        lo = lo - 1e-6 * span
        span = span * (1 + 2e-6)
    if n_axis is None:
        r_med = float(np.median(radii)) if len(radii) else float(span.max())
        n_axis = np.clip((span / max(2.0 * r_med, 1e-30)).astype(np.int64), 1,
                         MAX_CELLS_PER_AXIS)
    dim = np.broadcast_to(np.asarray(n_axis, dtype=np.int64), (3,)).astype(np.int32)
    return lo, dim / span, dim


def _cell_lists(centres, radii, lo, inv_cell, dim, periodic):
    """CSR lists ``cell -> regions whose sphere touches the cell`` (host, the
    catalogue is small next to the particle data)."""
    n = len(radii)
    pad = radii * 1e-6 + 1e-12          # rounding of the cell index on either side
    k0 = np.floor((centres - (radii + pad)[:, None] - lo) * inv_cell).astype(np.int64)
    k1 = np.floor((centres + (radii + pad)[:, None] - lo) * inv_cell).astype(np.int64)
    if periodic:
        cnt = np.minimum(k1 - k0 + 1, dim)
        k0 = np.where(cnt >= dim, 0, k0)
    else:
        k0 = np.clip(k0, 0, dim - 1)
        k1 = np.clip(k1, 0, dim - 1)
        cnt = k1 - k0 + 1
    tot = cnt.prod(axis=1)
    start = np.concatenate(([0], np.cumsum(tot)))
    reg = np.repeat(np.arange(n), tot)
    local = np.arange(start[-1]) - start[reg]
    cx, cy = cnt[reg, 0], cnt[reg, 1]
    ix = local % cx
    iy = (local // cx) % cy
    iz = local // (cx * cy)
    gx = (k0[reg, 0] + ix) % dim[0]
    gy = (k0[reg, 1] + iy) % dim[1]
    gz = (k0[reg, 2] + iz) % dim[2]
    cell = (gz * dim[1] + gy) * dim[0] + gx
    order = np.argsort(cell, kind='stable')
    n_cells = int(dim.prod())
    cell_start = np.zeros(n_cells + 1, dtype=np.int64)
    np.cumsum(np.bincount(cell, minlength=n_cells), out=cell_start[1:])
    return cell_start.astype(np.int32), reg[order].astype(np.int32)


def extract_regions(coordinates, region_positions, region_radii, box_size=None,
                    ids=None, velocities=None, masses=None, device=None,
                    to_host=True, cells_per_axis=None):
    """Particles within ``region_radii[j]`` of ``region_positions[j]`` for every
    region j (``example_script.py:50-58``).  ``coordinates`` / ``velocities`` /
    ``ids`` / ``masses`` may be numpy arrays or device tensors (N,3)/(N,).
    Returns the loader's dict; ``region_inds`` (the particle indices) is added.
    """
    ctx = DeviceContext(device)
    st = ctx.stream()

    def dev(a, dtype=None):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(ctx.device).contiguous().reshape(-1)
        return ctx.upload(np.ascontiguousarray(a, dtype=dtype))
    pos_np_dtype = coordinates.dtype if not isinstance(coordinates, torch.Tensor) \
        else np.dtype(str(coordinates.dtype).split('.')[-1])
    if np.dtype(pos_np_dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
        coordinates = np.asarray(coordinates, dtype=np.float64)
        pos_np_dtype = np.dtype(np.float64)
    pos_np_dtype = np.dtype(pos_np_dtype)
    d_pos = dev(coordinates)
    n = d_pos.numel() // 3
    centres = np.atleast_2d(np.asarray(region_positions))
    radii = np.atleast_1d(np.asarray(region_radii))
    n_h = len(radii)
    # numpy's promotion of `coordinates - position` (array - array)
    cdt = centres.dtype if centres.dtype.kind == 'f' else np.dtype(np.float64)
    frame = np.result_type(pos_np_dtype, cdt)
    c64 = np.ascontiguousarray(centres, dtype=np.float64)
    c32 = np.ascontiguousarray(centres, dtype=np.float32)
    r64 = np.ascontiguousarray(radii, dtype=np.float64)
    periodic = box_size is not None
    box = np.broadcast_to(np.asarray(box_size, dtype=np.float64), (3,)).copy() \
        if periodic else None
    lo, inv_cell, dim = _grid(c64, r64, box, cells_per_axis)
    cell_start, cell_regions = _cell_lists(c64, r64, lo, inv_cell, dim, periodic)

    d_c64, d_c32, d_r = ctx.upload(c64), ctx.upload(c32), ctx.upload(r64)
    d_cs, d_cr = ctx.upload(cell_start), ctx.upload(
        cell_regions if len(cell_regions) else np.zeros(1, dtype=np.int32))
    counter = ctx.empty(1, torch.int64)
    vp = C.c_void_p
    geo = (np.ascontiguousarray(lo, dtype=np.float64),
           np.ascontiguousarray(inv_cell, dtype=np.float64),
           np.ascontiguousarray(dim, dtype=np.int32))

    def pairs(keys, cap):
        check(lib.oa_region_pairs(
            ptr(d_pos), _lib.dtype_code(pos_np_dtype), n, ptr(d_c64), ptr(d_c32),
            ptr(d_r), _lib.dtype_code(frame), ptr(d_cs), ptr(d_cr),
            geo[0].ctypes.data_as(vp), geo[1].ctypes.data_as(vp),
            geo[2].ctypes.data_as(vp),
            box.ctypes.data_as(vp) if periodic else None, int(periodic),
            ptr(keys), cap, ptr(counter), st))
        ctx.launches += 1
        return int(counter.item())
    m = pairs(None, 0)                               # count, then emit
    keys = ctx.empty(m, torch.int64)
    if m:
        got = pairs(keys, m)
        assert got == m
        bits = 32 + max(n_h - 1, 1).bit_length()
        keys, _ = ctx.sort_pairs(keys, ctx.empty(m, torch.int64), m, bits)
    d_seg = ctx.upload(np.arange(n_h, dtype=np.int64) << 32)
    d_off = ctx.empty(n_h, torch.int64)
    check(lib.oa_segment_offsets(ptr(keys), m, None, ptr(d_seg), n_h, ptr(d_off), st))
    ctx.launches += 1
    offsets = d_off[:n_h].cpu().numpy()

    def gather(src, elem_bytes, rows3, dtype):
        if src is None:
            return None
        out = ctx.empty((3 if rows3 else 1) * m, dtype)[:(3 if rows3 else 1) * m]
        check(lib.oa_gather_by_key(ptr(src), elem_bytes, int(rows3), ptr(keys), m,
                                   ptr(out), st))
        ctx.launches += 1
        return out
    fdt = torch.float64 if pos_np_dtype == np.float64 else torch.float32
    out_pos = gather(d_pos, pos_np_dtype.itemsize, True, fdt)
    d_vel = dev(velocities)
    out_vel = gather(d_vel, d_vel.element_size(), True, d_vel.dtype) \
        if d_vel is not None else None
    d_ids = dev(ids)
    out_ids = gather(d_ids, d_ids.element_size(), False, d_ids.dtype) \
        if d_ids is not None else None
    mass_arr = masses is not None and not np.isscalar(masses)
    d_m = dev(masses) if mass_arr else None
    out_m = gather(d_m, d_m.element_size(), False, d_m.dtype) if mass_arr else masses
    inds = keys[:m] & 0xFFFFFFFF
    snap = {'region_offsets': offsets}
    if periodic:
        snap['box_size'] = box_size
    if not to_host:
        snap.update(pos=out_pos, vel=out_vel, ids=out_ids, mass=out_m if mass_arr
                    else None, region_inds=inds, n=m)
        return snap
    snap['coordinates'] = out_pos.cpu().numpy().reshape(-1, 3)
    if out_vel is not None:
        snap['velocities'] = out_vel.cpu().numpy().reshape(-1, 3)
    if out_ids is not None:
        snap['ids'] = out_ids.cpu().numpy()
    if masses is not None:
        snap['masses'] = out_m.cpu().numpy() if mass_arr else masses
    snap['region_inds'] = inds.cpu().numpy()
    return snap
