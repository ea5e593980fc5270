"""Synthetic orbit-tracking workloads.

Every particle follows an analytic rosette about its host halo,

    r(t)   = a (1 - e cos(w t + phi))
    psi(t) = kappa w t + psi0          (in-plane angle, random plane)

so it passes a pericentre whenever ``w t + phi`` crosses a multiple of 2 pi and
an apocentre half a radial period later.  Halo centres drift linearly through a
periodic box (exercising the minimum-image wrap, reference ``utils.py:24-33``),
the catalogue bulk velocity is that drift, particles outside the region radius
are omitted from the block (natural entered/departed churn) and the block order
is re-shuffled every snapshot so that ID matching is non-trivial.  All random
elements come from a counter-based splitmix64 hash of ``(seed, particle)``.

The object exposes exactly the callback protocol of the reference entry points
(``regions`` and ``load_snapshot_data``; reference ``track_orbits.py:27-61``).
"""
import numpy as np

SEED = 20261018
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over='ignore'):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _uniform(seed, stream, idx):
    """U[0,1) from hash(seed, stream, idx) with 53 random bits."""
    with np.errstate(over='ignore'):
        k = splitmix64(np.uint64(seed) * np.uint64(0x100000001B3)
                       + np.uint64(stream))
        z = splitmix64(np.asarray(idx, dtype=np.uint64) ^ k)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def halo_sizes(n_particles, n_halos, largest_frac=0.05, slope=1.9):
    """Power-law halo sizes summing to n_particles (largest ~ largest_frac)."""
    if n_halos == 1:
        return np.array([n_particles], dtype=np.int64)
    rank = np.arange(1, n_halos + 1, dtype=np.float64)
    w = rank ** (-1.0 / (slope - 1.0))
    w = w / w.sum()
    if w[0] > largest_frac:
        # flatten the head so that the largest halo holds ~largest_frac
        w = np.minimum(w, largest_frac)
        w = w / w.sum()
    if n_particles < n_halos:
        raise ValueError("too many halos for this particle count")
    # every halo gets one particle, the rest follows the power law
    sizes = np.floor(w * (n_particles - n_halos)).astype(np.int64) + 1
    sizes[0] += n_particles - sizes.sum()
    return sizes


class SynthSim:
    """A reproducible synthetic simulation with the reference callback API."""

    def __init__(self, n_particles, n_halos, n_snap, seed=SEED, box=100.0,
                 dtype=np.float32, catalogue_dtype=np.float64,
                 first_snapshot=10, nfw=False, hubble=False,
                 catalogue_bulk=True, mass_array=False, late_halos=0.0,
                 periodic=True, id_stride=1, id_offset=0, region_radius=None,
                 box_vector=False):
        self.N = int(n_particles)
        self.n_halos = int(n_halos)
        self.n_snap = int(n_snap)
        self.seed = int(seed)
        self.box = float(box)
        self.dtype = np.dtype(dtype)
        self.cat_dtype = np.dtype(catalogue_dtype)
        self.hubble = hubble
        self.catalogue_bulk = catalogue_bulk
        self.mass_array = mass_array
        self.periodic = periodic
        self.box_vector = box_vector
        self.snapshot_numbers = np.arange(
            first_snapshot, first_snapshot + n_snap)
        self.id_stride, self.id_offset = int(id_stride), int(id_offset)

        self.sizes = halo_sizes(self.N, self.n_halos)
        self.starts = np.concatenate(([0], np.cumsum(self.sizes)))
        hidx = np.arange(self.n_halos)
        # region radius: scaled so that regions are small next to the box
        if region_radius is None:
            region_radius = 0.02 * self.box * (
                self.sizes / self.sizes.max()) ** (1.0 / 3.0) + 0.002 * self.box
        self.radius = np.broadcast_to(
            np.asarray(region_radius, dtype=np.float64),
            (self.n_halos,)).copy()
        self.c0 = np.stack(
            [_uniform(seed, 101 + k, hidx) for k in range(3)], axis=1) \
            * self.box
        self.vh = (np.stack(
            [_uniform(seed, 111 + k, hidx) for k in range(3)], axis=1)
            - 0.5) * 0.02 * self.box
        # halo ids change from snapshot to snapshot like a real merger tree
        self.birth = np.zeros(self.n_halos, dtype=np.int64)
        if late_halos > 0 and self.n_halos > 1:
            late = _uniform(seed, 121, hidx) < late_halos
            late[0] = False
            self.birth[late] = 1 + (
                _uniform(seed, 122, hidx)[late] * max(n_snap - 2, 1)
            ).astype(np.int64)
        t = np.arange(n_snap)[:, None]
        self.main_branches = np.where(
            t >= self.birth[None, :],
            (self.snapshot_numbers[:, None] * 1000003 + hidx[None, :]
             ) % 2000000011, -1).astype(np.int64)
        self.nfw = nfw
        self._halo_of = None

    # -- per-particle orbit elements ---------------------------------------
    def _elements(self, u, h):
        s = self.seed
        R = self.radius[h]
        if self.nfw:
            # invert the NFW enclosed-mass profile (c = 10) on a table
            c = 10.0
            x = np.linspace(0.0, 1.2, 4097)
            m = np.log1p(c * x) - c * x / (1.0 + c * x)
            a = np.interp(_uniform(s, 1, u) * m[-1], m, x) * R
            a = np.maximum(a, 0.01 * R)
        else:
            a = (0.05 + 1.15 * _uniform(s, 1, u)) * R
        e = 0.1 + 0.7 * _uniform(s, 2, u)
        w = 0.1 + 0.9 * _uniform(s, 3, u)
        phi = 2 * np.pi * _uniform(s, 4, u)
        psi0 = 2 * np.pi * _uniform(s, 5, u)
        kappa = 0.55 + 0.4 * _uniform(s, 6, u)
        # random orbital plane: orthonormal pair (e1, e2)
        cz = 2 * _uniform(s, 7, u) - 1
        az = 2 * np.pi * _uniform(s, 8, u)
        sz = np.sqrt(np.maximum(1 - cz * cz, 0))
        n = np.stack([sz * np.cos(az), sz * np.sin(az), cz], axis=1)
        ref = np.where(np.abs(n[:, 2:3]) < 0.9,
                       np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
        e1 = np.cross(n, ref)
        e1 /= np.linalg.norm(e1, axis=1)[:, None]
        e2 = np.cross(n, e1)
        return a, e, w, phi, psi0, kappa, e1, e2

    def halo_of(self):
        if self._halo_of is None:
            self._halo_of = np.repeat(
                np.arange(self.n_halos, dtype=np.int64), self.sizes)
        return self._halo_of

    def halo_centre(self, t):
        c = self.c0 + self.vh * float(t)
        if self.periodic:
            c = np.mod(c, self.box)
        return c

    def cosmology(self, t):
        if self.hubble:
            z = 0.5 * (1.0 - t / max(self.n_snap - 1, 1))
            return dict(redshift=float(z), H0=0.07, Omega_m=0.3, Omega_L=0.7)
        return dict(redshift=0.0, H0=0.0, Omega_m=0.3, Omega_L=0.7)

    # -- the reference callback protocol -------------------------------------
    def regions(self, snapshot_number, halo_ids):
        t = int(np.searchsorted(self.snapshot_numbers, snapshot_number))
        cols = self._cols(t, halo_ids)
        pos = self.halo_centre(t)[cols].astype(self.cat_dtype)
        rad = self.radius[cols].astype(self.cat_dtype)
        if self.catalogue_bulk:
            return pos, rad, self.vh[cols].astype(self.cat_dtype)
        return pos, rad, None

    def regions_onthefly(self, snapshot_number, halo_ids):
        return self.regions(snapshot_number, halo_ids)[:2]

    def _cols(self, t, halo_ids):
        halo_ids = np.atleast_1d(np.asarray(halo_ids))
        row = self.main_branches[t]
        order = np.argsort(row)
        return order[np.searchsorted(row[order], halo_ids)]

    def load_snapshot_data(self, snapshot_number, region_positions,
                           region_radii, cols=None):
        t = int(np.searchsorted(self.snapshot_numbers, snapshot_number))
        if cols is None:
            # the loader contract only hands us positions: recover the halo
            # columns from the (exactly reproduced) centres
            cen = self.halo_centre(t).astype(self.cat_dtype)
            cols = _match_rows(cen, np.atleast_2d(region_positions))
        cols = np.asarray(cols, dtype=np.int64)
        lens = self.sizes[cols]
        u = np.concatenate(
            [np.arange(self.starts[c], self.starts[c + 1]) for c in cols]
        ) if len(cols) else np.zeros(0, dtype=np.int64)
        slot = np.repeat(np.arange(len(cols), dtype=np.int64), lens)
        h = cols[slot]
        a, e, w, phi, psi0, kappa, e1, e2 = self._elements(u, h)
        ph = w * t + phi
        r = a * (1 - e * np.cos(ph))
        rdot = a * e * w * np.sin(ph)
        psi = kappa * w * t + psi0
        cp, sp = np.cos(psi)[:, None], np.sin(psi)[:, None]
        er = cp * e1 + sp * e2
        et = -sp * e1 + cp * e2
        xrel = r[:, None] * er
        vrel = rdot[:, None] * er + (r * kappa * w)[:, None] * et
        keep = r <= self.radius[h]
        # shuffle inside every block, keyed on (seed, snapshot, particle)
        key = (slot.astype(np.uint64) << np.uint64(40)) | (
            splitmix64(u.astype(np.uint64) ^ splitmix64(
                np.uint64(self.seed * 7919 + 104729 * (t + 1))))
            >> np.uint64(24))
        sel = np.flatnonzero(keep)
        sel = sel[np.argsort(key[sel], kind='stable')]
        x = self.halo_centre(t)[h[sel]] + xrel[sel]
        if self.periodic:
            x = np.mod(x, self.box)
        v = self.vh[h[sel]] + vrel[sel]
        counts = np.bincount(slot[sel], minlength=len(cols))
        snap = {
            'ids': (u[sel] * self.id_stride + self.id_offset).astype(np.int64),
            'coordinates': np.ascontiguousarray(x.astype(self.dtype)),
            'velocities': np.ascontiguousarray(v.astype(self.dtype)),
            'region_offsets': np.concatenate(
                ([0], np.cumsum(counts)[:-1])).astype(np.int64),
        }
        if self.mass_array:
            snap['masses'] = (
                0.5 + _uniform(self.seed, 9, u[sel])).astype(self.dtype)
        else:
            snap['masses'] = 1.0
        if self.periodic:
            snap['box_size'] = (self.box * np.ones(3)
                                if self.box_vector else self.box)
        snap.update(self.cosmology(t))
        return snap


def _match_rows(table, rows):
    """Row indices of `rows` inside `table` (exact match)."""
    tv = np.ascontiguousarray(table).view(
        [('', table.dtype)] * table.shape[1]).ravel()
    rv = np.ascontiguousarray(rows.astype(table.dtype)).view(
        [('', table.dtype)] * table.shape[1]).ravel()
    order = np.argsort(tv)
    pos = np.searchsorted(tv[order], rv)
    pos = np.clip(pos, 0, len(order) - 1)
    out = order[pos]
    if not np.array_equal(tv[out], rv):
        raise ValueError("region positions do not belong to this simulation")
    return out


class DeviceSynth:
    """The same kind of workload generated directly in HBM (throughput runs).

    Uses the generator kernels of ``csrc/oa_synth.cu`` and the library's radix
    sort for the per-snapshot block shuffle.  Sharding: rank ``r`` of ``w``
    owns particle IDs ``u * w + r`` (``id mod w == r``); every
    rank holds a sub-block of every halo.
    """

    def __init__(self, n_particles, n_halos, seed=SEED, box=100.0,
                 dtype=np.float32, catalogue_dtype=np.float32, rank=0,
                 world=1, device=None):
        import ctypes as C
        import torch
        from . import _lib
        self._C, self._torch, self._lib = C, torch, _lib
        self.host = SynthSim(n_particles, n_halos, 1, seed=seed, box=box,
                             dtype=dtype, catalogue_dtype=catalogue_dtype)
        self.dtype = np.dtype(dtype)
        self.cat_dtype = np.dtype(catalogue_dtype)
        self.device = torch.device(device if device is not None else
                                   'cuda:%d' % torch.cuda.current_device())
        self.rank, self.world = int(rank), int(world)
        h = self.host
        self.n_halos, self.M = h.n_halos, h.N

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.d_start = up(h.starts.astype(np.int64))
        self.d_radius = up(h.radius.astype(np.float64))
        self.d_c0 = up(h.c0.astype(np.float64).reshape(-1))
        self.d_vh = up(h.vh.astype(np.float64).reshape(-1))
        self.sort_bits = 40 + int(self.n_halos).bit_length()
        self.ws_bytes = _lib.lib.oa_sort_workspace_bytes(self.M)

    def _params(self, t):
        p = self._lib.SynthParams()
        p.seed = self.host.seed
        p.n_universe = self.M
        p.id_stride, p.id_offset = self.world, self.rank
        p.halo_start = self.d_start.data_ptr()
        p.halo_radius = self.d_radius.data_ptr()
        p.halo_c0 = self.d_c0.data_ptr()
        p.halo_vh = self.d_vh.data_ptr()
        p.box = self.host.box
        p.t = float(t)
        p.n_halos = self.n_halos
        p.periodic = 1
        return p

    def regions(self, t):
        """(centres, radii, bulk velocities) of all halos at snapshot t."""
        h = self.host
        return (h.halo_centre(t).astype(self.cat_dtype),
                h.radius.astype(self.cat_dtype), h.vh.astype(self.cat_dtype))

    def snapshot(self, t):
        """Generate snapshot ``t`` in HBM.  Returns ``(dev, n, offsets)`` with
        ``dev`` = dict of flat device tensors (pos, vel, ids)."""
        torch, lib, C = self._torch, self._lib.lib, self._C
        check, ptr = self._lib.check, self._lib.ptr
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        M = self.M

        def empty(n, dt):
            return torch.empty(int(n), dtype=dt, device=self.device)
        keys, vals = empty(M, torch.int64), empty(M, torch.int64)
        keys2, vals2 = empty(M, torch.int64), empty(M, torch.int64)
        counts = empty(self.n_halos, torch.int64)
        ws = empty(self.ws_bytes, torch.uint8)
        p = self._params(t)
        check(lib.oa_synth_keys(C.byref(p), ptr(keys), ptr(vals), ptr(counts),
                                st))
        check(lib.oa_sort_pairs_u64(ptr(keys), ptr(vals), ptr(keys2),
                                    ptr(vals2), M, 0, self.sort_bits, ptr(ws),
                                    self.ws_bytes, st))
        lens = counts.cpu().numpy()
        offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        n = int(offsets[-1])
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        # global block-order key of every particle: (halo << 40 | shuffle hash);
        # the order all ranks' particles would have in one unsharded block
        dev = {'pos': empty(3 * n, tdt), 'vel': empty(3 * n, tdt),
               'ids': empty(n, torch.int64), 'mass': None,
               'gpos': keys2[:n].clone() if self.world > 1 else None}
        check(lib.oa_synth_fill(C.byref(p), ptr(vals2), n,
                                int(self.dtype == np.float64), ptr(dev['pos']),
                                ptr(dev['vel']), ptr(dev['ids']), st))
        return dev, n, offsets
