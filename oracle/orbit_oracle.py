"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the orbit-tracking hot path of ``orbitanalysis`` v0.1
(the reference mounted at ``/root/reference``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module, and only as the checker / the CPU
arm.  The product package (``nbody_orbit_analysis_b200``) never imports it and
fails loudly when its CUDA library is missing.

Parity status: **pinned** against the unmodified reference executed in the
build container -- ``tests/golden/make_golden.py`` runs the real
``/root/reference/orbitanalysis`` (h5py/pathos stand-ins only, see
``oracle/reference_harness.py``) on seeded inputs and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` requires this restatement to
reproduce them (bit-exact integers, and bit-exact floats on the generating
host).  The reference itself ships no tests or golden vectors (SURVEY.md
section 4), so that is the only pin that exists.

The arithmetic lives in numpy (unpinned dependency of the reference,
``requirements.txt:1``); canonical version here is numpy 2.3.x with NEP-50
promotion (hard part 4).  Floating-point expressions are
kept in the reference's evaluation order so that rounding is identical.

Each function cites the reference lines it restates.
"""
import numpy as np

from nbody_orbit_analysis_b200 import storage as _default_storage

MODES = ('pericentric', 'apocentric')


# ---------------------------------------------------------------------------
# primitives (reference utils.py)
# ---------------------------------------------------------------------------

def hubble_parameter(z, H0, Omega_m, Omega_L, Omega_k=0):
    """H(z); reference ``utils.py:36-39``."""
    return H0 * np.sqrt(
        Omega_m * (1 + z)**3 + Omega_k * (1 + z)**2 + Omega_L)


def minimum_image(delta, box_size):
    """Single +-L wrap with strict inequalities, in place, per axis.

    Reference ``utils.py:24-33``: ``x[x > L/2] -= L`` then ``x[x < -L/2] += L``
    (the second test sees the already shifted values).  A scalar box is
    broadcast to three axes as ``box * np.ones(3)`` (float64).
    """
    if isinstance(box_size, (float, np.floating, int, np.integer)):
        box_size = box_size * np.ones(3)
    for axis, side in enumerate(box_size):
        col = delta[:, axis]
        hi = np.flatnonzero(col > side / 2)
        delta[hi, axis] -= side
        lo = np.flatnonzero(delta[:, axis] < -side / 2)
        delta[lo, axis] += side
    return delta


def ordered_match(a, b):
    """Positions in ``a`` of the elements of ``b``, in ``b``'s order.

    Reference ``utils.py:4-11`` (``myin1d``), valid for unique ``a`` and
    ``b`` with ``b`` a subset of ``a`` -- the only way the hot path calls it.
    """
    a = np.asarray(a)
    b = np.asarray(b)
    order = np.argsort(a, kind='stable')
    where = np.searchsorted(a[order], b)
    return order[where].astype(np.int64, copy=False) if len(a) else \
        np.zeros(0, dtype=np.int64)


# ---------------------------------------------------------------------------
# per-region kernels of track_orbits.py
# ---------------------------------------------------------------------------

def region_frame(snapshot, region_slice, region_position, region_bulk_vel, H):
    """Halo-frame unit vectors and radial velocities of one region block.

    Reference ``track_orbits.py:247-290``.  Returns ``(rhat, v_r, bulk_vel)``.
    """
    lo, hi = int(region_slice[0]), int(region_slice[1])
    delta = snapshot['coordinates'][lo:hi] - region_position
    if 'box_size' in snapshot:
        delta = minimum_image(delta, snapshot['box_size'])

    vel = snapshot['velocities'][lo:hi]
    if region_bulk_vel is not None:
        bulk_vel = region_bulk_vel
    elif isinstance(snapshot['masses'], np.ndarray):
        m = snapshot['masses'][lo:hi]
        bulk_vel = np.sum(m[:, np.newaxis] * vel, axis=0) / np.sum(m)
    else:
        bulk_vel = np.mean(vel, axis=0)
    region_vels = vel - bulk_vel + H * delta / (1 + snapshot['redshift'])

    rads = np.sqrt(np.einsum('...i,...i', delta, delta))
    rhats = delta / rads[:, np.newaxis]
    radial_vels = np.einsum('...i,...i', region_vels, rhats)
    return rhats, radial_vels, bulk_vel


def compare_radial_velocities(ids, ids_prev, radial_vels, radial_vels_prev,
                              rhat, rhat_prev, mode):
    """Match a block against the same halo's previous block and flag apsides.

    Reference ``track_orbits.py:293-327``.  Same output dict.
    """
    survives = np.isin(ids_prev, ids)
    inds_departed = np.flatnonzero(~survives)
    ids_prev_ = ids_prev[survives]
    vr_prev_ = radial_vels_prev[survives]
    rhat_prev_ = rhat_prev[survives]

    inds_match = ordered_match(ids, ids_prev_)
    vr_match = radial_vels[inds_match]
    rhat_match = rhat[inds_match]

    if mode == 'pericentric':
        flip = (vr_prev_ < 0) & (vr_match > 0)
    elif mode == 'apocentric':
        flip = (vr_prev_ > 0) & (vr_match < 0)
    apsis_inds = np.flatnonzero(flip)

    return {
        'apsis_inds': apsis_inds,
        'apsis_ids': ids_prev_[apsis_inds],
        'ids_match': ids[inds_match],
        'inds_match': inds_match,
        'inds_departed': inds_departed,
        'angle_changes': np.arccos(
            np.einsum('...i,...i', rhat_prev_, rhat_match)),
    }


def calc_angles(npart, angles_prev, apsis_dict):
    """float16 swept-angle accumulator; reference ``track_orbits.py:330-351``.
    """
    keep = np.ones(len(angles_prev), dtype=bool)
    keep[apsis_dict['inds_departed']] = False
    running = angles_prev[keep] + apsis_dict['angle_changes']
    apsis_angles = running[apsis_dict['apsis_inds']].copy()
    running[apsis_dict['apsis_inds']] = 0.0
    angles = np.zeros(npart)
    angles[apsis_dict['inds_match']] = running
    return angles.astype(np.float16), apsis_angles.astype(np.float16)


# ---------------------------------------------------------------------------
# track_orbits driver (reference track_orbits.py:9-244, 354-397)
# ---------------------------------------------------------------------------

class SnapshotState:
    """What the reference carries from one snapshot to the next
    (``track_orbits.py:234-240``)."""
    __slots__ = ('rhats', 'radial_vels', 'ids', 'angles', 'region_slices',
                 'halo_exists')


def track_snapshot(snapshot, halo_exists, region_positions, region_bulk_vels,
                   H, mode, prev):
    """One iteration of the per-halo loop + result assembly
    (reference ``track_orbits.py:147-217``).

    ``prev`` is the previous ``SnapshotState`` or ``None`` for the first
    processed snapshot.  Returns ``(state, out)`` where ``out`` is ``None`` for
    the first snapshot and otherwise a dict with the arrays that
    ``save_to_file`` receives.
    """
    region_offsets = list(snapshot['region_offsets']) + [len(snapshot['ids'])]
    region_slices = np.array(
        list(zip(region_offsets[:-1], region_offsets[1:])))
    ids = snapshot['ids']

    rhats, vrs, bulks, angles = [], [], [], []
    ev_ids, ev_angles = [], []
    for j, hind in enumerate(halo_exists):
        sl = region_slices[j]
        rh, vr, bv = region_frame(
            snapshot, sl, region_positions[j],
            None if region_bulk_vels is None else region_bulk_vels[j], H)
        npart = int(sl[1] - sl[0])
        ang = np.zeros(npart, dtype=np.float16)
        if prev is not None and hind in prev.halo_exists:
            k = int(np.flatnonzero(prev.halo_exists == hind)[0])
            plo, phi = prev.region_slices[k]
            d = compare_radial_velocities(
                ids[sl[0]:sl[1]], prev.ids[plo:phi], vr,
                prev.radial_vels[plo:phi], rh, prev.rhats[plo:phi], mode)
            ang, eang = calc_angles(npart, prev.angles[plo:phi], d)
            ev_ids.append(d['apsis_ids'])
            ev_angles.append(eang)
        rhats.append(rh)
        vrs.append(vr)
        bulks.append(bv)
        angles.append(ang)

    state = SnapshotState()
    state.rhats = np.concatenate(rhats)
    state.radial_vels = np.concatenate(vrs)
    state.ids = ids
    state.angles = np.concatenate(angles)
    state.region_slices = region_slices
    state.halo_exists = halo_exists

    out = None
    if prev is not None:
        out = {
            'apsis_offsets': np.cumsum([0] + [len(x) for x in ev_ids]),
            'apsis_ids': np.concatenate(ev_ids),
            'apsis_angles': np.concatenate(ev_angles),
            'hinds': np.where(np.isin(halo_exists, prev.halo_exists))[0],
            'bulk_velocities': np.array(bulks),
        }
    return state, out


def track_orbits(snapshot_numbers, main_branches, regions, load_snapshot_data,
                 savefile, mode='pericentric', checkpoint=False, resume=False,
                 npool=1, verbose=False, storage=None):
    """Restatement of the reference entry point ``track_orbits.py:9-244`` and
    its writers ``:354-397``.  ``npool`` is accepted and ignored (the pool only
    changes scheduling, not results)."""
    h5 = storage or _default_storage
    if len(main_branches) != len(snapshot_numbers):
        raise ValueError("len(main_branches) != len(snapshot_numbers)")
    if mode not in MODES:
        raise ValueError("mode must be 'pericentric' or 'apocentric'")

    main_branches = np.asarray(main_branches)
    if main_branches.ndim == 1:
        main_branches = main_branches[:, np.newaxis]
    snapshot_numbers = np.asarray(snapshot_numbers)
    order = np.argsort(snapshot_numbers)
    snapshot_numbers = snapshot_numbers[order]
    main_branches = main_branches[order]

    if resume:
        with h5.File(savefile, 'r') as hf:
            last = int(list(hf.keys())[-1].split('_')[1])
        first = int(np.flatnonzero(snapshot_numbers == last)[0])
        snapshot_numbers = snapshot_numbers[first:]
        main_branches = main_branches[first:]

    istart, started, prev = 0, False, None
    for i, (halo_ids, snap_no) in enumerate(
            zip(main_branches, snapshot_numbers)):
        halo_exists = np.flatnonzero(halo_ids != -1)
        if len(halo_exists) == 0:
            if not started:
                istart = i + 1
            continue
        halo_ids_ = halo_ids[halo_exists]
        region_positions, region_radii, region_bulk_vels = regions(
            snap_no, halo_ids_)
        snapshot = load_snapshot_data(snap_no, region_positions, region_radii)
        if len(snapshot['coordinates']) == 0:
            if not started:
                istart = i + 1
            continue
        started = True

        H = hubble_parameter(
            snapshot['redshift'], snapshot['H0'], snapshot['Omega_m'],
            snapshot['Omega_L'], snapshot.get('Omega_k', 0))

        if i == 0 and not resume:
            with h5.File(savefile, 'w') as hf:
                hf.attrs['mode'] = mode
                if 'box_size' in snapshot:
                    hf.attrs['box_size'] = snapshot['box_size']

        state, out = track_snapshot(
            snapshot, halo_exists, region_positions, region_bulk_vels, H, mode,
            prev if i > istart else None)

        if i > istart:
            hinds = out['hinds']
            with h5.File(savefile, 'r+') as hf:
                g = hf.create_group('snapshot_%03d' % snap_no)
                g.create_dataset('region_offsets', data=out['apsis_offsets'])
                g.create_dataset(mode[:-3] + 'er_IDs', data=out['apsis_ids'])
                g.create_dataset('angles', data=out['apsis_angles'])
                g.create_dataset('halo_IDs', data=halo_ids_[hinds])
                if snap_no != snapshot_numbers[-1]:
                    g.create_dataset(
                        'final_descendant_IDs',
                        data=main_branches[-1][prev.halo_exists])
                g.create_dataset('region_radii', data=region_radii[hinds])
                g.create_dataset(
                    'region_positions', data=region_positions[hinds])
                g.create_dataset(
                    'bulk_velocities', data=out['bulk_velocities'][hinds])
            if checkpoint:
                with h5.File(savefile + '.checkpoint', 'w') as hf:
                    hf.create_dataset('angles', data=state.angles)
        elif resume:
            with h5.File(savefile + '.checkpoint', 'r') as hf:
                state.angles = hf['angles'][:]
        prev = state


# ---------------------------------------------------------------------------
# on-the-fly variant (reference track_orbits_onthefly.py)
# ---------------------------------------------------------------------------

def repack(arr, length, inds):
    """Scatter ``arr`` into a -1-filled array of leading length ``length``
    (reference ``track_orbits_onthefly.py:61-68``)."""
    arr = np.asarray(arr)
    out = -np.ones((length,) + arr.shape[1:], dtype=arr.dtype)
    out[inds] = arr
    return out


def region_frame_onthefly(snapshot, region_slices, region_positions):
    """Reference ``track_orbits_onthefly.py:71-120``: no Hubble flow, bulk
    velocity always derived, work arrays in the snapshot's own dtypes."""
    coords = snapshot['coordinates']
    vels = snapshot['velocities']
    region_coords = np.empty(np.shape(coords), dtype=coords.dtype)
    for sl, pos in zip(region_slices, region_positions):
        delta = coords[slice(*sl)] - pos
        if 'box_size' in snapshot:
            delta = minimum_image(delta, snapshot['box_size'])
        region_coords[slice(*sl), :] = delta

    region_vels = np.empty(np.shape(vels), dtype=vels.dtype)
    bulk = []
    weighted = isinstance(snapshot['masses'], np.ndarray)
    for sl in region_slices:
        v = vels[slice(*sl)]
        if weighted:
            m = snapshot['masses'][slice(*sl)]
            b = np.sum(m[:, np.newaxis] * v, axis=0) / np.sum(m)
        else:
            b = np.mean(v, axis=0)
        region_vels[slice(*sl), :] = v - b
        bulk.append(b)

    rads = np.sqrt(np.einsum('...i,...i', region_coords, region_coords))
    rhats = region_coords / rads[:, np.newaxis]
    radial_vels = np.einsum('...i,...i', region_vels, rhats)
    return rhats, radial_vels, np.array(bulk)


def compare_onthefly(ids, ids_prev, radial_vels, radial_vels_prev, rhat,
                     rhat_prev, region_slices, region_slices_prev, mode):
    """Reference ``track_orbits_onthefly.py:123-205``."""
    tag = mode[:8] + 'er'
    ev_ids, ev_inds, entered, departed, matched, angles = \
        [], [], [], [], [], []
    empty = np.array([], dtype=ids.dtype)
    for slp, sl in zip(region_slices_prev, region_slices):
        cur = ids[slice(*sl)]
        if slp[1] - slp[0] > 0:
            old = ids_prev[slice(*slp)]
            survives = np.isin(old, cur)
            old_ = old[survives]
            vr_old = radial_vels_prev[slice(*slp)][survives]
            rh_old = rhat_prev[slice(*slp)][survives]
            m = ordered_match(cur, old_)
            vr_new = radial_vels[slice(*sl)][m]
            rh_new = rhat[slice(*sl)][m]
            if mode == 'pericentric':
                flip = (vr_old < 0) & (vr_new > 0)
            else:
                flip = (vr_old > 0) & (vr_new < 0)
            k = np.flatnonzero(flip)
            ev_inds.append(k)
            ev_ids.append(old_[k])
            entered.append(np.setdiff1d(cur, old))
            departed.append(np.setdiff1d(old, cur))
            matched.append(cur[m])
            angles.append(np.arccos(np.einsum('...i,...i', rh_old, rh_new)))
        else:
            entered.append(cur)
            for lst in (ev_inds, ev_ids, departed, matched, angles):
                lst.append(empty)

    def offs(lst):
        return np.cumsum([0] + [len(x) for x in lst])

    return {
        tag + '_ids': np.concatenate(ev_ids),
        tag + '_inds': np.concatenate(ev_inds),
        tag + '_offsets': offs(ev_ids),
        'entered_ids': np.concatenate(entered),
        'entered_offsets': offs(entered),
        'departed_ids': np.concatenate(departed),
        'departed_offsets': offs(departed),
        'matched_ids': np.concatenate(matched),
        'matched_offsets': offs(matched),
        'angle_changes': np.concatenate(angles),
    }


def track_orbits_onthefly(snapshot_number, progenitor_links, regions,
                          load_snapshot_data, savefile, mode='pericentric',
                          verbose=False, storage=None):
    """Reference ``track_orbits_onthefly.py:8-58`` + writer ``:208-252``."""
    h5 = storage or _default_storage
    if mode not in MODES:
        raise ValueError("mode must be 'pericentric' or 'apocentric'")
    ids, rhats, vrs, rpos, rrad, bulk, slices = [], [], [], [], [], [], []
    box_size = None
    for s, links in zip([snapshot_number, snapshot_number - 1],
                        progenitor_links):
        links = np.asarray(links)
        exists = np.flatnonzero(links != -1)
        pos, rad = regions(s, links[exists])
        pos_ = repack(pos, len(links), exists)
        rpos.append(pos_)
        rrad.append(repack(rad, len(links), exists))
        snapshot = load_snapshot_data(s, pos, rad)
        ids.append(snapshot['ids'])
        offsets = list(snapshot['region_offsets']) + [len(snapshot['ids'])]
        sl = repack(np.array(list(zip(offsets[:-1], offsets[1:]))),
                    len(links), exists)
        slices.append(sl)
        rh, vr, bv = region_frame_onthefly(snapshot, sl, pos_)
        rhats.append(rh)
        vrs.append(vr)
        bulk.append(bv)
        box_size = snapshot['box_size'] if 'box_size' in snapshot else None

    d = compare_onthefly(ids[0], ids[1], vrs[0], vrs[1], rhats[0], rhats[1],
                         slices[0], slices[1], mode)
    tag = mode[:8] + 'er'
    with h5.File(savefile.format('%0.3d' % snapshot_number), 'w') as hf:
        hf.create_dataset(tag + '_offsets', data=d[tag + '_offsets'])
        hf.create_dataset(tag + '_IDs', data=d[tag + '_ids'])
        hf.create_dataset('angles', data=d['angle_changes'])
        hf.create_dataset('entered_offsets', data=d['entered_offsets'])
        hf.create_dataset('entered_IDs', data=d['entered_ids'])
        hf.create_dataset('departed_offsets', data=d['departed_offsets'])
        hf.create_dataset('departed_IDs', data=d['departed_ids'])
        hf.create_dataset('progenitor_links', data=progenitor_links)
        hf.create_dataset('region_radii', data=rrad)
        hf.create_dataset('region_positions', data=rpos)
        hf.create_dataset('bulk_velocities', data=bulk)
        if box_size is not None:
            hf.attrs['box_size'] = box_size
    return d


# ---------------------------------------------------------------------------
# progenitors.py
# ---------------------------------------------------------------------------

def get_central_particle_ids(snapshot, halo_positions, n=100):
    """IDs of the ``n`` innermost particles of every region block, ordered by
    radius.  Reference ``progenitors.py:5-56`` (float64 work array, default
    ``argsort``)."""
    offsets = list(snapshot['region_offsets']) + [len(snapshot['ids'])]
    bounds = list(zip(offsets[:-1], offsets[1:]))
    coords = snapshot['coordinates']
    delta = np.empty(np.shape(coords))
    for (lo, hi), pos in zip(bounds, halo_positions):
        d = coords[lo:hi] - pos
        if 'box_size' in snapshot:
            d = minimum_image(d, snapshot['box_size'])
        delta[lo:hi, :] = d
    rads = np.sqrt(np.einsum('...i,...i', delta, delta))
    picks = [snapshot['ids'][np.argsort(rads[lo:hi])[:n] + lo]
             for lo, hi in bounds]
    starts = np.cumsum([0] + [len(p) for p in picks])[:-1]
    return np.hstack(picks), starts


def find_main_progenitors(halo_pids, halo_offsets, tracked_pids,
                          tracked_offsets):
    """Plurality host halo of every descendant's tracked particles.

    Reference ``progenitors.py:59-117``: a tracked ID that occurs more than
    once is only counted at its first occurrence (``:82-84``); a descendant
    with no tracked particle in any halo gets -1 (``:112-113``); ties go to the
    smallest halo index (sorted ``unique`` + first ``argmax``, ``:107-115``).
    Returns a list like the reference.
    """
    halo_pids = np.asarray(halo_pids)
    tracked_pids = np.asarray(tracked_pids)
    halo_offsets = np.asarray(halo_offsets)
    tracked_offsets = np.asarray(tracked_offsets)
    M = len(tracked_pids)

    _, first = np.unique(tracked_pids, return_index=True)
    is_first = np.zeros(M, dtype=bool)
    is_first[first] = True

    order = np.argsort(halo_pids, kind='stable')
    sorted_pids = halo_pids[order]
    where = np.searchsorted(sorted_pids, tracked_pids)
    where_c = np.minimum(where, max(len(sorted_pids) - 1, 0))
    found = is_first & (where < len(sorted_pids))
    if len(sorted_pids):
        found &= sorted_pids[where_c] == tracked_pids
    else:
        found[:] = False
    host = np.full(M, -1, dtype=np.int64)
    host[found] = np.searchsorted(
        halo_offsets, order[where_c[found]], side='right') - 1

    bounds = list(tracked_offsets) + [M]
    result = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        h = host[lo:hi]
        h = h[h != -1]
        if len(h) == 0:
            result.append(-1)
        else:
            vals, counts = np.unique(h, return_counts=True)
            result.append(vals[np.argmax(counts)])
    return result


# ---------------------------------------------------------------------------
# postprocessing.py (Apsides)
# ---------------------------------------------------------------------------

class Apsides:
    """Reference ``postprocessing.py:8-240``."""

    def __init__(self, filename, storage=None):
        self._h5 = storage or _default_storage
        self.filename = filename
        with self._h5.File(filename, 'r') as hf:
            keys = list(hf.keys())
            self.snapshot_numbers = np.array(
                [int(k.split('_')[1]) for k in keys])
            self.final_halo_ids = hf[keys[-1]]['halo_IDs'][:]
            self.mode = hf.attrs['mode']
            if 'box_size' in hf.attrs:
                self.box_size = hf.attrs['box_size']

    def collate_apsides(self, halo_ids=None, snapshot_number=None,
                        angle_cut=np.pi / 4, save_final_counts=False,
                        data_type=None, savefile=None, verbose=False):
        """Reference ``postprocessing.py:30-174``."""
        tag = self.mode[:-3] + 'er'
        if halo_ids is None:
            halo_ids = self.final_halo_ids
        elif len(np.intersect1d(self.final_halo_ids, halo_ids)) < len(
                halo_ids):
            self.missing_halo_ids = np.setdiff1d(
                halo_ids, self.final_halo_ids)
            raise ValueError("halo_ids contains halos that were not tracked")
        if snapshot_number is None:
            last = len(self.snapshot_numbers) - 1
        else:
            last = int(np.flatnonzero(
                self.snapshot_numbers == snapshot_number)[0])

        pools = None
        for s in self.snapshot_numbers[:last + 1]:
            is_final = s == self.snapshot_numbers[-1]
            with self._h5.File(self.filename, 'r') as hf:
                g = hf['snapshot_%03d' % s]
                region_positions = g['region_positions'][:]
                region_radii = g['region_radii'][:]
                bulk_velocities = g['bulk_velocities'][:]
                halo_ids_current = g['halo_IDs'][:]
                halo_ids_final = halo_ids_current if is_final else \
                    g['final_descendant_IDs'][:]
                common = np.intersect1d(halo_ids_final, halo_ids)
                hinds1 = ordered_match(halo_ids_final, common)
                hinds2 = ordered_match(halo_ids, common)
                if len(g[tag + '_IDs']) == 0:
                    continue
                if pools is None:
                    idtype = g[tag + '_IDs'].dtype if data_type is None \
                        else data_type
                    pools = [np.array([], dtype=idtype) for _ in halo_ids]
                offs = g['region_offsets'][:]
                ev_ids = g[tag + '_IDs'][:]
                ev_ang = g['angles'][:]
                for h1, h2 in zip(hinds1, hinds2):
                    lo, hi = offs[h1], offs[h1 + 1]
                    sel = ev_ang[lo:hi] > angle_cut
                    pools[h2] = np.append(pools[h2], ev_ids[lo:hi][sel])

            uniq, counts, lens = [], [], []
            for i, pool in enumerate(pools):
                u, c = np.unique(pool, return_counts=True)
                uniq.append(u)
                counts.append(c)
                if i in hinds2:
                    lens.append(len(u))
            with self._h5.File(savefile, 'a') as hf:
                g = hf.create_group('snapshot_%03d' % s)
                g.create_dataset('particle_IDs', data=np.concatenate(uniq))
                g.create_dataset(tag + '_counts', data=np.concatenate(counts))
                g.create_dataset(
                    'halo_offsets', data=np.cumsum([0] + lens)[:-1])
                if not is_final:
                    g.create_dataset(
                        'final_descendant_IDs', data=halo_ids_final[hinds1])
                g.create_dataset('halo_IDs', data=halo_ids_current[hinds1])
                g.create_dataset(
                    'halo_positions', data=region_positions[hinds1])
                g.create_dataset(
                    'halo_velocities', data=bulk_velocities[hinds1])
                g.create_dataset('region_radii', data=region_radii[hinds1])

        if save_final_counts:
            self.save_final_apsis_counts(savefile)

    def save_final_apsis_counts(self, collated_file, snapshot_numbers=None,
                                verbose=False):
        """Reference ``postprocessing.py:176-240``."""
        tag = self.mode[:-3] + 'er'
        with self._h5.File(collated_file, 'r+') as hf:
            keys = np.array(list(hf.keys()))
            fin = hf[keys[-1]]
            ids_final = fin['particle_IDs'][:]
            counts_final = fin[tag + '_counts'][:]
            halo_ids = fin['halo_IDs'][:]
            fo = list(fin['halo_offsets'][:]) + [len(ids_final)]
            if snapshot_numbers is None:
                todo = keys[:-1]
            else:
                nums = np.array([int(k.split('_')[-1]) for k in keys])
                todo = keys[np.isin(nums, snapshot_numbers)]
            for key in todo:
                g = hf[key]
                ids = g['particle_IDs'][:]
                desc = g['final_descendant_IDs'][:]
                so = list(g['halo_offsets'][:]) + [len(ids)]
                hinds = ordered_match(halo_ids, desc)
                retro = np.empty(len(ids))
                for h2, h1 in enumerate(hinds):
                    flo, fhi = fo[h1], fo[h1 + 1]
                    lo, hi = so[h2], so[h2 + 1]
                    k = ordered_match(ids_final[flo:fhi], ids[lo:hi])
                    retro[lo:hi] = counts_final[flo:fhi][k]
                g.create_dataset(tag + '_counts_final', data=retro)


# ---------------------------------------------------------------------------
# region extraction (the example loader, example_script.py:36-67)
# ---------------------------------------------------------------------------
def extract_regions(coordinates, region_positions, region_radii, box_size=None):
    """Particle indices within each region and ``region_offsets``: restates the
    selection loop of the reference's example loader (``example_script.py:
    50-58``) with the reference's helpers (``utils.py:13-33``): for every
    region, ``np.argwhere(vector_norm(recenter_coordinates(coordinates -
    position, box_size)) < radius).flatten()``.  O(N x n_regions): test sizes
    only."""
    region_inds = []
    for position, radius in zip(np.atleast_2d(region_positions),
                                np.atleast_1d(region_radii)):
        d = coordinates - position
        if box_size is not None:
            boxsize = box_size
            if isinstance(boxsize, (float, np.floating, int, np.integer)):
                boxsize = boxsize * np.ones(3)
            for dim, bs in enumerate(boxsize):
                d[np.argwhere((d[:, dim] > bs / 2)), dim] -= bs
                d[np.argwhere((d[:, dim] < -bs / 2)), dim] += bs
        r = np.sqrt(np.einsum('...i,...i', d, d))
        region_inds.append(np.argwhere(r < radius).flatten())
    region_lens = [len(inds) for inds in region_inds]
    region_offsets = np.cumsum([0] + region_lens)[:-1]
    region_inds = np.hstack(region_inds).astype(int) if region_inds else \
        np.zeros(0, dtype=int)
    return region_inds, region_offsets
