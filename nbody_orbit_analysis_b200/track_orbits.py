"""Drop-in ``track_orbits`` entry point on B200.

Same signature, callback protocol and result-file layout as the reference
``orbitanalysis/track_orbits.py:9-11``; the per-halo
numpy loop (``:147-217``) is replaced by ``OrbitTracker.step`` -- one fused CUDA
kernel per snapshot plus an ordered event compaction.  Host code here only does
what the reference's driver does outside ``track(j)``: argument validation,
snapshot ordering, resume bookkeeping and writing the HDF5 layout.

Deliberate, documented deviations (DESIGN.md "Quirks"):
* ``npool`` is accepted and ignored (it only changes scheduling);
* the savefile is initialised at the first *processed* snapshot rather than
  only when row 0 is processed (the reference crashes on ``'r+'`` otherwise,
  ``track_orbits.py:140,375``);
* a derived bulk velocity (``regions`` returned ``None``) is accumulated in
  float64 (reference: sequential accumulation in the input dtype).
"""
import time

import numpy as np

from . import storage
from .tracker import OrbitTracker, require_cuda
from .utils import hubble_parameter


def track_orbits(snapshot_numbers, main_branches, regions, load_snapshot_data,
                 savefile, mode='pericentric', checkpoint=False, resume=False,
                 npool=1, verbose=True, device=None):
    """Track the orbits of particles in gravitating systems (GPU path).

    Parameters are those of the reference (``track_orbits.py:13-71``):

    snapshot_numbers : (n_snap,) array_like
    main_branches : (n_snap, n_halo) array_like, -1 where no progenitor exists
    regions : callable(snapshot_number, halo_ids) ->
        (positions (n,3), radii (n,), bulk_velocities (n,3) or None)
    load_snapshot_data : callable(snapshot_number, positions, radii) -> dict
        with ``ids``, ``coordinates``, ``velocities``, ``masses``,
        ``region_offsets``, optional ``box_size``, ``redshift``, ``H0``,
        ``Omega_m``, ``Omega_L``, optional ``Omega_k``
    savefile : str -- HDF5 result file (layout of ``track_orbits.py:366-397``)
    mode : 'pericentric' | 'apocentric'
    checkpoint, resume : as in the reference (``:93-101, 229-232, 390-394``)
    npool : ignored
    device : optional torch device (extension; default current CUDA device)
    """
    if len(main_branches) != len(snapshot_numbers):
        raise ValueError(
            "Number of halo main branch nodes does not equal the number of "
            "snapshot numbers supplied. Must have len(main_branches) == "
            "len(snapshot_numbers).")
    if mode not in ('pericentric', 'apocentric'):
        raise ValueError(
            "Orbit detection mode not recognized. Please specify either "
            "'pericentric' or 'apocentric'.")
    require_cuda()

    t_start = time.time()
    main_branches = np.asarray(main_branches)
    if main_branches.ndim == 1:
        main_branches = main_branches[:, np.newaxis]
    snapshot_numbers = np.asarray(snapshot_numbers)
    order = np.argsort(snapshot_numbers)
    snapshot_numbers = snapshot_numbers[order]
    main_branches = main_branches[order]

    if resume:
        if verbose:
            print('Resuming from file...\n')
        with storage.File(savefile, 'r') as hf:
            last = int(list(hf.keys())[-1].split('_')[1])
        first = int(np.flatnonzero(snapshot_numbers == last)[0])
        snapshot_numbers = snapshot_numbers[first:]
        main_branches = main_branches[first:]

    tracker = OrbitTracker(mode=mode, device=device)
    tag = mode[:-3] + 'er'
    istart, started, initialised = 0, False, bool(resume)
    prev_halo_exists = None

    for i, (halo_ids, snap_no) in enumerate(
            zip(main_branches, snapshot_numbers)):
        if verbose:
            print('-' * 30, '\n')
            print('Snapshot {}\n'.format('%03d' % snap_no))

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

        if not initialised:
            with storage.File(savefile, 'w') as hf:
                hf.attrs['mode'] = mode
                if 'box_size' in snapshot:
                    hf.attrs['box_size'] = snapshot['box_size']
            initialised = True
            if verbose:
                print('Savefile initialized\n')

        t0 = time.time()
        if i <= istart:
            tracker.prev = None       # first processed snapshot: no matching
        res = tracker.step(
            snapshot, halo_exists, np.asarray(region_positions),
            region_bulk_vels, H, want_angles=checkpoint)
        if verbose:
            print('Finished {} detection for snapshot {} in {} s\n'.format(
                tag, '%03d' % snap_no, time.time() - t0))

        if i > istart:
            if res.apsis_ids is None or len(res.hinds) == 0:
                # reference: np.concatenate([]) of an empty list (:216)
                raise ValueError("need at least one array to concatenate")
            t0 = time.time()
            hinds = res.hinds
            with storage.File(savefile, 'r+') as hf:
                g = hf.create_group('snapshot_%03d' % snap_no)
                g.create_dataset('region_offsets', data=res.apsis_offsets)
                g.create_dataset(tag + '_IDs', data=res.apsis_ids)
                g.create_dataset('angles', data=res.apsis_angles)
                g.create_dataset('halo_IDs', data=halo_ids_[hinds])
                if snap_no != snapshot_numbers[-1]:
                    g.create_dataset(
                        'final_descendant_IDs',
                        data=main_branches[-1][prev_halo_exists])
                g.create_dataset(
                    'region_radii', data=np.asarray(region_radii)[hinds])
                g.create_dataset(
                    'region_positions',
                    data=np.asarray(region_positions)[hinds])
                g.create_dataset(
                    'bulk_velocities', data=res.bulk_velocities[hinds])
            if checkpoint:
                with storage.File(savefile + '.checkpoint', 'w') as hf:
                    hf.create_dataset('angles', data=res.angles)
            if verbose:
                print('Saved to file ({} s)\n'.format(time.time() - t0))
        elif resume:
            with storage.File(savefile + '.checkpoint', 'r') as hf:
                tracker.load_angles(hf['angles'][:])
        prev_halo_exists = halo_exists

    if verbose:
        print('Finished {} detection for all snapshots in {} s\n'.format(
            tag, time.time() - t_start))
