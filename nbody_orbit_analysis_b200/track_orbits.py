"""Drop-in ``track_orbits`` entry point on B200.

Same signature, callback protocol and result-file layout as the reference
``orbitanalysis/track_orbits.py:9-11``; the per-halo
numpy loop (``:147-217``) is replaced by ``OrbitTracker.submit/collect`` -- one
fused CUDA kernel per snapshot plus an ordered event compaction.  Host code here
only does what the reference's driver does outside ``track(j)``: argument
validation, snapshot ordering, resume bookkeeping and writing the HDF5 layout.

The driver is a software pipeline: while snapshot s is on the GPU, the user's
``regions`` / ``load_snapshot_data`` callbacks run for s+1 and the result group
of s-1 is written, so callbacks, host->device copies, kernels and file writes
overlap (the reference is strictly serial, ``track_orbits.py:104-240``).

Multi-GPU (one process per GPU, ``torch.distributed`` initialised by the
launcher, e.g. ``torchrun``): particles are sharded by ``id mod world``
(SURVEY.md section 8(e)); rank 0's catalogue is broadcast, every rank tracks its
shard, the event lists are merged in the reference's order
(``track_orbits.py:199-217, 315-316``) and rank 0 writes the file.  The loader
may return the whole snapshot (it is sharded here) or this rank's shard
together with ``snapshot['_gpos']`` = each particle's position in the unsharded
arrays (loader-side sharding: no rank ever holds the whole snapshot).

Deliberate, documented deviations (DESIGN.md "Quirks"):
* ``npool`` is accepted and ignored (it only changes scheduling);
* the savefile is initialised at the first *processed* snapshot rather than
  only when row 0 is processed (the reference crashes on ``'r+'`` otherwise,
  ``track_orbits.py:140,375``);
* a derived bulk velocity (``regions`` returned ``None``) is accumulated in
  float64 (reference: sequential accumulation in the input dtype); sharded runs
  need catalogue bulk velocities;
* the ``ValueError`` of a snapshot without matched halos (``:216``) is raised
  when that snapshot's results are collected, one loop iteration later.
"""
import time
from collections import deque

import numpy as np

from . import storage
from .sharded import shard_snapshot
from .tracker import OrbitTracker, require_cuda
from .utils import hubble_parameter


def track_orbits(snapshot_numbers, main_branches, regions, load_snapshot_data,
                 savefile, mode='pericentric', checkpoint=False, resume=False,
                 npool=1, verbose=True, device=None, comm=None):
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
    checkpoint, resume : as in the reference (``:93-101, 229-232, 390-394``).
        Extension: ``checkpoint='state'`` also saves the carried device state
        (records + ID table) next to the reference's
        ``angles`` dataset; ``resume=True`` then continues at the snapshot AFTER
        the last saved one without reloading it (with a plain checkpoint the
        last saved snapshot is re-processed, like in the reference)
    npool : ignored
    device : optional torch device (extension; default current CUDA device)
    comm : extension.  None (default): shard over the ranks of an initialised
        ``torch.distributed`` process group, single GPU otherwise; False: never
        shard; a ``sharded.Comm``: use it
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

    restored = None
    if resume:
        if verbose:
            print('Resuming from file...\n')
        with storage.File(savefile, 'r') as hf:
            last = int(list(hf.keys())[-1].split('_')[1])
        first = int(np.flatnonzero(snapshot_numbers == last)[0])
        restored = _read_state(savefile + '.checkpoint', last)
        if restored is not None:
            first += 1                # the carried state of `last` is restored
        snapshot_numbers = snapshot_numbers[first:]
        main_branches = main_branches[first:]

    tracker = OrbitTracker(mode=mode, device=device)
    comm = _communicator(comm)
    sharded_run = comm is not None
    rank = comm.rank if sharded_run else 0
    writer = rank == 0
    verbose = verbose and writer
    if sharded_run:
        tracker.events_on_device = True
    ckpt_name = savefile + '.checkpoint' + ('.%d' % rank if sharded_run else '')
    tag = mode[:-3] + 'er'
    istart, started, initialised = 0, False, bool(resume)
    prev_halo_exists = None
    last_snap = snapshot_numbers[-1] if len(snapshot_numbers) else None
    if restored is not None:
        if sharded_run:
            raise ValueError("device-state checkpoints are single-GPU")
        tracker.load_state(restored)
        prev_halo_exists = np.asarray(restored['halo_exists'])
        istart, started = -1, True

    writer_thread = _Writer()

    def write(e, res):
        """Result group of one snapshot (``track_orbits.py:366-397``): checked
        here, written by the writer thread while the next snapshot is staged
        (its arrays stay valid: they live in the tracker's pinned result ring,
        which comes round three submits later, and at most one write is in
        flight)."""
        if res.apsis_ids is None or len(res.hinds) == 0:
            # reference: np.concatenate([]) of an empty list (:216)
            raise ValueError("need at least one array to concatenate")
        if res.host_ready is not None:
            res.host_ready.synchronize()
        writer_thread.submit(write_files, e, res)

    def write_files(e, res):
        if checkpoint:
            with storage.File(ckpt_name, 'w') as hf:
                hf.create_dataset('angles', data=e.angles)
                if e.state is not None:
                    g = hf.create_group('b200_state')
                    g.attrs['snapshot'] = int(e.snap_no)
                    for k, v in e.state.items():
                        g.create_dataset(k, data=v)
        if not writer:
            return
        t0 = time.time()
        hinds = res.hinds
        with storage.File(savefile, 'r+') as hf:
            g = hf.create_group('snapshot_%03d' % e.snap_no)
            g.create_dataset('region_offsets', data=res.apsis_offsets)
            g.create_dataset(tag + '_IDs', data=res.apsis_ids)
            g.create_dataset('angles', data=res.apsis_angles)
            g.create_dataset('halo_IDs', data=e.halo_ids[hinds])
            if e.snap_no != last_snap:
                g.create_dataset('final_descendant_IDs',
                                 data=main_branches[-1][e.prev_halo_exists])
            g.create_dataset('region_radii', data=e.radii[hinds])
            g.create_dataset('region_positions', data=e.positions[hinds])
            g.create_dataset('bulk_velocities', data=res.bulk_velocities[hinds])
        if verbose:
            print('Saved snapshot {} to file ({} s)\n'.format(
                '%03d' % e.snap_no, time.time() - t0))

    pending = deque()      # submitted, not collected
    merging = deque()      # collected, event exchange in flight (multi-GPU)

    def finish(e):
        res = tracker.collect(e.p)
        e.p = None
        e.angles = res.angles
        if verbose:
            print('Finished {} detection for snapshot {} ({} s after its '
                  'submission)\n'.format(tag, '%03d' % e.snap_no,
                                         time.time() - e.t0))
        if not e.has_events:
            return
        if not sharded_run:
            write(e, res)
            return
        if res.apsis_offsets is None or len(res.hinds) == 0:
            raise ValueError("need at least one array to concatenate")
        merging.append((e, comm.start_merge(tracker, res, to_host=True)))
        while len(merging) > 1:
            e0, h = merging.popleft()
            write(e0, comm.finish_merge(h))

    try:
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
            if sharded_run:
                # the catalogue comes from rank 0; the other ranks do
                # not call `regions`
                cat = regions(snap_no, halo_ids_) if writer else (None, None, None)
                region_positions, region_radii, region_bulk_vels = \
                    _broadcast_catalogue(comm, cat, len(halo_ids_))
                if region_bulk_vels is None:
                    raise ValueError(
                        "sharded tracking needs catalogue bulk velocities "
                        "(regions() returned None for them)")
            else:
                region_positions, region_radii, region_bulk_vels = regions(
                    snap_no, halo_ids_)
            snapshot = load_snapshot_data(snap_no, region_positions, region_radii)
            gpos = None
            if sharded_run:
                if '_gpos' in snapshot:
                    gpos = np.ascontiguousarray(snapshot['_gpos'], dtype=np.int64)
                else:
                    snapshot, gpos = shard_snapshot(snapshot, rank, comm.world)
                empty = _all_empty(comm, len(snapshot['coordinates']))
            else:
                empty = len(snapshot['coordinates']) == 0
            if empty:
                if not started:
                    istart = i + 1
                continue
            started = True

            H = hubble_parameter(
                snapshot['redshift'], snapshot['H0'], snapshot['Omega_m'],
                snapshot['Omega_L'], snapshot.get('Omega_k', 0))

            if not initialised:
                if writer:
                    with storage.File(savefile, 'w') as hf:
                        hf.attrs['mode'] = mode
                        if 'box_size' in snapshot:
                            hf.attrs['box_size'] = snapshot['box_size']
                initialised = True
                if verbose:
                    print('Savefile initialized\n')

            if i <= istart:
                tracker.prev = None       # first processed snapshot: no matching
            e = _Entry()
            e.snap_no, e.halo_ids, e.t0 = snap_no, halo_ids_, time.time()
            # (copies: the group is written after the next callbacks have run,
            # and a catalogue reader may reuse its arrays)
            e.positions, e.radii = np.array(region_positions), \
                np.array(region_radii)
            if region_bulk_vels is not None:
                region_bulk_vels = np.array(region_bulk_vels)
            e.prev_halo_exists, e.has_events = prev_halo_exists, i > istart
            e.p = tracker.submit(
                snapshot, halo_exists, e.positions, region_bulk_vels, H,
                want_angles=checkpoint, gpos=gpos)
            if resume and i <= istart:
                with storage.File(ckpt_name, 'r') as hf:
                    tracker.load_angles(hf['angles'][:])
            e.state = tracker.save_state() if checkpoint == 'state' else None
            pending.append(e)
            # results of the PREVIOUS snapshot: collected and written while this one
            # is on the GPU
            while len(pending) > 1:
                finish(pending.popleft())
            prev_halo_exists = halo_exists

        while pending:
            finish(pending.popleft())
        while merging:
            e0, h = merging.popleft()
            write(e0, comm.finish_merge(h))
    except BaseException:
        # (the result file ends at a whole group; the error that stopped the run
        # is the one the caller sees)
        writer_thread.wait(swallow=True)
        raise
    writer_thread.wait()

    if verbose:
        print('Finished {} detection for all snapshots in {} s\n'.format(
            tag, time.time() - t_start))


def _read_state(path, snapshot):
    """The device state saved by ``checkpoint='state'`` for ``snapshot``, or
    None (plain checkpoint, other snapshot, no file)."""
    import os
    if not os.path.exists(path):
        return None
    with storage.File(path, 'r') as hf:
        if 'b200_state' not in list(hf.keys()):
            return None
        g = hf['b200_state']
        if int(g.attrs['snapshot']) != int(snapshot):
            return None
        return {k: g[k][:] for k in g.keys()}


class _Writer:
    """Runs the result-file writes of the driver on a background thread, one at a
    time and in order: ``submit`` first waits for the previous write.  The file
    layer releases the GIL inside its large ``write`` calls, and the staging
    copy / CUDA calls of the main thread release it too, so writing snapshot
    s-1 overlaps staging snapshot s+1.  An exception raised by a write surfaces
    at the next ``submit`` / ``wait`` on the calling thread."""

    def __init__(self):
        self._thread = None
        self._error = None

    def _run(self, fn, args):
        try:
            fn(*args)
        except BaseException as exc:          # re-raised on the driver's thread
            self._error = exc

    def submit(self, fn, *args):
        import os
        import threading
        self.wait()
        if os.environ.get('OA_WRITER_THREAD', '1') == '0':
            fn(*args)                         # (escape hatch: write in line)
            return
        self._thread = threading.Thread(target=self._run, args=(fn, args),
                                        name='orbit-b200-writer')
        self._thread.start()

    def wait(self, swallow=False):
        thread, self._thread = self._thread, None
        if thread is not None:
            thread.join()
        error, self._error = self._error, None
        if error is not None and not swallow:
            raise error


class _Entry:
    """A submitted snapshot: what its result group needs besides the events."""
    pass


def _communicator(comm):
    """``comm`` argument -> ``sharded.Comm`` or None (single GPU).  Default:
    shard when ``torch.distributed`` is initialised with more than one rank."""
    if comm is False:
        return None
    if comm is None or comm is True:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()
                and dist.get_world_size() > 1):
            return None
        from .sharded import Comm
        return Comm()
    return comm


def _broadcast_catalogue(comm, cat, n_h):
    """Rank 0's ``regions()`` output on every rank, dtypes included."""
    import torch
    import torch.distributed as dist
    pos, rad, bulk = cat
    meta = torch.zeros(4, dtype=torch.int64, device=comm.device)
    if comm.rank == 0:
        pos, rad = np.asarray(pos), np.asarray(rad)
        code = {np.dtype(np.float32): 1, np.dtype(np.float64): 2}
        meta[0] = code.get(pos.dtype, 2)
        meta[1] = code.get(rad.dtype, 2)
        meta[2] = 0 if bulk is None else code.get(np.asarray(bulk).dtype, 2)
        meta[3] = len(rad)
    dist.broadcast(meta, src=0)
    m = meta.cpu().numpy()
    dt = {1: np.float32, 2: np.float64}
    if comm.rank != 0:
        n = int(m[3])
        pos = np.zeros((n, 3), dtype=dt[int(m[0])])
        rad = np.zeros(n, dtype=dt[int(m[1])])
        bulk = np.zeros((n, 3), dtype=dt[int(m[2])]) if m[2] else None
    else:
        pos = pos.astype(dt[int(m[0])], copy=False)
        rad = rad.astype(dt[int(m[1])], copy=False)
        bulk = None if bulk is None else \
            np.asarray(bulk).astype(dt[int(m[2])], copy=False)
    return comm.broadcast_catalogue(pos, rad, bulk)


def _all_empty(comm, n_local):
    """True when no rank holds a particle of this snapshot."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([n_local], dtype=torch.int64, device=comm.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item()) == 0
