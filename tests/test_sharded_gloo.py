"""Multi-process test of the ID-sharded exchange (CPU, gloo, world_size 2).

Every rank takes the particles with ``id mod world == rank`` of the same
synthetic snapshots (``sharded.shard_snapshot``), produces its LOCAL apsis
events -- here with the CPU oracle, on the GPU box with the fused kernel -- and
runs the same exchange/merge code as the NCCL path (``Comm.broadcast_catalogue``
+ ``Comm.merge``: all-gather of variable-length event records, all-reduce of the
per-halo counts, ordering by position in the unsharded previous snapshot).  The
merged lists must equal the events of the unsharded oracle run bit for bit
(reference order, ``track_orbits.py:315-316``).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_halos, out_dir):
    sys.path.insert(0, REPO)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nbody_orbit_analysis_b200 import sharded
        from nbody_orbit_analysis_b200.synth import SynthSim
        from oracle import orbit_oracle as oracle

        sim = SynthSim(24000, n_halos, 4, dtype=np.float32,
                       catalogue_dtype=np.float32)
        comm = sharded.Comm(world, rank, device=torch.device('cpu'))
        exists = np.arange(n_halos)
        prev_local = prev_full = None
        gpos_prev = None
        for t, snap_no in enumerate(sim.snapshot_numbers):
            cat = sim.regions(snap_no, sim.main_branches[t])
            # only rank 0 "owns" the catalogue; the others get it by broadcast
            if rank != 0:
                cat = tuple(np.zeros_like(c) for c in cat)
            pos, rad, bulk = comm.broadcast_catalogue(*cat)
            ref_cat = sim.regions(snap_no, sim.main_branches[t])
            assert np.array_equal(pos, ref_cat[0])
            assert np.array_equal(bulk, ref_cat[2])
            full = sim.load_snapshot_data(snap_no, pos, rad)
            local, gpos = sharded.shard_snapshot(full, rank, world)
            assert np.all(local['ids'] % world == rank)

            state_l, out_l = oracle.track_snapshot(
                local, exists, pos, bulk, 0.0, 'pericentric', prev_local)
            state_f, out_f = oracle.track_snapshot(
                full, exists, pos, bulk, 0.0, 'pericentric', prev_full)
            if t > 0:
                # position of every local event particle in the unsharded
                # previous snapshot = the merge key
                ids_prev_local = prev_local_ids
                idx = {int(i): k for k, i in enumerate(ids_prev_local)}
                # events are (halo, id) pairs; a particle may sit in two halos,
                # so resolve the previous index inside the event's own block
                keys = np.empty(len(out_l['apsis_ids']), dtype=np.int64)
                offs = out_l['apsis_offsets']
                for h in range(n_halos):
                    lo, hi = prev_local_offsets[h], prev_local_offsets[h + 1]
                    block = {int(i): lo + k for k, i in
                             enumerate(ids_prev_local[lo:hi])}
                    for e in range(offs[h], offs[h + 1]):
                        keys[e] = gpos_prev[block[int(out_l['apsis_ids'][e])]]
                del idx

                def order(k):
                    return torch.sort(k, stable=True)[1]

                def take(src, perm):
                    return src[perm]
                ids_m, ang_m, off_m = comm.merge(
                    torch.from_numpy(keys),
                    torch.from_numpy(out_l['apsis_ids'].astype(np.int64)),
                    torch.from_numpy(out_l['apsis_angles'].view(np.int16)),
                    torch.from_numpy(np.diff(offs)), order, take)
                assert np.array_equal(ids_m.numpy(), out_f['apsis_ids'])
                assert np.array_equal(off_m, out_f['apsis_offsets'])
                assert np.array_equal(ang_m.numpy().view(np.float16),
                                      out_f['apsis_angles'], equal_nan=True)
                assert len(out_f['apsis_ids']) > 0
            prev_local, prev_full = state_l, state_f
            prev_local_ids = np.asarray(local['ids'])
            prev_local_offsets = np.append(local['region_offsets'],
                                           len(local['ids']))
            gpos_prev = gpos
        open(os.path.join(out_dir, 'ok_%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_halos', [5, 40])
def test_sharded_exchange_reproduces_unsharded_order(n_halos, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_halos, str(tmp_path)),
             nprocs=world, join=True)
    assert all(os.path.exists(str(tmp_path / ('ok_%d' % r)))
               for r in range(world))


def test_shard_snapshot_keeps_block_structure():
    sys.path.insert(0, REPO)
    from nbody_orbit_analysis_b200 import sharded
    rng = np.random.default_rng(3)
    ids = rng.permutation(1000).astype(np.int64)
    snap = {'ids': ids, 'coordinates': rng.random((1000, 3)),
            'velocities': rng.random((1000, 3)), 'masses': 1.0,
            'region_offsets': np.array([0, 300, 300, 720])}
    seen = []
    for r in range(3):
        loc, gpos = sharded.shard_snapshot(snap, r, 3)
        assert np.array_equal(loc['ids'], ids[gpos])
        offs = np.append(loc['region_offsets'], len(loc['ids']))
        full = np.append(snap['region_offsets'], 1000)
        for h in range(4):
            blk = gpos[offs[h]:offs[h + 1]]
            assert np.all((blk >= full[h]) & (blk < full[h + 1]))
        seen.append(gpos)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(1000))


# ---------------------------------------------------------------------------
# asynchronous exchange (Comm.start_merge / finish_merge), host side
# ---------------------------------------------------------------------------
def _async_worker(rank, world, port, mode, out_dir):
    """The bench's pipeline (exchange of snapshot k started after the submit of
    k+1 and finished one snapshot later) with the kernels replaced by their
    numpy restatements (tests/exchange_emul.py) and NCCL by gloo.  Event counts
    jump by orders of magnitude between snapshots, so the send buffers overflow
    and exchanges are repeated, also two in a row."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nbody_orbit_analysis_b200 import sharded
        import exchange_emul as emul
        emul.install(sharded)
        comm = sharded.Comm(world, rank, device=torch.device('cpu'))
        trk = emul.EmulTracker()
        n_prev, n_halos = 120000, 7
        counts = [0, 50, 3000, 2900, 40000, 90000, 100, 0, 39000, 95000, 5]
        rng = np.random.default_rng(11)
        expected, pending, finished = {}, None, []
        caps = []

        def finish(h):
            res = comm.finish_merge(h)
            if res.host_ready is not None:
                res.host_ready.synchronize()
            exp_ids, exp_ang, exp_off = expected.pop(res.step)
            assert res.n_events == len(exp_ids), (res.step, res.n_events, len(exp_ids), comm._cap)
            assert np.array_equal(res.apsis_offsets, exp_off)
            lo, hi = res.host_slice if res.host_slice is not None \
                else (0, res.n_events)
            if mode is True:
                assert (lo, hi) == (0, len(exp_ids))
            assert np.array_equal(res.apsis_ids, exp_ids[lo:hi])
            assert np.array_equal(res.apsis_angles.view(np.int16),
                                  exp_ang[lo:hi].view(np.int16))
            assert np.array_equal(res.d_ids.numpy(), exp_ids[lo:hi])
            finished.append((res.step, lo, hi, res.n_events))
            caps.append(comm._cap)

        for k, m in enumerate(counts):
            # unsharded previous snapshot: blocks of the halos, one owner per row
            starts = np.sort(rng.choice(n_prev, n_halos - 1, replace=False))
            starts = np.concatenate(([0], starts))
            pid = rng.permutation(n_prev).astype(np.int64) + 10 ** 12
            ev_pos = np.sort(rng.choice(n_prev, m, replace=False))
            ev_ang = rng.standard_normal(n_prev).astype(np.float16)
            expected[k] = (pid[ev_pos], ev_ang[ev_pos], np.append(
                np.searchsorted(ev_pos, starts), m).astype(np.int64))
            # this rank's share
            mine = np.flatnonzero(pid % world == rank)
            ev_local = np.flatnonzero(np.isin(mine, ev_pos))
            res = emul.EmulResult(
                k, mine.astype(np.int64), ev_local.astype(np.int64),
                pid[mine][ev_local], ev_ang[mine][ev_local],
                np.searchsorted(mine, starts))
            trk._step = k + 2                       # snapshot k+1 submitted
            h = comm.start_merge(trk, res, to_host=mode)
            if pending is not None:
                finish(pending)
            pending = h
        finish(pending)
        assert [f[0] for f in finished] == list(range(len(counts)))
        # every rank went through the same capacities ...
        all_caps = [None] * world
        dist.all_gather_object(all_caps, caps)
        assert all(c == all_caps[0] for c in all_caps)
        # ... and the slices of the ranks tile the global lists
        all_fin = [None] * world
        dist.all_gather_object(all_fin, finished)
        if mode is not True:
            for k in range(len(counts)):
                edges = sorted((f[k][1], f[k][2]) for f in all_fin)
                assert edges[0][0] == 0 and edges[-1][1] == counts[k]
                assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        # capacity follows the data instead of doubling without bound
        assert caps[-1] <= 8 * comm.HEADROOM * max(counts) / world + 8192
        open(os.path.join(out_dir, 'ok_%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,mode', [(2, 'slice'), (4, 'slice'), (3, 'slice'),
                                        (3, True)])
def test_async_exchange_host_logic(world, mode, tmp_path):
    mp.spawn(_async_worker, args=(world, _free_port(), mode, str(tmp_path)),
             nprocs=world, join=True)
    assert all(os.path.exists(str(tmp_path / ('ok_%d' % r)))
               for r in range(world))


def test_catalogue_staging_ring_is_round_robin():
    """A staging buffer of the catalogue broadcast comes round again only
    CAT_RING broadcasts later (two broadcasts are in flight at a time)."""
    sys.path.insert(0, REPO)
    from nbody_orbit_analysis_b200 import sharded
    comm = sharded.Comm(2, 0, device=torch.device('cpu'))
    seq = [comm._cat_buffer(5, lambda: torch.zeros(35, dtype=torch.float64))
           .data_ptr() for _ in range(3 * comm.CAT_RING)]
    other = comm._cat_buffer(9, lambda: torch.zeros(63, dtype=torch.float64))
    assert other.data_ptr() not in seq
    for i in range(len(seq) - comm.CAT_RING + 1):
        assert len(set(seq[i:i + comm.CAT_RING])) == comm.CAT_RING


def _batch_worker(rank, world, port, K, out_dir):
    """Several snapshots per exchange (Comm.stage_merge / finish_batch): the
    per-snapshot global lists must be those of the unbatched exchange."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['OA_EXCHANGE_BATCH'] = str(K)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nbody_orbit_analysis_b200 import sharded
        import exchange_emul as emul
        emul.install(sharded)
        comm = sharded.Comm(world, rank, device=torch.device('cpu'))
        assert comm.batch_size == K
        trk = emul.EmulTracker()
        n_prev = 50000
        counts = [0, 300, 4000, 3800, 20000, 35000, 90, 0, 25000, 5, 700, 1200, 33000]
        halos = [7, 7, 5, 9, 7, 7, 3, 7, 7, 1, 7, 8, 7]
        rng = np.random.default_rng(21)
        expected, finished = {}, []

        def check(results):
            for res in results:
                exp_ids, exp_ang, exp_off = expected.pop(res.step)
                assert res.n_events == len(exp_ids), (res.step, res.n_events)
                assert np.array_equal(res.apsis_offsets, exp_off)
                lo, hi = res.host_slice
                assert 0 <= lo <= hi <= res.n_events
                assert np.array_equal(res.apsis_ids, exp_ids[lo:hi])
                assert np.array_equal(res.apsis_angles.view(np.int16),
                                      exp_ang[lo:hi].view(np.int16))
                assert np.array_equal(res.d_ids.numpy(), exp_ids[lo:hi])
                finished.append((res.step, lo, hi, res.n_events))

        for k, (m, n_halos) in enumerate(zip(counts, halos)):
            starts = np.concatenate(([0], np.sort(rng.choice(
                np.arange(1, n_prev), n_halos - 1, replace=False)))) \
                if n_halos > 1 else np.array([0])
            pid = rng.permutation(n_prev).astype(np.int64) + 10 ** 12
            ev_pos = np.sort(rng.choice(n_prev, m, replace=False))
            ev_ang = rng.standard_normal(n_prev).astype(np.float16)
            expected[k] = (pid[ev_pos], ev_ang[ev_pos], np.append(
                np.searchsorted(ev_pos, starts), m).astype(np.int64))
            mine = np.flatnonzero(pid % world == rank)
            ev_local = np.flatnonzero(np.isin(mine, ev_pos))
            res = emul.EmulResult(
                k, mine.astype(np.int64), ev_local.astype(np.int64),
                pid[mine][ev_local], ev_ang[mine][ev_local],
                np.searchsorted(mine, starts))
            trk._step = k + 2
            comm.stage_merge(trk, res)
            while len(comm._launched) > 1:           # lag one batch behind
                check(comm.finish_batch(comm._launched[0]))
        comm.launch_batch(trk)                         # the partial last batch
        while comm._launched:
            check(comm.finish_batch(comm._launched[0]))
        assert [f[0] for f in finished] == list(range(len(counts)))
        all_fin = [None] * world
        dist.all_gather_object(all_fin, finished)
        for k in range(len(counts)):                   # slices tile every list
            edges = sorted((f[k][1], f[k][2]) for f in all_fin)
            covered = 0
            for a, z in edges:
                assert a == covered or a == z
                covered = max(covered, z)
            assert covered == counts[k]
        open(os.path.join(out_dir, 'ok_%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,K', [(2, 4), (3, 5)])
def test_batched_exchange(world, K, tmp_path):
    mp.spawn(_batch_worker, args=(world, _free_port(), K, str(tmp_path)),
             nprocs=world, join=True)
    assert all(os.path.exists(str(tmp_path / ('ok_%d' % r)))
               for r in range(world))
