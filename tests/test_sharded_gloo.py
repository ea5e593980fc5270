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
