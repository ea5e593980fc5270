"""The HOST side of ``OrbitTracker`` / ``track_orbits`` in the CPU container
(tests/fake_cuda.py): CUDA streams and events are no-ops on CPU tensors, the
partitioned-join kernel is the g++ build of its own stage code, the ordered
selection is numpy.  ``track_orbits(..., impl pjoin)`` must then write the same
file as the oracle (reference ``track_orbits.py:9-244``); the hash-table branch
(whose kernel has no host twin) must at least run through its host code.
"""
import os

import numpy as np
import pytest

import fake_cuda
from test_pjoin_emul import emul          # noqa: F401  (fixture)
from test_gpu_track import compare_track_trees
from nbody_orbit_analysis_b200 import pjoin, storage
from nbody_orbit_analysis_b200.synth import SynthSim
from oracle import orbit_oracle as oracle


@pytest.fixture
def pjoin_env(monkeypatch):
    monkeypatch.setenv('OA_TRACK_IMPL', 'pjoin')


CASES = [
    (30000, 9, 5, {}),
    (30000, 9, 5, {'late_halos': 0.4}),
    (20000, 40, 4, {'hubble': True}),
    (20000, 3, 4, {'periodic': False, 'nfw': True}),
    (15000, 4, 4, {'catalogue_bulk': False}),
]


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
@pytest.mark.parametrize('case', CASES, ids=[
    'plain', 'late_halos', 'hubble', 'nfw_nonperiodic', 'derived_bulk'])
def test_track_orbits_pjoin_matches_oracle(emul, pjoin_env, monkeypatch, case,
                                           mode, tmp_path):
    n, nh, ns, kw = case
    derived = not kw.get('catalogue_bulk', True)
    # small partitions, so that the partitioned stages run at these sizes
    monkeypatch.setattr(pjoin, 'TARGET', 400)
    monkeypatch.setattr(pjoin, 'LAG_PARTICLES', 1 << 12)
    sim = SynthSim(n, nh, ns, dtype=np.float32, catalogue_dtype=np.float32, **kw)
    f_dev, f_cpu = str(tmp_path / 'dev.h5'), str(tmp_path / 'cpu.h5')
    args = (sim.snapshot_numbers, sim.main_branches, sim.regions,
            sim.load_snapshot_data)
    from nbody_orbit_analysis_b200 import track_orbits
    with fake_cuda.install(emul) as fake:
        track_orbits.track_orbits(*args, f_dev, mode=mode, verbose=False,
                                  device='cpu')
        assert fake.calls.count('oa_pjoin_step') == ns
        assert 'oa_track_fused' not in fake.calls
    oracle.track_orbits(*args, f_cpu, mode=mode, storage=storage)
    got, exp = storage.tree(f_dev), storage.tree(f_cpu)
    assert sum(len(v) for k, v in exp.items() if k.endswith('er_IDs')) > 0
    compare_track_trees(got, exp, data_f64=False, derived_bulk=derived)


def test_pjoin_rejects_what_it_does_not_cover(emul, pjoin_env):
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    from nbody_orbit_analysis_b200._lib import OrbitB200Error
    sim = SynthSim(5000, 3, 2, dtype=np.float32, catalogue_dtype=np.float64)
    pos, rad, bulk = sim.regions(sim.snapshot_numbers[0], sim.main_branches[0])
    snap = sim.load_snapshot_data(sim.snapshot_numbers[0], pos, rad)
    with fake_cuda.install(emul):
        with pytest.raises(ValueError):
            OrbitTracker(impl='pjoin', onthefly=True, device='cpu')
        trk = OrbitTracker(device='cpu')
        assert trk.impl == 'pjoin'
        with pytest.raises(OrbitB200Error):        # float64 catalogue
            trk.step(snap, np.arange(3), pos, bulk, 0.0)
        trk = OrbitTracker(device='cpu')
        with pytest.raises(OrbitB200Error):        # checkpoint angles
            trk.step(snap, np.arange(3), pos.astype(np.float32),
                     bulk.astype(np.float32), 0.0, want_angles=True)


def test_hash_branch_host_code_runs(emul, tmp_path):
    """submit / collect of the default implementation with no-op kernels: the
    host code (buffers, region table, packed copies, launch arguments, result
    assembly) runs through; numbers are meaningless."""
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    sim = SynthSim(8000, 5, 3, dtype=np.float32, catalogue_dtype=np.float32)
    with fake_cuda.install(emul) as fake:
        trk = OrbitTracker(device='cpu')
        assert trk.impl == 'hash'
        pend = []
        for t, sn in enumerate(sim.snapshot_numbers):
            pos, rad, bulk = sim.regions(sn, sim.main_branches[t])
            snap = sim.load_snapshot_data(sn, pos, rad)
            pend.append(trk.submit(snap, np.arange(5), pos, bulk, 0.0,
                                   want_angles=(t == 1), diagnostics=(t == 2)))
            res = trk.collect(pend[-1])
            assert res.n == len(snap['ids'])
            if t > 0:
                assert res.n_events == 0 and len(res.apsis_offsets) == 6
        assert fake.calls.count('oa_track_fused') == 3
        assert os.environ.get('OA_TRACK_IMPL') is None


class HostSynth:
    """Host twin of ``synth.DeviceSynth`` (same interface, SynthSim data)."""

    def __init__(self, n_particles, n_halos, rank=0, world=1, **kw):
        import torch
        self._torch = torch
        self.host = SynthSim(n_particles, n_halos, 64, dtype=np.float32,
                             catalogue_dtype=np.float32)
        self.n_halos = n_halos

    def regions(self, t):
        h = self.host
        return (h.halo_centre(t).astype(np.float32), h.radius.astype(np.float32),
                h.vh.astype(np.float32))

    def snapshot(self, t):
        torch = self._torch
        pos, rad, _ = self.regions(t)
        s = self.host.load_snapshot_data(self.host.snapshot_numbers[t], pos, rad,
                                         cols=np.arange(self.n_halos))
        n = len(s['ids'])
        dev = {'pos': torch.from_numpy(s['coordinates'].reshape(-1).copy()),
               'vel': torch.from_numpy(s['velocities'].reshape(-1).copy()),
               'ids': torch.from_numpy(s['ids'].copy()), 'mass': None,
               'gpos': None}
        return dev, n, np.append(s['region_offsets'], n).astype(np.int64)


def test_bench_end_to_end_on_fake_cuda(emul, pjoin_env, monkeypatch, capsys):
    """``bench.py``'s GPU arm from argument parsing to the JSON line, on the CPU:
    device-resident run, end-to-end run from host arrays, roofline block and the
    CPU baseline whose sample must agree with the (emulated) GPU path."""
    import argparse
    import json
    import bench
    from nbody_orbit_analysis_b200 import synth
    monkeypatch.setattr(synth, 'DeviceSynth', HostSynth)
    monkeypatch.setattr(pjoin, 'TARGET', 400)
    monkeypatch.setattr(pjoin, 'LAG_PARTICLES', 1 << 12)
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    monkeypatch.setenv('OA_BENCH_CLOCK_PERIOD', '0.05')
    args = argparse.Namespace(
        gpus=1, steps=3, warmup=3, impl='b200', particles=10000, halos=8,
        mode='pericentric', depth=2, profile=False, no_e2e=False, no_cpu=False,
        cpu_particles=3000)
    with fake_cuda.install(emul):
        bench.run_b200(args)
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines()
                       if ln.startswith('{')][-1])
    assert line['metric'] == 'particle-snapshots/sec' and line['value'] > 0
    assert line['n_gpus'] == 1 and line['steps'] == 3 and line['warmup'] == 3
    assert line['track_impl'] == 'pjoin' and line['gpu_launches'] > 0
    assert line['events_per_step'] > 0
    assert line['e2e']['value'] > 0 and line['e2e']['h2d_bytes_per_step'] > 0
    assert line['e2e']['events_per_step'] == line['events_per_step']
    # `value`: results left in HBM; the pass that also copies them to the host
    # is reported beside it and must see the same events
    assert 'HBM' in line['config']['results']
    vh = line['value_results_to_host']
    assert vh['value'] > 0 and vh['events_equal_device_run'] is True
    assert line['parity'] == 'ok'
    # the same data through the drop-in entry point (pageable arrays, file write)
    ep = line['e2e_entry_point']
    assert ep['value'] > 0 and ep['snapshots'] == 3 and ep['file_bytes_per_step'] > 0
    r = line['roofline']
    assert r['bound'] == 'hbm' and r['achieved'] > 0 and 'oa_pjoin' in r['kernel']
    assert line['cpu_baseline']['parity_vs_gpu_on_sample'] == 'ok'
    assert line['cpu_baseline']['sample_events'] > 0
    assert set(line['host_phases_ms_per_step']) >= {'submit', 'collect'}


def test_bench_default_implementation_on_the_numpy_twin(monkeypatch, capsys):
    """``bench.py`` with the DEFAULT implementation, its kernel replaced by the
    numpy twin (tests/hash_twin.py): the three passes (results left in HBM,
    results copied to the host, end to end from host arrays) see the same events,
    and the CPU-baseline sample agrees with the oracle."""
    import argparse
    import json
    import bench
    import hash_twin
    from nbody_orbit_analysis_b200 import synth
    monkeypatch.setattr(synth, 'DeviceSynth', HostSynth)
    for k in ('WORLD_SIZE', 'OA_TRACK_IMPL', 'OA_BENCH_TO_HOST'):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv('OA_BENCH_CLOCK_PERIOD', '0.05')
    args = argparse.Namespace(
        gpus=1, steps=3, warmup=3, impl='b200', particles=8000, halos=6,
        mode='pericentric', depth=2, profile=False, no_e2e=False, no_cpu=False,
        cpu_particles=2000)
    with fake_cuda.install(None, hash_twin.TwinLib):
        bench.run_b200(args)
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines()
                       if ln.startswith('{')][-1])
    assert line['track_impl'] == 'hash' and line['parity'] == 'ok'
    assert line['events_per_step'] > 100
    assert line['value_results_to_host']['events_equal_device_run'] is True
    assert line['e2e']['events_equal_device_run'] is True
    assert line['e2e']['events_per_step'] == line['events_per_step']
    cpu = line['cpu_baseline']
    assert cpu['parity_vs_gpu_on_sample'] == 'ok' and cpu['sample_events'] > 0
    assert 'error' not in line['e2e_entry_point']


def test_bench_hash_branch_runs_on_fake_cuda(emul, monkeypatch, capsys):
    """The default implementation through ``bench.py`` with no-op kernels:
    every line of the GPU arm executes (numbers are meaningless)."""
    import argparse
    import json
    import bench
    from nbody_orbit_analysis_b200 import synth
    monkeypatch.setattr(synth, 'DeviceSynth', HostSynth)
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    monkeypatch.delenv('OA_TRACK_IMPL', raising=False)
    monkeypatch.setenv('OA_BENCH_CLOCK_PERIOD', '0.05')
    args = argparse.Namespace(
        gpus=1, steps=3, warmup=3, impl='b200', particles=8000, halos=6,
        mode='pericentric', depth=2, profile=False, no_e2e=False, no_cpu=False,
        cpu_particles=2000)
    with fake_cuda.install(emul) as fake:
        bench.run_b200(args)
        assert fake.calls.count('oa_track_fused') >= 2 * 7 + 7
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines()
                       if ln.startswith('{')][-1])
    assert line['track_impl'] == 'hash' and 'oa_track_kernel' in line['roofline']['kernel']
    assert line['cpu_baseline']['parity_vs_gpu_on_sample'] in ('ok', 'MISMATCH')
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup',
                'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'clocks', 'gpu_launches', 'e2e',
                'roofline', 'cpu_baseline'):
        assert key in line, key


# ---------------------------------------------------------------------------
# bench.py, multi-rank, on the CPU: gloo + emulated kernels
# ---------------------------------------------------------------------------
class ShardedHostSynth(HostSynth):
    """Every rank generates the same universe and keeps ``id mod world ==
    rank`` (what DeviceSynth does on the device); ``gpos`` = position in the
    unsharded snapshot."""

    def __init__(self, n_particles, n_halos, rank=0, world=1, **kw):
        HostSynth.__init__(self, n_particles * world, n_halos)
        self.rank, self.world = rank, world

    def snapshot(self, t):
        from nbody_orbit_analysis_b200 import sharded
        torch = self._torch
        pos, rad, _ = self.regions(t)
        full = self.host.load_snapshot_data(
            self.host.snapshot_numbers[t], pos, rad, cols=np.arange(self.n_halos))
        loc, gpos = sharded.shard_snapshot(full, self.rank, self.world)
        n = len(loc['ids'])
        dev = {'pos': torch.from_numpy(loc['coordinates'].reshape(-1).copy()),
               'vel': torch.from_numpy(loc['velocities'].reshape(-1).copy()),
               'ids': torch.from_numpy(loc['ids'].copy()), 'mass': None,
               'gpos': torch.from_numpy(gpos.copy())}
        return dev, n, np.append(loc['region_offsets'], n).astype(np.int64)


def _bench_rank(rank, world, port, emul_path, out_dir, batch=1, prepack=False,
                impl='pjoin'):
    import argparse
    import contextlib
    import io
    import sys
    import ctypes as C
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), OA_TRACK_IMPL=impl,
                      OA_BENCH_CLOCK_PERIOD='0.05', OA_FAKE_CTAS='1',
                      OA_EXCHANGE_BATCH=str(batch),
                      OA_EXCHANGE_PREPACK='1' if prepack else '0')
    if batch > 1 or prepack:
        # (one device-resident pass is enough for the variants: no second pass
        # with the results copied to the host)
        os.environ['OA_BENCH_TO_HOST'] = '0'
    else:
        os.environ.pop('OA_BENCH_TO_HOST', None)
    import bench
    import exchange_emul
    import fake_cuda as fc
    from nbody_orbit_analysis_b200 import pjoin as pj, sharded, synth
    emul_lib = C.CDLL(emul_path)
    emul_lib.pj_emul_step.argtypes = [C.POINTER(pj.PJoinArgs), C.c_int]
    synth.DeviceSynth = ShardedHostSynth
    pj.TARGET, pj.LAG_PARTICLES = 400, 1 << 12
    exchange_emul.install(sharded)
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init(
        'gloo', rank=rank, world_size=world)
    args = argparse.Namespace(
        gpus=world, steps=3, warmup=3, impl='b200', particles=6000, halos=7,
        mode='pericentric', depth=2, profile=False,
        no_e2e=world > 2 or batch > 1 or prepack,   # (variants: one pass each)
        no_cpu=batch > 1 or prepack, cpu_particles=2000)
    buf = io.StringIO()
    lib_class = None
    if impl == 'hash':         # default implementation: numpy twin of its kernel
        import hash_twin
        lib_class = hash_twin.TwinLib
    with fc.install(emul_lib, lib_class), contextlib.redirect_stdout(buf):
        bench.run_b200(args)
    with open(os.path.join(out_dir, 'out_%d' % rank), 'w') as fh:
        fh.write(buf.getvalue())


@pytest.mark.parametrize('world,batch,prepack,impl', [
    (2, 1, False, 'pjoin'), (2, 3, False, 'pjoin'), (2, 1, True, 'pjoin'),
    (2, 1, False, 'hash')])
def test_bench_multi_rank_on_fake_cuda(emul, world, batch, prepack, impl,
                                       tmp_path, monkeypatch, capsys):
    """The multi-GPU arm of ``bench.py`` with `world` ranks on the CPU: sharded
    snapshots, catalogue broadcast one snapshot ahead, asynchronous all-to-all
    exchange (numpy restatements of its kernels, gloo) and per-rank slices --
    and the same global event count as ONE rank tracking the whole universe
    (also with three snapshots per exchange, OA_EXCHANGE_BATCH=3)."""
    import json
    import torch.multiprocessing as mp
    from test_sharded_gloo import _free_port
    mp.spawn(_bench_rank, args=(world, _free_port(), emul._name, str(tmp_path),
                                batch, prepack, impl), nprocs=world, join=True)
    lines = [ln for ln in open(str(tmp_path / 'out_0')).read().splitlines()
             if ln.startswith('{')]
    multi = json.loads(lines[-1])
    assert multi['n_gpus'] == world and multi['value'] > 0
    assert multi['events_per_step'] > 0
    assert multi['parity'] == 'ok'
    if batch == 1 and not prepack:
        vh = multi['value_results_to_host']
        assert vh['value'] > 0 and vh['events_equal_device_run'] is True
    else:
        assert 'value_results_to_host' not in multi
    base = batch == 1 and not prepack
    if base:
        # a sample of whole halos, put back together from all ranks' shards,
        # through both exchange paths against the oracle on rank 0
        par = multi['parity_multi_gpu']
        assert par['parity_vs_oracle'] == 'ok' and par['sample_events'] > 0
        assert par['all_gather'] == par['all_to_all'] == 'ok'
        assert par['batched_all_to_all'] == 'ok'
    if world == 2 and base:
        assert multi['e2e']['value'] > 0
        assert multi['e2e']['events_per_step'] == multi['events_per_step']
    assert {'catalogue', 'submit', 'collect', 'start_merge'} | (
        {'finish_merge'} if batch == 1 else set()) <= set(
            multi['host_phases_ms_per_step'])
    for r in range(1, world):          # only rank 0 prints
        assert '{' not in open(str(tmp_path / ('out_%d' % r))).read()

    if prepack:
        # (OA_EXCHANGE_PREPACK=1: the pack kernels ran at submit time)
        assert multi['exchange_prepacked'] >= 2
    if world != 2:
        return
    # one rank, whole universe
    import argparse
    import bench
    from nbody_orbit_analysis_b200 import synth

    class Whole(HostSynth):
        def __init__(self, n_particles, n_halos, **kw):
            HostSynth.__init__(self, n_particles * world, n_halos)
    monkeypatch.setattr(synth, 'DeviceSynth', Whole)
    monkeypatch.setattr(pjoin, 'TARGET', 400)
    monkeypatch.setattr(pjoin, 'LAG_PARTICLES', 1 << 12)
    monkeypatch.setenv('OA_TRACK_IMPL', 'pjoin')
    monkeypatch.setenv('OA_BENCH_CLOCK_PERIOD', '0.05')
    monkeypatch.setenv('OA_FAKE_CTAS', '1')
    for k in ('WORLD_SIZE', 'RANK', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(
        gpus=1, steps=3, warmup=3, impl='b200', particles=6000, halos=7,
        mode='pericentric', depth=2, profile=False, no_e2e=True, no_cpu=True,
        cpu_particles=2000)
    with fake_cuda.install(emul):
        bench.run_b200(args)
    single = json.loads([ln for ln in capsys.readouterr().out.splitlines()
                         if ln.startswith('{')][-1])
    assert single['events_per_step'] == multi['events_per_step']


@pytest.mark.parametrize('dt,cdt', [(np.float32, np.float32),
                                    (np.float64, np.float64),
                                    (np.float32, np.float64)])
def test_onthefly_and_f64_branches_host_code_runs(emul, dt, cdt):
    """Host code of the remaining ``submit`` variants with no-op kernels:
    on-the-fly (derived bulk velocity, per-match angle buffer), float64 frames,
    mass arrays, vanishing halos."""
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    sim = SynthSim(6000, 5, 3, dtype=dt, catalogue_dtype=cdt, mass_array=True)
    with fake_cuda.install(emul) as fake:
        for onthefly in (True, False):
            trk = OrbitTracker(device='cpu', onthefly=onthefly, impl='hash')
            for t, sn in enumerate(sim.snapshot_numbers):
                exists = np.arange(5) if t != 1 else np.array([0, 2, 3])
                pos, rad, bulk = sim.regions(sn, sim.main_branches[t][exists])
                snap = sim.load_snapshot_data(sn, pos, rad)
                res = trk.step(snap, exists, pos, None if onthefly else bulk,
                               0.0 if onthefly else 0.07, diagnostics=onthefly)
                assert res.n == len(snap['ids'])
                assert list(res.hinds) == ([] if t == 0 else
                                           list(range(len(exists))) if t == 1
                                           else [0, 2, 3])
        assert fake.calls.count('oa_bulk_velocity') == 3


def test_pjoin_empty_block_and_vanishing_halo(emul, pjoin_env, monkeypatch, tmp_path):
    """A halo disappears for one snapshot and comes back; a (large, partitioned)
    block becomes empty from the second snapshot on -- partition counts never
    shrink, so an empty block keeps its partitions (all of size 0)."""
    from nbody_orbit_analysis_b200 import track_orbits
    monkeypatch.setattr(pjoin, 'TARGET', 300)
    monkeypatch.setattr(pjoin, 'LAG_PARTICLES', 1 << 12)
    sim = SynthSim(20000, 6, 6, dtype=np.float32, catalogue_dtype=np.float32)
    mb = sim.main_branches.copy()
    mb[2, 4] = -1          # halo 4 vanishes at t=2 and returns at t=3
    base_load = sim.load_snapshot_data
    calls = {'n': 0}

    def load(snap_no, pos, rad):
        s = base_load(snap_no, pos, rad)
        calls['n'] += 1
        offs = np.append(s['region_offsets'], len(s['ids']))
        if len(offs) > 3 and snap_no != sim.snapshot_numbers[0]:
            lo, hi = offs[1], offs[2]              # empty the second block
            keep = np.ones(len(s['ids']), dtype=bool)
            keep[lo:hi] = False
            for k in ('ids', 'coordinates', 'velocities'):
                s[k] = s[k][keep]
            offs[2:] -= hi - lo
            s['region_offsets'] = offs[:-1]
        return s

    f_dev, f_cpu = str(tmp_path / 'd.h5'), str(tmp_path / 'c.h5')
    with fake_cuda.install(emul):
        track_orbits.track_orbits(sim.snapshot_numbers, mb, sim.regions, load,
                                  f_dev, verbose=False, device='cpu')
    oracle.track_orbits(sim.snapshot_numbers, mb, sim.regions, load, f_cpu,
                        storage=storage)
    compare_track_trees(storage.tree(f_dev), storage.tree(f_cpu), data_f64=False)


# ---------------------------------------------------------------------------
# the drop-in entry point, sharded over 2 ranks (gloo), against the oracle
# ---------------------------------------------------------------------------
def _entry_rank(rank, world, port, emul_path, out_dir, loader_side,
                impl='pjoin'):
    import ctypes as C
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      OA_TRACK_IMPL=impl, OA_FAKE_CTAS='1')
    import exchange_emul
    import fake_cuda as fc
    from nbody_orbit_analysis_b200 import pjoin as pj, sharded, track_orbits
    emul_lib, lib_class = None, None
    if impl == 'pjoin':
        emul_lib = C.CDLL(emul_path)
        emul_lib.pj_emul_step.argtypes = [C.POINTER(pj.PJoinArgs), C.c_int]
    else:                      # default implementation: numpy twin of its kernel
        import hash_twin
        lib_class = hash_twin.TwinLib
    pj.TARGET, pj.LAG_PARTICLES = 400, 1 << 12
    exchange_emul.install(sharded)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sim = SynthSim(24000, 7, 5, dtype=np.float32, catalogue_dtype=np.float32,
                   late_halos=0.3)
    calls = {'regions': 0}

    def regions(sn, halo_ids):
        calls['regions'] += 1
        return sim.regions(sn, halo_ids)

    def loader(sn, pos, rad):
        snap = sim.load_snapshot_data(sn, pos, rad)
        if loader_side:                 # this rank's shard + global positions
            snap, gpos = sharded.shard_snapshot(snap, rank, world)
            snap['_gpos'] = gpos
        return snap
    out = os.path.join(out_dir, 'sharded.h5')
    with fc.install(emul_lib, lib_class):
        track_orbits.track_orbits(sim.snapshot_numbers, sim.main_branches,
                                  regions, loader, out, verbose=False,
                                  device='cpu')
    # the catalogue is read on rank 0 only and broadcast
    assert (calls['regions'] > 0) == (rank == 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('loader_side,impl', [(False, 'pjoin'), (True, 'pjoin'),
                                              (True, 'hash')])
def test_track_orbits_sharded_entry_point(emul, tmp_path, loader_side, impl):
    """``track_orbits`` under a 2-rank process group: catalogue broadcast from
    rank 0, particles sharded by ID (by the driver, or by the loader with
    ``_gpos``), events merged in the reference order, rank 0 writes the file --
    which must equal the oracle's single-process file."""
    import torch.multiprocessing as mp
    from test_sharded_gloo import _free_port
    mp.spawn(_entry_rank, args=(2, _free_port(), emul._name, str(tmp_path),
                                loader_side, impl), nprocs=2, join=True)
    sim = SynthSim(24000, 7, 5, dtype=np.float32, catalogue_dtype=np.float32,
                   late_halos=0.3)
    f_cpu = str(tmp_path / 'cpu.h5')
    oracle.track_orbits(sim.snapshot_numbers, sim.main_branches, sim.regions,
                        sim.load_snapshot_data, f_cpu, storage=storage)
    got, exp = storage.tree(str(tmp_path / 'sharded.h5')), storage.tree(f_cpu)
    assert sum(len(v) for k, v in exp.items() if k.endswith('er_IDs')) > 0
    compare_track_trees(got, exp, data_f64=False, derived_bulk=False)


def test_loader_error_leaves_whole_groups(emul, pjoin_env, monkeypatch, tmp_path):
    """The result groups are written by a background thread: when a callback
    raises, the caller sees that error and the file holds exactly the groups
    of the snapshots finished before it, complete and equal to the oracle's."""
    from nbody_orbit_analysis_b200 import track_orbits
    monkeypatch.setattr(pjoin, 'TARGET', 400)
    monkeypatch.setattr(pjoin, 'LAG_PARTICLES', 1 << 12)
    sim = SynthSim(20000, 6, 6, dtype=np.float32, catalogue_dtype=np.float32)

    def load(snap_no, pos, rad):
        if snap_no == sim.snapshot_numbers[4]:
            raise RuntimeError('snapshot file missing')
        return sim.load_snapshot_data(snap_no, pos, rad)
    f_dev, f_cpu = str(tmp_path / 'd.h5'), str(tmp_path / 'c.h5')
    with fake_cuda.install(emul):
        with pytest.raises(RuntimeError, match='snapshot file missing'):
            track_orbits.track_orbits(sim.snapshot_numbers, sim.main_branches,
                                      sim.regions, load, f_dev, verbose=False,
                                      device='cpu')
    oracle.track_orbits(sim.snapshot_numbers, sim.main_branches, sim.regions,
                        sim.load_snapshot_data, f_cpu, storage=storage)
    got, exp = storage.tree(f_dev), storage.tree(f_cpu)
    def group(k):
        return k.strip('/').split('/')[0]
    groups = sorted({group(k) for k in got if group(k).startswith('snapshot_')})
    assert groups == ['snapshot_%03d' % sim.snapshot_numbers[t] for t in (1, 2)]
    exp = {k: v for k, v in exp.items()
           if not group(k).startswith('snapshot_') or group(k) in groups}
    compare_track_trees(got, exp, data_f64=False)


# ---------------------------------------------------------------------------
# property test: the drop-in driver (host code + emulated kernel) vs the oracle
# ---------------------------------------------------------------------------
def _hyp():
    from hypothesis import HealthCheck, given, settings, strategies as st
    cfg = st.fixed_dictionaries(dict(
        n_particles=st.integers(500, 6000), n_halos=st.integers(1, 12),
        n_snap=st.integers(2, 6), seed=st.integers(1, 2 ** 31),
        nfw=st.booleans(), hubble=st.booleans(), catalogue_bulk=st.booleans(),
        mass_array=st.booleans(), periodic=st.booleans(),
        late_halos=st.sampled_from([0.0, 0.0, 0.4])))
    return HealthCheck, given, settings, st, cfg


_HC, _given, _settings, _st, _cfg = _hyp()


@_settings(max_examples=30, deadline=None, derandomize=True, database=None,
           suppress_health_check=list(_HC))
@_given(kw=_cfg, mode=_st.sampled_from(['pericentric', 'apocentric']),
        reverse=_st.booleans(), vanish=_st.integers(0, 40),
        target=_st.sampled_from([64, 400, 4096]))
def test_track_orbits_property_vs_oracle(emul, tmp_path_factory, kw, mode,
                                         reverse, vanish, target):
    """Drawn configurations through ``track_orbits`` (pipelined driver, region
    table, partition plan, staging, writer thread; kernel = the g++ build of the
    partitioned join) against the oracle: same file.  Includes input rows in
    reverse order, halos that vanish for a snapshot, derived and mass-weighted
    bulk velocities, open boxes."""
    from nbody_orbit_analysis_b200 import track_orbits
    tmp = tmp_path_factory.mktemp('prop')
    sim = SynthSim(dtype=np.float32, catalogue_dtype=np.float32, **kw)
    snaps, mb = sim.snapshot_numbers.copy(), sim.main_branches.copy()
    if vanish and sim.n_snap > 2:        # one halo missing at one inner snapshot
        mb[1 + vanish % (sim.n_snap - 2), vanish % sim.n_halos] = -1
    if reverse:
        snaps, mb = snaps[::-1].copy(), mb[::-1].copy()
    f_dev, f_cpu = str(tmp / 'dev.h5'), str(tmp / 'cpu.h5')
    saved = (pjoin.TARGET, pjoin.LAG_PARTICLES, os.environ.get('OA_TRACK_IMPL'))
    pjoin.TARGET, pjoin.LAG_PARTICLES = target, 1 << 12
    os.environ['OA_TRACK_IMPL'] = 'pjoin'
    try:
        with np.errstate(all='ignore'):
            try:
                oracle.track_orbits(snaps, mb, sim.regions,
                                    sim.load_snapshot_data, f_cpu, mode=mode,
                                    storage=storage)
                failed = False
            except ValueError:           # a snapshot without matched halos
                failed = True
            with fake_cuda.install(emul):
                if failed:
                    with pytest.raises(ValueError):
                        track_orbits.track_orbits(
                            snaps, mb, sim.regions, sim.load_snapshot_data,
                            f_dev, mode=mode, verbose=False, device='cpu')
                    return
                track_orbits.track_orbits(snaps, mb, sim.regions,
                                          sim.load_snapshot_data, f_dev,
                                          mode=mode, verbose=False, device='cpu')
    finally:
        pjoin.TARGET, pjoin.LAG_PARTICLES = saved[:2]
        if saved[2] is None:
            os.environ.pop('OA_TRACK_IMPL', None)
        else:
            os.environ['OA_TRACK_IMPL'] = saved[2]
    compare_track_trees(storage.tree(f_dev), storage.tree(f_cpu), data_f64=False,
                        derived_bulk=not kw['catalogue_bulk'])
