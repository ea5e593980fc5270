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
            # marks are whatever the buffers hold: make them "no event"
            pend.append(trk.submit(snap, np.arange(5), pos, bulk, 0.0,
                                   want_angles=(t == 1), diagnostics=(t == 2)))
            trk.prev.mark.fill_(-32768)
            res = trk.collect(pend[-1])
            assert res.n == len(snap['ids'])
            if t > 0:
                assert res.n_events == 0 and len(res.apsis_offsets) == 6
        assert fake.calls.count('oa_track_fused') == 3
        assert os.environ.get('OA_TRACK_IMPL') is None
