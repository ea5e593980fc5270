"""The DEFAULT implementation's host code on the CPU: the driver-level tests of
the GPU suite (``tests/test_gpu_track.py``) executed under ``tests/fake_cuda.py``
with the fused hash-table kernel replaced by its numpy twin
(``tests/hash_twin.py``: the oracle's per-region functions applied to the
buffers of the launch arguments).

What this checks without a GPU: region table and launch arguments of
``oa_track_fused``, record / mark / ring-buffer bookkeeping, the ordered event
selection and result assembly, checkpoints of the angle accumulators, resume,
device-state checkpoints, the two-slot ring, float64 and mixed-dtype frames,
derived and mass-weighted bulk velocities -- against the golden fixtures of the
unmodified reference and against the oracle, with the SAME assertions the
``-m gpu`` tests make.  What it does not check is the CUDA kernel.
"""
import numpy as np
import pytest

import fake_cuda
import hash_twin
import test_gpu_track as gpu_tests
from fixture_io import list_fixtures


@pytest.fixture
def twin(monkeypatch):
    monkeypatch.delenv('OA_TRACK_IMPL', raising=False)
    with fake_cuda.install(None, hash_twin.TwinLib) as fake:
        yield fake


@pytest.mark.parametrize('name', list_fixtures('track_'))
def test_reference_fixtures_through_the_default_host_path(twin, name, tmp_path):
    gpu_tests.test_track_orbits_matches_reference_fixture(name, tmp_path)
    assert 'oa_track_fused' in twin.calls


@pytest.mark.parametrize('name', list_fixtures('kernels_'))
def test_reference_vectors_through_the_default_host_path(twin, name):
    """diagnostic outputs, ``load_angles``, per-particle match indices."""
    gpu_tests.test_frame_and_match_against_reference_vectors(name)
    assert 'oa_set_record_angles' in twin.calls


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
@pytest.mark.parametrize('case', [
    (12000, 9, 5, np.float32, np.float32, {}),
    (12000, 9, 5, np.float32, np.float64, {'late_halos': 0.3}),
    (8000, 5, 4, np.float64, np.float64, {'hubble': True}),
    (8000, 40, 4, np.float32, np.float32, {'hubble': True}),
    (6000, 4, 4, np.float32, np.float64, {'catalogue_bulk': False}),
    (6000, 4, 4, np.float64, np.float64,
     {'catalogue_bulk': False, 'mass_array': True}),
    (6000, 1, 4, np.float32, np.float32, {'nfw': True, 'periodic': False}),
    (1500, 300, 3, np.float32, np.float32, {}),
], ids=['f32c32', 'f32c64_late', 'f64_hubble', 'f32_hubble_40h', 'f32_nobulk',
        'f64_massarr', 'nfw_nonperiodic', 'tiny_blocks'])
def test_oracle_cases_through_the_default_host_path(twin, case, mode, tmp_path):
    gpu_tests.test_track_orbits_matches_oracle(case, mode, tmp_path)


def test_empty_block_and_vanishing_halo(twin, tmp_path):
    gpu_tests.test_empty_block_and_vanishing_halo(tmp_path)


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
def test_device_state_checkpoint_resume(twin, mode, tmp_path):
    """``checkpoint='state'``: crash inside the loader, resume at the next
    snapshot from the saved records."""
    gpu_tests.test_device_state_checkpoint_resume(mode, tmp_path)


def test_two_slot_ring(twin, tmp_path, monkeypatch):
    gpu_tests.test_two_slot_ring_memory_plan(tmp_path, monkeypatch)


def _property():
    from hypothesis import HealthCheck, given, settings, strategies as st
    cfg = st.fixed_dictionaries(dict(
        n_particles=st.integers(500, 5000), n_halos=st.integers(1, 12),
        n_snap=st.integers(2, 6), seed=st.integers(1, 2 ** 31),
        dtype=st.sampled_from([np.float32, np.float64]),
        catalogue_dtype=st.sampled_from([np.float32, np.float64]),
        nfw=st.booleans(), hubble=st.booleans(), catalogue_bulk=st.booleans(),
        mass_array=st.booleans(), periodic=st.booleans(),
        box_vector=st.booleans(),
        late_halos=st.sampled_from([0.0, 0.0, 0.4])))
    deco = settings(max_examples=40, deadline=None, derandomize=True,
                    database=None, suppress_health_check=list(HealthCheck))
    return deco, given, st, cfg


_deco, _given, _st, _cfg = _property()


@_deco
@_given(kw=_cfg, mode=_st.sampled_from(['pericentric', 'apocentric']),
        checkpoint=_st.booleans(), reverse=_st.booleans(),
        vanish=_st.integers(0, 40))
def test_property_default_host_path_vs_oracle(tmp_path_factory, kw, mode,
                                              checkpoint, reverse, vanish):
    """Drawn configurations (every dtype combination, open / periodic / (3,)
    box, Hubble flow, derived / mass-weighted bulk velocity, halos that appear
    late or vanish for a snapshot, reversed input rows, checkpoints) through
    ``track_orbits`` with the default implementation's host code."""
    import os
    from nbody_orbit_analysis_b200 import storage, track_orbits
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    from parity import f16_ulps
    tmp = tmp_path_factory.mktemp('twin_prop')
    kw = dict(kw, box_vector=kw['box_vector'] and kw['periodic'])
    sim = SynthSim(**kw)
    snaps, mb = sim.snapshot_numbers.copy(), sim.main_branches.copy()
    if vanish and sim.n_snap > 2:
        mb[1 + vanish % (sim.n_snap - 2), vanish % sim.n_halos] = -1
    if reverse:
        snaps, mb = snaps[::-1].copy(), mb[::-1].copy()
    f_dev, f_cpu = str(tmp / 'dev.h5'), str(tmp / 'cpu.h5')
    saved = os.environ.pop('OA_TRACK_IMPL', None)
    try:
        with np.errstate(all='ignore'):
            try:
                oracle.track_orbits(snaps, mb, sim.regions,
                                    sim.load_snapshot_data, f_cpu, mode=mode,
                                    checkpoint=checkpoint, storage=storage)
                failed = False
            except ValueError:
                failed = True
            with fake_cuda.install(None, hash_twin.TwinLib):
                if failed:
                    with pytest.raises(ValueError):
                        track_orbits.track_orbits(
                            snaps, mb, sim.regions, sim.load_snapshot_data,
                            f_dev, mode=mode, checkpoint=checkpoint,
                            verbose=False)
                    return
                track_orbits.track_orbits(
                    snaps, mb, sim.regions, sim.load_snapshot_data, f_dev,
                    mode=mode, checkpoint=checkpoint, verbose=False)
    finally:
        if saved is not None:
            os.environ['OA_TRACK_IMPL'] = saved
    gpu_tests.compare_track_trees(
        storage.tree(f_dev), storage.tree(f_cpu),
        data_f64=kw['dtype'] == np.float64,
        derived_bulk=not kw['catalogue_bulk'])
    if checkpoint:
        a = storage.tree(f_dev + '.checkpoint')['/angles']
        b = storage.tree(f_cpu + '.checkpoint')['/angles']
        assert a.shape == b.shape and f16_ulps(a, b).max(initial=0) == 0


@pytest.fixture
def twin_onthefly(twin, monkeypatch):
    """The on-the-fly driver holds its own reference to the library."""
    from nbody_orbit_analysis_b200 import track_orbits_onthefly
    monkeypatch.setattr(track_orbits_onthefly, 'lib', twin)
    return twin


@pytest.mark.parametrize('name', list_fixtures('onthefly_'))
def test_onthefly_reference_fixtures_through_the_host_path(twin_onthefly, name,
                                                           tmp_path):
    """a-12 on the CPU: the on-the-fly driver (two launches of the tracking
    step, ordered selections, sorted entered / departed lists, dtype rules of
    the reference's concatenations) against the reference's own files."""
    import test_gpu_onthefly as otf
    otf.test_onthefly_matches_reference_fixture(name, tmp_path)


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
@pytest.mark.parametrize('case', [
    (12000, 9, np.float32, np.float32, {}, (), ()),
    (12000, 9, np.float32, np.float64, {}, (2, 7), (5,)),
    (8000, 9, np.float64, np.float64, {'mass_array': True}, (0,), (8,)),
    (3000, 400, np.float32, np.float32, {}, (3, 4, 5), (10,)),
], ids=['f32', 'f32c64_missing', 'f64_massarr_missing', 'tiny_blocks'])
def test_onthefly_oracle_cases_through_the_host_path(twin_onthefly, case, mode,
                                                     tmp_path):
    import test_gpu_onthefly as otf
    otf.test_onthefly_matches_oracle(case, mode, tmp_path)


@_deco
@_given(kw=_cfg, mode=_st.sampled_from(['pericentric', 'apocentric']),
        drop=_st.lists(_st.tuples(_st.integers(0, 1), _st.integers(0, 11)),
                       max_size=3))
def test_property_onthefly_host_path_vs_oracle(tmp_path_factory, kw, mode, drop):
    """Drawn configurations through the on-the-fly entry point (halos missing
    at s or at s-1, every dtype combination, mass arrays, open boxes)."""
    import os
    import test_gpu_onthefly as otf
    from nbody_orbit_analysis_b200 import storage, track_orbits_onthefly
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    tmp = tmp_path_factory.mktemp('twin_otf')
    kw = dict(kw, late_halos=0.0, hubble=False,
              box_vector=kw['box_vector'] and kw['periodic'])
    sim = SynthSim(**kw)
    t = sim.n_snap - 1
    links = np.stack([sim.main_branches[t], sim.main_branches[t - 1]])
    for row, col in drop:
        links[row, col % sim.n_halos] = -1
    if not (links[0] != -1).any() or not (links[1] != -1).any():
        return
    snap_no = int(sim.snapshot_numbers[t])
    f_dev, f_cpu = str(tmp / 'd_{}.h5'), str(tmp / 'c_{}.h5')
    saved = os.environ.pop('OA_TRACK_IMPL', None)
    real_lib = track_orbits_onthefly.lib
    try:
        with np.errstate(all='ignore'):
            oracle.track_orbits_onthefly(
                snap_no, links, sim.regions_onthefly, sim.load_snapshot_data,
                f_cpu, mode=mode, storage=storage)
            with fake_cuda.install(None, hash_twin.TwinLib) as fake:
                track_orbits_onthefly.lib = fake
                track_orbits_onthefly.track_orbits(
                    snap_no, links, sim.regions_onthefly, sim.load_snapshot_data,
                    f_dev, mode=mode, verbose=False)
    finally:
        track_orbits_onthefly.lib = real_lib
        if saved is not None:
            os.environ['OA_TRACK_IMPL'] = saved
    name = '%0.3d' % snap_no
    otf.compare_onthefly_trees(storage.tree(f_dev.format(name)),
                               storage.tree(f_cpu.format(name)),
                               kw['dtype'] == np.float64)
