"""GPU parity tests of the two consumers of the tracking path (``-m gpu``):
``progenitors.py`` and ``postprocessing.Apsides`` (a-14),
against the golden files written by the UNMODIFIED reference and against the
CPU oracle on further seeded inputs.  Everything here is integer work: the
results must be bit-exact (``*_counts_final`` is float64 holding integers)."""
import numpy as np
import pytest

from fixture_io import expected_tree, list_fixtures, load_fixture
from parity import assert_same_array, assert_same_tree

pytestmark = pytest.mark.gpu


def _mods():
    from nbody_orbit_analysis_b200 import (h5shim, postprocessing, progenitors,
                                           storage)
    from oracle import orbit_oracle as oracle
    return h5shim, postprocessing, progenitors, storage, oracle


@pytest.mark.parametrize('name', list_fixtures('progen_'))
def test_progenitors_match_reference_fixture(name):
    _, _, progenitors, _, _ = _mods()
    fx = load_fixture(name)
    snap = {'ids': fx['in/ids'], 'coordinates': fx['in/coordinates'],
            'region_offsets': fx['in/region_offsets'],
            'box_size': float(fx['in/box_size'])}
    cids, coffs = progenitors.get_central_particle_ids(
        snap, fx['in/halo_positions'], n=int(fx['in/n']))
    assert_same_array('central_ids', cids, fx['out/central_ids'])
    assert_same_array('central_offsets', coffs, fx['out/central_offsets'])
    res = progenitors.find_main_progenitors(
        fx['in/halo_pids'], fx['in/halo_offsets'], cids, coffs)
    assert [int(x) for x in res] == fx['out/main_progenitors'].tolist()
    # known-answer test of SURVEY.md 8(a-13): tie -> smallest index, no hits
    # -> -1, repeated IDs only count at their first occurrence
    kat = progenitors.find_main_progenitors(
        np.array([10, 11, 12, 13, 20, 21, 22, 23]), np.array([0, 4]),
        np.array([10, 11, 20, 21, 99, 98, 97, 96, 22, 23, 20, 10]),
        np.array([0, 4, 8]))
    assert [int(x) for x in kat] == [0, -1, 1]


@pytest.mark.parametrize('dtype,cdtype,periodic', [
    (np.float32, np.float32, True), (np.float32, np.float64, True),
    (np.float64, np.float64, False)])
def test_central_ids_match_oracle(dtype, cdtype, periodic):
    _, _, progenitors, _, oracle = _mods()
    from nbody_orbit_analysis_b200.synth import SynthSim
    sim = SynthSim(60000, 40, 3, dtype=dtype, catalogue_dtype=cdtype,
                   periodic=periodic)
    t = 1
    pos, rad, _ = sim.regions(sim.snapshot_numbers[t], sim.main_branches[t])
    snap = sim.load_snapshot_data(sim.snapshot_numbers[t], pos, rad)
    for n in (1, 25, 100000):
        got = progenitors.get_central_particle_ids(snap, pos, n=n)
        exp = oracle.get_central_particle_ids(snap, pos, n=n)
        assert_same_array('central_ids', got[0], exp[0])
        assert_same_array('central_offsets', got[1], np.asarray(exp[1]))


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_find_main_progenitors_random(seed):
    _, _, progenitors, _, oracle = _mods()
    rng = np.random.default_rng(seed)
    n_halos, n_desc = 300, 260
    lens = rng.integers(0, 400, n_halos)
    lens[rng.integers(0, n_halos, 20)] = 0            # empty halos
    N = int(lens.sum())
    halo_pids = rng.permutation(10 * N)[:N].astype(np.int64) - 7   # unique
    halo_offsets = np.concatenate(([0], np.cumsum(lens)))[:-1]
    blocks = []
    for d in range(n_desc):
        k = int(rng.integers(0, 60))
        src = rng.integers(0, 3)
        if src == 0 or N == 0:                        # strangers
            blk = rng.integers(20 * N, 30 * N, k)
        else:                                         # mostly from 1-3 halos
            hs = rng.integers(0, n_halos, 3)
            pool = np.concatenate([halo_pids[halo_offsets[h]:halo_offsets[h] +
                                             lens[h]] for h in hs] +
                                  [rng.integers(20 * N, 30 * N, 5)])
            blk = rng.choice(pool, k) if len(pool) else np.zeros(0, int)
        blocks.append(np.asarray(blk, dtype=np.int64))
    t_lens = np.array([len(b) for b in blocks])
    tracked = np.concatenate(blocks)                  # with duplicates
    t_off = np.concatenate(([0], np.cumsum(t_lens)))[:-1]
    got = progenitors.find_main_progenitors(halo_pids, halo_offsets, tracked,
                                            t_off)
    exp = oracle.find_main_progenitors(halo_pids, halo_offsets, tracked, t_off)
    assert [int(x) for x in got] == [int(x) for x in exp]
    assert any(int(x) >= 0 for x in exp) and any(int(x) == -1 for x in exp)


def _write_src(h5shim, fx, path):
    with h5shim.File(path, 'w') as hf:
        for k, v in expected_tree(fx, 'in').items():
            if k.startswith('/__attr__/'):
                hf.attrs[k[len('/__attr__/'):]] = v if v.dtype.kind != 'U' \
                    else str(v)
            else:
                hf.create_dataset(k, data=v)


@pytest.mark.parametrize('name', list_fixtures('post_'))
def test_postprocessing_matches_reference_fixture(name, tmp_path):
    h5shim, postprocessing, _, storage, _ = _mods()
    fx = load_fixture(name)
    src = str(tmp_path / 'src.h5')
    _write_src(h5shim, fx, src)
    kw = dict(fx['meta']['kwargs'])
    if 'halo_ids' in kw:
        kw['halo_ids'] = np.array(kw['halo_ids'])
    collated = str(tmp_path / 'col.h5')
    postprocessing.Apsides(src).collate_apsides(savefile=collated,
                                                verbose=False, **kw)
    assert_same_tree(storage.tree(collated), expected_tree(fx))


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
def test_postprocessing_matches_oracle_end_to_end(mode, tmp_path):
    """track_orbits on the GPU -> Apsides on the GPU, against the oracle's
    Apsides on the same event file (counts, final counts, angle cut)."""
    h5shim, postprocessing, _, storage, oracle = _mods()
    from nbody_orbit_analysis_b200.synth import SynthSim
    from nbody_orbit_analysis_b200.track_orbits import track_orbits
    sim = SynthSim(40000, 9, 8, dtype=np.float32, catalogue_dtype=np.float32)
    events = str(tmp_path / 'events.h5')
    track_orbits(sim.snapshot_numbers, sim.main_branches, sim.regions,
                 sim.load_snapshot_data, events, mode=mode, verbose=False)
    for k, kw in enumerate((dict(save_final_counts=True),
                            dict(angle_cut=0.3, save_final_counts=True,
                                 halo_ids=np.array(
                                     sim.main_branches[-1][[6, 1, 4]])))):
        f_gpu = str(tmp_path / ('gpu%d.h5' % k))
        f_cpu = str(tmp_path / ('cpu%d.h5' % k))
        postprocessing.Apsides(events).collate_apsides(
            savefile=f_gpu, verbose=False, **kw)
        oracle.Apsides(events, storage=storage).collate_apsides(
            savefile=f_cpu, **kw)
        got, exp = storage.tree(f_gpu), storage.tree(f_cpu)
        assert any(k_.endswith('_counts_final') for k_ in exp)
        assert sum(len(v) for k_, v in exp.items()
                   if k_.endswith('particle_IDs')) > 0
        assert_same_tree(got, exp)


def test_postprocessing_rejects_untracked_halos(tmp_path):
    h5shim, postprocessing, _, _, _ = _mods()
    fx = load_fixture(list_fixtures('post_')[0])
    src = str(tmp_path / 'src.h5')
    _write_src(h5shim, fx, src)
    aps = postprocessing.Apsides(src)
    with pytest.raises(ValueError):
        aps.collate_apsides(halo_ids=np.array([-12345]),
                            savefile=str(tmp_path / 'x.h5'), verbose=False)
    assert list(aps.missing_halo_ids) == [-12345]
