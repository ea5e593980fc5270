"""GPU region extraction (``regions.extract_regions``, csrc/oa_regions.cu)
against the reference selection (golden fixtures made with the reference's own
utils, ``example_script.py:50-58``) and against the oracle on larger seeded
inputs; then the extracted snapshot through the tracking path."""
import numpy as np
import pytest

from fixture_io import list_fixtures, load_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('cells', [None, 1, 5, 64])
@pytest.mark.parametrize('name', list_fixtures('regions_'))
def test_extract_regions_matches_reference(name, cells):
    from nbody_orbit_analysis_b200.regions import extract_regions
    fx = load_fixture(name)
    box = float(fx['in/box_size'])
    x = fx['in/coordinates']
    n = len(x)
    ids = np.arange(n, dtype=np.int64) * 7 + 3
    vel = (x * 2 + 1).astype(x.dtype)
    snap = extract_regions(x, fx['in/centres'], fx['in/radii'],
                           box_size=box if box > 0 else None, ids=ids,
                           velocities=vel, masses=1.0, cells_per_axis=cells)
    inds = fx['out/region_inds']
    assert np.array_equal(snap['region_inds'], inds)
    assert np.array_equal(snap['region_offsets'], fx['out/region_offsets'])
    assert np.array_equal(snap['ids'], ids[inds])
    assert np.array_equal(snap['coordinates'], x[inds])
    assert np.array_equal(snap['velocities'], vel[inds])
    assert snap['masses'] == 1.0


@pytest.mark.parametrize('dtype,cdtype', [(np.float32, np.float32),
                                          (np.float32, np.float64),
                                          (np.float64, np.float64)])
def test_extract_regions_matches_oracle_at_scale(dtype, cdtype):
    """200 k particles, 300 overlapping regions in a periodic box: identical
    indices and offsets (bit-exact selection), mass array gathered."""
    from nbody_orbit_analysis_b200.regions import extract_regions
    from oracle import orbit_oracle as oracle
    rng = np.random.default_rng(5)
    L, n, n_h = 100.0, 200000, 300
    centres = rng.uniform(0, L, (n_h, 3)).astype(cdtype)
    radii = (rng.uniform(0.5, 4.0, n_h) ** 1.5).astype(cdtype)
    host = rng.integers(0, n_h, n)
    x = ((centres[host].astype(np.float64) + rng.normal(0, 2.0, (n, 3))) % L
         ).astype(dtype)
    m = rng.uniform(1, 2, n).astype(dtype)
    snap = extract_regions(x, centres, radii, box_size=L, masses=m,
                           ids=np.arange(n, dtype=np.int64))
    inds, offs = oracle.extract_regions(x, centres, radii, L)
    assert len(inds) > n // 4
    assert np.array_equal(snap['region_inds'], inds)
    assert np.array_equal(snap['region_offsets'], offs)
    assert np.array_equal(snap['masses'], m[inds])


def test_extracted_snapshot_tracks_like_the_loader(tmp_path):
    """A loader built on ``extract_regions`` (the whole snapshot -> regions on
    the GPU) gives the same ``track_orbits`` file as SynthSim's own loader."""
    from nbody_orbit_analysis_b200 import storage, track_orbits
    from nbody_orbit_analysis_b200.regions import extract_regions
    from nbody_orbit_analysis_b200.synth import SynthSim
    from test_gpu_track import compare_track_trees
    sim = SynthSim(40000, 6, 5, dtype=np.float32, catalogue_dtype=np.float64)
    whole = {}

    def loader(sn, pos, rad):
        # every particle of the universe once (blocks of SynthSim overlap only
        # through the selection radius)
        if sn not in whole:
            full = sim.load_snapshot_data(sn, pos, np.full(len(rad), 1e9))
            _, first = np.unique(full['ids'], return_index=True)
            whole[sn] = {k: full[k][first] for k in
                         ('ids', 'coordinates', 'velocities')}
            whole[sn]['meta'] = {k: full[k] for k in full if k not in
                                 ('ids', 'coordinates', 'velocities',
                                  'region_offsets')}
        w = whole[sn]
        snap = extract_regions(w['coordinates'], pos, rad,
                               box_size=w['meta'].get('box_size'), ids=w['ids'],
                               velocities=w['velocities'], masses=1.0)
        snap.update({k: v for k, v in w['meta'].items() if k != 'masses'})
        return snap

    def loader_ref(sn, pos, rad):
        from oracle import orbit_oracle as oracle
        loader(sn, pos, rad)
        w = whole[sn]
        inds, offs = oracle.extract_regions(w['coordinates'], pos, rad,
                                            w['meta'].get('box_size'))
        snap = {'ids': w['ids'][inds], 'coordinates': w['coordinates'][inds],
                'velocities': w['velocities'][inds], 'masses': 1.0,
                'region_offsets': offs}
        snap.update({k: v for k, v in w['meta'].items() if k != 'masses'})
        return snap
    f_a, f_b = str(tmp_path / 'a.h5'), str(tmp_path / 'b.h5')
    args = (sim.snapshot_numbers, sim.main_branches, sim.regions)
    track_orbits.track_orbits(*args, loader, f_a, verbose=False)
    track_orbits.track_orbits(*args, loader_ref, f_b, verbose=False)
    got, exp = storage.tree(f_a), storage.tree(f_b)
    assert sum(len(v) for k, v in exp.items() if k.endswith('er_IDs')) > 0
    compare_track_trees(got, exp, data_f64=False, derived_bulk=False)
