"""GPU parity tests of the on-the-fly entry point (``-m gpu``).

``nbody_orbit_analysis_b200.track_orbits_onthefly.track_orbits`` against

* the golden files written by the UNMODIFIED reference
  (``tests/golden/onthefly_*.npz``; reference ``track_orbits_onthefly.py``), and
* the CPU oracle on seeded synthetic inputs, both modes, with halos that lack a
  progenitor / a descendant.

Integer datasets (IDs, offsets, links) are bit-exact, including the sorted
entered / departed lists and the ``apocentrer_*`` dataset names; ``angles``
(raw arccos of every matched particle) within rel 1e-5 (fp32) / 1e-6 (fp64)
plus an absolute term for arccos near 1; derived bulk velocities to 3e-5 of
their scale (the reference accumulates sequentially in the input dtype,
SURVEY.md 7.5).
"""
import numpy as np
import pytest

from fixture_io import Replay, expected_tree, list_fixtures, load_fixture
from parity import RTOL, assert_same_array

pytestmark = pytest.mark.gpu


def compare_onthefly_trees(got, exp, data_f64):
    assert sorted(got) == sorted(exp), (
        sorted(set(got) - set(exp)), sorted(set(exp) - set(got)))
    rtol = RTOL['float64' if data_f64 else 'float32']
    for k in sorted(exp):
        g, e = np.asarray(got[k]), np.asarray(exp[k])
        if k.endswith('/angles'):
            assert g.dtype == e.dtype, (k, g.dtype, e.dtype)
            assert g.shape == e.shape, (k, g.shape, e.shape)
            # arccos(x) near x = 1 amplifies the last-ulp difference of the
            # dot product: |d acos| ~ ulp / sqrt(2 ulp)
            atol = 2e-3 if not data_f64 else 1e-7
            both_nan = np.isnan(g) & np.isnan(e)
            ok = both_nan | (np.abs(g - e) <= atol + rtol * np.abs(e))
            assert ok.all(), '%s: %d of %d differ' % (k, (~ok).sum(), ok.size)
        elif k.endswith('/bulk_velocities'):
            # derived on the device in numpy's summation order: bit-identical
            assert g.dtype == e.dtype and g.shape == e.shape, k
            assert np.array_equal(g, e, equal_nan=True), k
        elif e.dtype.kind == 'f':
            assert_same_array(k, g, e, exact_float=False, rtol=rtol)
        else:
            assert_same_array(k, g, e)


@pytest.mark.parametrize('name', list_fixtures('onthefly_'))
def test_onthefly_matches_reference_fixture(name, tmp_path):
    from nbody_orbit_analysis_b200 import storage, track_orbits_onthefly
    fx = load_fixture(name)
    rp = Replay(fx)
    snap_no = int(fx['in/snapshot_number'])
    savefile = str(tmp_path / 'otf_{}.h5')
    track_orbits_onthefly.track_orbits(
        snap_no, fx['in/progenitor_links'], rp.regions, rp.load_snapshot_data,
        savefile, mode=fx['meta']['mode'], verbose=False)
    got = storage.tree(savefile.format('%0.3d' % snap_no))
    compare_onthefly_trees(got, expected_tree(fx),
                           fx['meta']['sim']['dtype'] == 'float64')


CASES = [
    (50000, 23, np.float32, np.float32, {}, (), ()),
    (50000, 23, np.float32, np.float64, {}, (2, 7), (5,)),
    (30000, 9, np.float64, np.float64, {'mass_array': True}, (0,), (8,)),
    (6000, 1500, np.float32, np.float32, {}, (3, 4, 5), (10,)),
]


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
@pytest.mark.parametrize('case', CASES, ids=[
    'f32', 'f32c64_missing', 'f64_massarr_missing', 'tiny_blocks'])
def test_onthefly_matches_oracle(case, mode, tmp_path):
    from nbody_orbit_analysis_b200 import storage, track_orbits_onthefly
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    n, nh, dt, cdt, kw, drop_prev, drop_now = case
    sim = SynthSim(n, nh, 5, dtype=dt, catalogue_dtype=cdt, **kw)
    t = 3
    links = np.stack([sim.main_branches[t], sim.main_branches[t - 1]])
    for c in drop_prev:
        links[1, c] = -1
    for c in drop_now:
        links[0, c] = -1
    snap_no = int(sim.snapshot_numbers[t])
    f_gpu, f_cpu = str(tmp_path / 'g_{}.h5'), str(tmp_path / 'c_{}.h5')
    track_orbits_onthefly.track_orbits(
        snap_no, links, sim.regions_onthefly, sim.load_snapshot_data, f_gpu,
        mode=mode, verbose=False)
    with np.errstate(all='ignore'):
        oracle.track_orbits_onthefly(
            snap_no, links, sim.regions_onthefly, sim.load_snapshot_data,
            f_cpu, mode=mode, storage=storage)
    got = storage.tree(f_gpu.format('%0.3d' % snap_no))
    exp = storage.tree(f_cpu.format('%0.3d' % snap_no))
    tag = 'pericenter' if mode == 'pericentric' else 'apocentrer'
    assert '/%s_IDs' % tag in got
    assert len(exp['/%s_IDs' % tag]) > 0
    compare_onthefly_trees(got, exp, dt == np.float64)


def test_onthefly_bad_mode():
    from nbody_orbit_analysis_b200 import track_orbits_onthefly
    with pytest.raises(ValueError):
        track_orbits_onthefly.track_orbits(
            3, np.zeros((2, 1), dtype=int), None, None, 'x_{}.h5', mode='both')
