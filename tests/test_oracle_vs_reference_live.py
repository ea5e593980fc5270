"""Property tests: the CPU oracle against the UNMODIFIED reference, run live.

The committed fixtures (tests/golden, test_oracle_golden.py) pin the oracle on
22 fixed cases.  Here hypothesis draws the configuration -- particle and halo
counts, number of snapshots, dtypes, periodic / open box, scalar or (3,) box,
Hubble flow, catalogue or derived bulk velocity, mass arrays, halos appearing
mid-run, both modes -- and the oracle must write, bit for bit, the file the
reference package writes on the same callbacks (SURVEY.md sections 4 and 8(c):
"parity is pinned only by executing the reference itself on shared synthetic
inputs").  Covers ``track_orbits`` (reference ``track_orbits.py:9-244``), the
on-the-fly entry point (``track_orbits_onthefly.py:8-252``), ``Apsides``
(``postprocessing.py:30-240``) and ``progenitors.py:5-117``.

The reference is imported from the mounted tree in the build container or from
the copy ``make -C oracle ref`` vendors into ``oracle/_ref`` (which also travels
to the GPU box); without either the module is skipped.  CPU only.
"""
import warnings

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from parity import assert_same_tree

from nbody_orbit_analysis_b200 import storage
from nbody_orbit_analysis_b200.synth import SynthSim
from oracle import orbit_oracle as oracle
from oracle import reference_harness

pytestmark = pytest.mark.skipif(
    not reference_harness.available(),
    reason='the unmodified reference is neither mounted nor vendored '
           '(make -C oracle ref)')

SEEN = {'track_files': 0, 'track_events': 0, 'track_errors': 0, 'otf_files': 0,
        'otf_events': 0, 'collated': 0, 'progenitor_lists': 0, 'progenitors_found': 0}
SETTINGS = dict(deadline=None, derandomize=True, database=None,
                suppress_health_check=list(HealthCheck))


@pytest.fixture(scope='module')
def ref():
    warnings.filterwarnings('ignore', category=DeprecationWarning)
    return reference_harness.load_reference()


sims = st.fixed_dictionaries(dict(
    n_particles=st.integers(300, 4000),
    n_halos=st.integers(1, 9),
    n_snap=st.integers(2, 5),
    seed=st.integers(1, 2 ** 31),
    dtype=st.sampled_from([np.float32, np.float64]),
    catalogue_dtype=st.sampled_from([np.float32, np.float64]),
    nfw=st.booleans(), hubble=st.booleans(), catalogue_bulk=st.booleans(),
    mass_array=st.booleans(), periodic=st.booleans(),
    box_vector=st.booleans(),
    late_halos=st.sampled_from([0.0, 0.0, 0.4]),
))


def _storage_of_reference(ref):
    """The module the reference writes through (real h5py if this box has it,
    else the stand-in reference_harness registered)."""
    import sys
    return sys.modules['h5py']


@settings(max_examples=60, **SETTINGS)
@given(kw=sims, mode=st.sampled_from(['pericentric', 'apocentric']),
       checkpoint=st.booleans(), reverse=st.booleans())
def test_track_orbits_equals_reference(ref, tmp_path_factory, kw, mode,
                                       checkpoint, reverse):
    tmp = tmp_path_factory.mktemp('live')
    if kw['box_vector'] and not kw['periodic']:
        kw = dict(kw, box_vector=False)
    sim = SynthSim(**kw)
    snaps, mb = sim.snapshot_numbers.copy(), sim.main_branches.copy()
    if reverse:                      # the driver sorts by snapshot number
        snaps, mb = snaps[::-1].copy(), mb[::-1].copy()
    if not (mb[np.argsort(snaps)][0] != -1).any():
        return                       # reference quirk (:140): savefile never made
    f_ref, f_ora = str(tmp / 'ref.h5'), str(tmp / 'ora.h5')
    backend = _storage_of_reference(ref)
    with np.errstate(all='ignore'):
        try:
            ref.track_orbits.track_orbits(
                snaps, mb, sim.regions, sim.load_snapshot_data, f_ref,
                mode=mode, checkpoint=checkpoint, npool=None, verbose=False)
            failed = None
        except ValueError as exc:    # e.g. a snapshot without matched halos (:216)
            failed = exc
        if failed is not None:
            with pytest.raises(ValueError):
                oracle.track_orbits(snaps, mb, sim.regions,
                                    sim.load_snapshot_data, f_ora, mode=mode,
                                    checkpoint=checkpoint, storage=backend)
            SEEN['track_errors'] += 1
            return
        oracle.track_orbits(snaps, mb, sim.regions, sim.load_snapshot_data,
                            f_ora, mode=mode, checkpoint=checkpoint,
                            storage=backend)
    got, exp = storage.tree(f_ora), storage.tree(f_ref)
    if checkpoint:
        for k, v in storage.tree(f_ref + '.checkpoint').items():
            exp['/__checkpoint__' + k] = v
        for k, v in storage.tree(f_ora + '.checkpoint').items():
            got['/__checkpoint__' + k] = v
    assert_same_tree(got, exp)
    SEEN['track_files'] += 1
    SEEN['track_events'] += sum(len(v) for k, v in exp.items()
                                if k.endswith('er_IDs'))


@settings(max_examples=20, **SETTINGS)
@given(kw=sims, mode=st.sampled_from(['pericentric', 'apocentric']),
       cut=st.integers(2, 4))
def test_resume_equals_reference(ref, tmp_path_factory, kw, mode, cut):
    """checkpoint=True up to snapshot `cut`, then resume=True over the whole
    list (reference ``track_orbits.py:93-101, 229-232, 390-394``)."""
    tmp = tmp_path_factory.mktemp('live_resume')
    kw = dict(kw, n_snap=5, late_halos=0.0,
              box_vector=kw['box_vector'] and kw['periodic'])
    sim = SynthSim(**kw)
    snaps, mb = sim.snapshot_numbers, sim.main_branches
    backend = _storage_of_reference(ref)
    files = []
    with np.errstate(all='ignore'):
        for run, extra in ((ref.track_orbits.track_orbits,
                            dict(npool=None, verbose=False)),
                           (oracle.track_orbits, dict(storage=backend))):
            f = str(tmp / ('%d.h5' % len(files)))
            run(snaps[:cut], mb[:cut], sim.regions, sim.load_snapshot_data, f,
                mode=mode, checkpoint=True, **extra)
            run(snaps, mb, sim.regions, sim.load_snapshot_data, f, mode=mode,
                checkpoint=True, resume=True, **extra)
            tree = storage.tree(f)
            for k, v in storage.tree(f + '.checkpoint').items():
                tree['/__checkpoint__' + k] = v
            files.append(tree)
    assert_same_tree(files[1], files[0])
    SEEN['resumed'] = SEEN.get('resumed', 0) + 1


@settings(max_examples=40, **SETTINGS)
@given(kw=sims, mode=st.sampled_from(['pericentric', 'apocentric']),
       drop=st.lists(st.tuples(st.integers(0, 1), st.integers(0, 8)),
                     max_size=3))
def test_onthefly_equals_reference(ref, tmp_path_factory, kw, mode, drop):
    tmp = tmp_path_factory.mktemp('live_otf')
    kw = dict(kw, late_halos=0.0, hubble=False)
    if kw['box_vector'] and not kw['periodic']:
        kw['box_vector'] = False
    sim = SynthSim(**kw)
    t = sim.n_snap - 1
    links = np.stack([sim.main_branches[t], sim.main_branches[t - 1]])
    for row, col in drop:            # halos missing at s or at s-1
        links[row, col % sim.n_halos] = -1
    if not (links[0] != -1).any() or not (links[1] != -1).any():
        return
    snap_no = int(sim.snapshot_numbers[t])
    f_ref, f_ora = str(tmp / 'ref_{}.h5'), str(tmp / 'ora_{}.h5')
    backend = _storage_of_reference(ref)
    with np.errstate(all='ignore'):
        ref.onthefly.track_orbits(snap_no, links, sim.regions_onthefly,
                                  sim.load_snapshot_data, f_ref, mode=mode,
                                  verbose=False)
        oracle.track_orbits_onthefly(snap_no, links, sim.regions_onthefly,
                                     sim.load_snapshot_data, f_ora, mode=mode,
                                     storage=backend)
    name = '%0.3d' % snap_no
    exp = storage.tree(f_ref.format(name))
    assert_same_tree(storage.tree(f_ora.format(name)), exp)
    SEEN['otf_files'] += 1
    SEEN['otf_events'] += sum(len(v) for k, v in exp.items()
                              if k.endswith('er_IDs'))


@settings(max_examples=20, **SETTINGS)
@given(kw=sims, mode=st.sampled_from(['pericentric', 'apocentric']),
       angle_cut=st.sampled_from([0.0, np.pi / 4, 1.5]),
       final=st.booleans())
def test_collation_equals_reference(ref, tmp_path_factory, kw, mode, angle_cut,
                                    final):
    tmp = tmp_path_factory.mktemp('live_post')
    kw = dict(kw, n_snap=max(kw['n_snap'], 3), late_halos=0.0,
              box_vector=kw['box_vector'] and kw['periodic'])
    sim = SynthSim(**kw)
    src = str(tmp / 'events.h5')
    backend = _storage_of_reference(ref)
    with np.errstate(all='ignore'):
        ref.track_orbits.track_orbits(
            sim.snapshot_numbers, sim.main_branches, sim.regions,
            sim.load_snapshot_data, src, mode=mode, npool=None, verbose=False)
        f_ref, f_ora = str(tmp / 'c_ref.h5'), str(tmp / 'c_ora.h5')
        ref.postprocessing.Apsides(src).collate_apsides(
            angle_cut=angle_cut, save_final_counts=final, savefile=f_ref,
            verbose=False)
        oracle.Apsides(src, storage=backend).collate_apsides(
            angle_cut=angle_cut, save_final_counts=final, savefile=f_ora,
            verbose=False)
    assert_same_tree(storage.tree(f_ora), storage.tree(f_ref))
    SEEN['collated'] += 1


@settings(max_examples=30, **SETTINGS)
@given(kw=sims, n=st.integers(1, 40), t=st.integers(1, 4))
def test_progenitors_equal_reference(ref, kw, n, t):
    kw = dict(kw, late_halos=0.0,
              box_vector=kw['box_vector'] and kw['periodic'])
    sim = SynthSim(**kw)
    t = min(t, sim.n_snap - 1)
    pos, rad, _ = sim.regions(sim.snapshot_numbers[t], sim.main_branches[t])
    snap = sim.load_snapshot_data(sim.snapshot_numbers[t], pos, rad)
    pos0, rad0, _ = sim.regions(sim.snapshot_numbers[t - 1],
                                sim.main_branches[t - 1])
    snap0 = sim.load_snapshot_data(sim.snapshot_numbers[t - 1], pos0, rad0)
    with np.errstate(all='ignore'):
        cids_r, coffs_r = ref.progenitors.get_central_particle_ids(
            snap, pos, n=n)
        cids_o, coffs_o = oracle.get_central_particle_ids(snap, pos, n=n)
        assert np.array_equal(cids_r, cids_o) and cids_r.dtype == cids_o.dtype
        assert np.array_equal(np.asarray(coffs_r), np.asarray(coffs_o))
        res_r = ref.progenitors.find_main_progenitors(
            snap0['ids'], snap0['region_offsets'], cids_r, coffs_r)
        res_o = oracle.find_main_progenitors(
            snap0['ids'], snap0['region_offsets'], cids_o, coffs_o)
    assert [int(x) for x in res_r] == [int(x) for x in res_o]
    SEEN['progenitor_lists'] += 1
    SEEN['progenitors_found'] += sum(int(x) >= 0 for x in res_r)


def test_zz_the_drawn_cases_were_not_trivial():
    """(runs last in this module) the properties above compared real work."""
    assert SEEN['track_files'] >= 30 and SEEN['track_events'] > 2000, SEEN
    assert SEEN['otf_files'] >= 20 and SEEN['otf_events'] > 500, SEEN
    assert SEEN['collated'] >= 10 and SEEN['progenitor_lists'] >= 15, SEEN
    assert SEEN.get('resumed', 0) >= 10, SEEN
    assert SEEN['progenitors_found'] > 10, SEEN
    print('live oracle-vs-reference cases:', SEEN)


def test_vendored_reference_is_unmodified():
    """``oracle/_ref`` (what travels to the GPU box for the CPU arm) holds the
    reference's Python files byte for byte -- checked wherever the mounted tree
    and the vendored copy both exist (the build container)."""
    import filecmp
    import os
    here = os.path.dirname(os.path.abspath(reference_harness.__file__))
    vendored = os.path.join(here, '_ref', 'orbitanalysis')
    mounted = os.path.join(os.environ.get('OA_REFERENCE_ROOT', '/root/reference'),
                           'orbitanalysis')
    if not (os.path.isdir(vendored) and os.path.isdir(mounted)):
        pytest.skip('needs both the mounted reference and oracle/_ref')
    names = sorted(f for f in os.listdir(mounted) if f.endswith('.py'))
    assert names and names == sorted(
        f for f in os.listdir(vendored) if f.endswith('.py'))
    for name in names:
        assert filecmp.cmp(os.path.join(mounted, name),
                           os.path.join(vendored, name), shallow=False), name
