"""Pins the CPU oracle to the outputs of the UNMODIFIED reference.

The fixtures under tests/golden were written by tests/golden/make_golden.py,
which runs /root/reference/orbitanalysis itself.  Integers must match exactly;
floats are compared bit-for-bit as well (same numpy, same expressions), which
is what makes the oracle a trustworthy stand-in for the reference on the GPU
box where /root/reference does not exist.
"""
import numpy as np
import pytest

from fixture_io import Replay, expected_tree, list_fixtures, load_fixture
from parity import assert_same_array, assert_same_tree

from nbody_orbit_analysis_b200 import h5shim, storage
from oracle import orbit_oracle as oracle


@pytest.mark.parametrize('name', list_fixtures('track_'))
def test_track_orbits_matches_reference(name, tmp_path):
    fx = load_fixture(name)
    meta = fx['meta']
    rp = Replay(fx)
    savefile = str(tmp_path / 'o.h5')
    snaps, mb = fx['in/snapshot_numbers'], fx['in/main_branches']
    if 0 in np.flatnonzero((np.atleast_2d(mb.T).T == -1).all(axis=1)) or \
            name == 'track_peri_leading_empty':
        with h5shim.File(savefile, 'w') as hf:   # see make_golden.run_track
            hf.attrs['mode'] = meta['mode']
            hf.attrs['box_size'] = fx['out/__attr__/box_size']
    if meta['resume_at'] is None:
        oracle.track_orbits(snaps, mb, rp.regions, rp.load_snapshot_data,
                            savefile, mode=meta['mode'],
                            checkpoint=meta['checkpoint'], storage=h5shim)
    else:
        k = meta['resume_at']
        oracle.track_orbits(snaps[:k], mb[:k], rp.regions,
                            rp.load_snapshot_data, savefile, mode=meta['mode'],
                            checkpoint=True, storage=h5shim)
        oracle.track_orbits(snaps, mb, rp.regions, rp.load_snapshot_data,
                            savefile, mode=meta['mode'], checkpoint=True,
                            resume=True, storage=h5shim)
    got = storage.tree(savefile)
    if meta['checkpoint'] or meta['resume_at'] is not None:
        for k, v in storage.tree(savefile + '.checkpoint').items():
            got['/__checkpoint__' + k] = v
    assert_same_tree(got, expected_tree(fx))


@pytest.mark.parametrize('name', list_fixtures('onthefly_'))
def test_onthefly_matches_reference(name, tmp_path):
    fx = load_fixture(name)
    rp = Replay(fx)
    snap_no = int(fx['in/snapshot_number'])
    savefile = str(tmp_path / 'otf_{}.h5')
    with np.errstate(all='ignore'):
        oracle.track_orbits_onthefly(
            snap_no, fx['in/progenitor_links'], rp.regions,
            rp.load_snapshot_data, savefile, mode=fx['meta']['mode'],
            storage=h5shim)
    got = storage.tree(savefile.format('%0.3d' % snap_no))
    assert_same_tree(got, expected_tree(fx))


@pytest.mark.parametrize('name', list_fixtures('post_'))
def test_postprocessing_matches_reference(name, tmp_path):
    fx = load_fixture(name)
    src = str(tmp_path / 'src.h5')
    with h5shim.File(src, 'w') as hf:
        for k, v in expected_tree(fx, 'in').items():
            if k.startswith('/__attr__/'):
                hf.attrs[k[len('/__attr__/'):]] = v if v.dtype.kind != 'U' \
                    else str(v)
            else:
                hf.create_dataset(k, data=v)
    kw = dict(fx['meta']['kwargs'])
    if 'halo_ids' in kw:
        kw['halo_ids'] = np.array(kw['halo_ids'])
    collated = str(tmp_path / 'col.h5')
    oracle.Apsides(src, storage=h5shim).collate_apsides(
        savefile=collated, **kw)
    assert_same_tree(storage.tree(collated), expected_tree(fx))


@pytest.mark.parametrize('name', list_fixtures('progen_'))
def test_progenitors_match_reference(name):
    fx = load_fixture(name)
    snap = {'ids': fx['in/ids'], 'coordinates': fx['in/coordinates'],
            'region_offsets': fx['in/region_offsets'],
            'box_size': float(fx['in/box_size'])}
    cids, coffs = oracle.get_central_particle_ids(
        snap, fx['in/halo_positions'], n=int(fx['in/n']))
    assert_same_array('central_ids', cids, fx['out/central_ids'])
    assert_same_array('central_offsets', coffs, fx['out/central_offsets'])
    res = oracle.find_main_progenitors(
        fx['in/halo_pids'], fx['in/halo_offsets'], cids, coffs)
    assert [int(x) for x in res] == fx['out/main_progenitors'].tolist()
    kat = oracle.find_main_progenitors(
        np.array([10, 11, 12, 13, 20, 21, 22, 23]), np.array([0, 4]),
        np.array([10, 11, 20, 21, 99, 98, 97, 96, 22, 23, 20, 10]),
        np.array([0, 4, 8]))
    assert [int(x) for x in kat] == fx['out/kat'].tolist() == [0, -1, 1]


@pytest.mark.parametrize('name', list_fixtures('kernels_'))
def test_region_kernels_match_reference(name):
    fx = load_fixture(name)
    H = fx['in/H'][()]
    frames = {}
    for tag in ('prev', 'cur'):
        x = fx['in/%s/coordinates' % tag]
        snap = {'coordinates': x, 'velocities': fx['in/%s/velocities' % tag],
                'masses': 1.0, 'box_size': float(fx['in/box_size']),
                'redshift': float(fx['in/redshift'])}
        rh, vr, _ = oracle.region_frame(
            snap, (0, len(x)), fx['in/centre'], fx['in/bulk'], H)
        assert_same_array(tag + '/rhat', rh, fx['out/%s/rhat' % tag])
        assert_same_array(tag + '/vr', vr, fx['out/%s/vr' % tag])
        frames[tag] = (rh, vr)
    for mode in oracle.MODES:
        d = oracle.compare_radial_velocities(
            fx['in/cur/ids'], fx['in/prev/ids'], frames['cur'][1],
            frames['prev'][1], frames['cur'][0], frames['prev'][0], mode)
        for k, v in d.items():
            assert_same_array(mode + '/' + k, v, fx['out/%s/%s' % (mode, k)])
        ang, eang = oracle.calc_angles(
            len(fx['in/cur/ids']), fx['in/angles_prev_' + mode], d)
        assert_same_array('angles', ang, fx['out/%s/angles' % mode])
        assert_same_array('apsis_angles', eang,
                          fx['out/%s/apsis_angles' % mode])
    assert oracle.ordered_match(
        np.array([7, 5, 3, 1]), np.array([3, 7])).tolist() == \
        fx['out/myin1d'].tolist() == [2, 0]
    assert_same_array('recenter', oracle.minimum_image(
        fx['in/recenter'].copy(), 100.0), fx['out/recenter'])
    a, ea = oracle.calc_angles(
        5, np.array([1, 9, 2, 3], dtype=np.float16),
        {'inds_departed': np.array([1]),
         'angle_changes': np.array([.3, np.nan, .2]),
         'apsis_inds': np.array([2]), 'inds_match': np.array([3, 0, 1])})
    assert_same_array('kat_angles', a, fx['out/kat_angles'])
    assert_same_array('kat_apsis', ea, fx['out/kat_apsis_angles'])


def test_hubble_parameter():
    assert oracle.hubble_parameter(0.0, 70.0, 0.3, 0.7) == 70.0
    h = oracle.hubble_parameter(1.0, 70.0, 0.3, 0.7, 0.0)
    assert h == 70.0 * np.sqrt(0.3 * 8 + 0.7)


@pytest.mark.parametrize('name', list_fixtures('regions_'))
def test_region_extraction_matches_reference(name):
    """oracle.extract_regions against the selection loop of the reference's
    example loader run with the reference's own utils (make_golden.py
    run_regions; example_script.py:50-58)."""
    fx = load_fixture(name)
    box = float(fx['in/box_size'])
    inds, offs = oracle.extract_regions(
        fx['in/coordinates'], fx['in/centres'], fx['in/radii'],
        box if box > 0 else None)
    assert len(fx['out/region_inds']) > 100
    assert_same_array('region_inds', inds, fx['out/region_inds'])
    assert_same_array('region_offsets', offs, fx['out/region_offsets'])


@pytest.mark.parametrize('name', list_fixtures('regions_'))
def test_region_cell_lists_cover_every_member(name):
    """Host side of the GPU region extraction: every (particle, region) pair of
    the reference selection has the region in the list of the particle's cell
    (several grid resolutions, periodic wrap included)."""
    from nbody_orbit_analysis_b200 import regions
    fx = load_fixture(name)
    box = float(fx['in/box_size'])
    periodic = box > 0
    x = fx['in/coordinates'].astype(np.float64)
    c64 = fx['in/centres'].astype(np.float64)
    r64 = fx['in/radii'].astype(np.float64)
    inds = fx['out/region_inds']
    offs = np.append(fx['out/region_offsets'], len(inds))
    reg = np.repeat(np.arange(len(r64)), np.diff(offs))
    for cells in (None, 1, 7, 64):
        lo, inv, dim = regions._grid(c64, r64, np.full(3, box) if periodic else None,
                                     cells)
        start, lists = regions._cell_lists(c64, r64, lo, inv, dim, periodic)
        k = np.floor((x[inds] - lo) * inv).astype(np.int64)
        if periodic:
            k %= dim
        else:
            assert (k >= 0).all() and (k < dim).all()
        cell = (k[:, 2] * dim[1] + k[:, 1]) * dim[0] + k[:, 0]
        for i in range(0, len(inds), 7):
            assert reg[i] in lists[start[cell[i]]:start[cell[i] + 1]]
