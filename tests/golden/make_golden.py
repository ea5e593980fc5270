"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors of its own (SURVEY.md section
4), so these fixtures -- outputs of the real ``orbitanalysis`` package on
seeded synthetic inputs -- are what pins the oracle (``oracle/orbit_oracle.py``)
and, through it, the CUDA path.  h5py / pathos are replaced by the stand-ins in
``oracle/reference_harness.py``; nothing else is touched.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(HERE))

warnings.filterwarnings('ignore', category=DeprecationWarning)

from oracle.reference_harness import load_reference  # noqa: E402
from nbody_orbit_analysis_b200.synth import SynthSim  # noqa: E402
from nbody_orbit_analysis_b200 import storage  # noqa: E402
from fixture_io import Recorder, save_fixture  # noqa: E402

ref = load_reference()
TMP = tempfile.mkdtemp(prefix='oa_golden_')


def run_track(name, sim_kwargs, mode='pericentric', checkpoint=False,
              reverse=False, squeeze=False, resume_at=None, skip_rows=()):
    sim = SynthSim(**sim_kwargs)
    rec = Recorder(sim.regions, sim.load_snapshot_data)
    snaps = sim.snapshot_numbers.copy()
    mb = sim.main_branches.copy()
    for r in skip_rows:           # rows where no halo exists at all
        mb[r, :] = -1
    if squeeze:
        mb = mb[:, 0]
    if reverse:
        snaps, mb = snaps[::-1].copy(), mb[::-1].copy()
    savefile = os.path.join(TMP, name + '.h5')
    if 0 in skip_rows:
        # reference quirk (track_orbits.py:140): the savefile is only
        # initialised when the very first row is processed; with leading
        # all-empty rows the later 'r+' open fails.  Pre-initialise it the way
        # the reference itself would so that the istart logic can be pinned.
        ref.track_orbits.initialize_savefile(
            savefile, mode, sim.box if sim.periodic else None, False)
    if resume_at is None:
        ref.track_orbits.track_orbits(
            snaps, mb, rec.regions, rec.load_snapshot_data, savefile,
            mode=mode, checkpoint=checkpoint, npool=None, verbose=False)
    else:
        ref.track_orbits.track_orbits(
            snaps[:resume_at], mb[:resume_at], rec.regions,
            rec.load_snapshot_data, savefile, mode=mode, checkpoint=True,
            npool=None, verbose=False)
        ref.track_orbits.track_orbits(
            snaps, mb, rec.regions, rec.load_snapshot_data, savefile,
            mode=mode, checkpoint=True, resume=True, npool=None,
            verbose=False)
    out = storage.tree(savefile)
    if checkpoint or resume_at is not None:
        for k, v in storage.tree(savefile + '.checkpoint').items():
            out['/__checkpoint__' + k] = v
    inputs = dict(rec.data)
    inputs['in/snapshot_numbers'] = snaps
    inputs['in/main_branches'] = mb
    meta = dict(kind='track', mode=mode, checkpoint=bool(checkpoint),
                resume_at=resume_at, sim=_jsonable(sim_kwargs))
    path = save_fixture('track_' + name, meta, inputs, out)
    nev = sum(len(v) for k, v in out.items() if k.endswith('er_IDs'))
    print('%-28s %8.1f KB  events=%d' % (
        os.path.basename(path), os.path.getsize(path) / 1024, nev))
    return savefile, sim


def run_onthefly(name, sim_kwargs, t, mode='pericentric', drop_prev=(),
                 drop_now=()):
    sim = SynthSim(**sim_kwargs)
    rec = Recorder(sim.regions_onthefly, sim.load_snapshot_data)
    links = np.stack([sim.main_branches[t], sim.main_branches[t - 1]])
    for c in drop_prev:
        links[1, c] = -1
    for c in drop_now:
        links[0, c] = -1
    snap_no = int(sim.snapshot_numbers[t])
    savefile = os.path.join(TMP, name + '_{}.h5')
    ref.onthefly.track_orbits(
        snap_no, links, rec.regions, rec.load_snapshot_data, savefile,
        mode=mode, verbose=False)
    out = storage.tree(savefile.format('%0.3d' % snap_no))
    inputs = dict(rec.data)
    inputs['in/progenitor_links'] = links
    inputs['in/snapshot_number'] = np.asarray(snap_no)
    meta = dict(kind='onthefly', mode=mode, sim=_jsonable(sim_kwargs))
    path = save_fixture('onthefly_' + name, meta, inputs, out)
    print('%-28s %8.1f KB  keys=%s' % (
        os.path.basename(path), os.path.getsize(path) / 1024,
        sorted(k for k in out if 'IDs' in k)))


def run_postprocessing(name, savefile, **kw):
    aps = ref.postprocessing.Apsides(savefile)
    collated = os.path.join(TMP, name + '_collated.h5')
    aps.collate_apsides(savefile=collated, verbose=False, **kw)
    src = storage.tree(savefile)
    inputs = {'in' + k: v for k, v in src.items()}
    out = storage.tree(collated)
    meta = dict(kind='post', kwargs=_jsonable(kw))
    path = save_fixture('post_' + name, meta, inputs, out)
    print('%-28s %8.1f KB' % (
        os.path.basename(path), os.path.getsize(path) / 1024))


def run_progenitors(name, sim_kwargs, t, n):
    sim = SynthSim(**sim_kwargs)
    halo_ids = sim.main_branches[t]
    pos, rad, _ = sim.regions(sim.snapshot_numbers[t], halo_ids)
    snap = sim.load_snapshot_data(sim.snapshot_numbers[t], pos, rad)
    cids, coffs = ref.progenitors.get_central_particle_ids(snap, pos, n=n)
    # candidate progenitor halos one snapshot earlier, in a scrambled order
    pos0, rad0, _ = sim.regions(sim.snapshot_numbers[t - 1],
                                sim.main_branches[t - 1])
    snap0 = sim.load_snapshot_data(sim.snapshot_numbers[t - 1], pos0, rad0)
    res = ref.progenitors.find_main_progenitors(
        snap0['ids'], snap0['region_offsets'], cids, coffs)
    inputs = {
        'in/ids': snap['ids'], 'in/coordinates': snap['coordinates'],
        'in/region_offsets': snap['region_offsets'],
        'in/box_size': np.asarray(snap['box_size']),
        'in/halo_positions': pos, 'in/n': np.asarray(n),
        'in/halo_pids': snap0['ids'],
        'in/halo_offsets': snap0['region_offsets'],
    }
    out = {'/central_ids': cids, '/central_offsets': np.asarray(coffs),
           '/main_progenitors': np.asarray(res, dtype=np.int64)}
    # the survey's hand-checkable known-answer test
    kat = ref.progenitors.find_main_progenitors(
        np.array([10, 11, 12, 13, 20, 21, 22, 23]), np.array([0, 4]),
        np.array([10, 11, 20, 21, 99, 98, 97, 96, 22, 23, 20, 10]),
        np.array([0, 4, 8]))
    out['/kat'] = np.asarray(kat, dtype=np.int64)
    path = save_fixture('progen_' + name, dict(kind='progen'), inputs, out)
    print('%-28s %8.1f KB  kat=%s' % (
        os.path.basename(path), os.path.getsize(path) / 1024, kat))


def run_kernels(name, dtype, cdtype, seed):
    """Function-level vectors for the per-region kernels."""
    rng = np.random.default_rng(seed)
    n_prev, n_cur = 700, 720
    pool = rng.permutation(5000)[:900]
    ids_prev = pool[:n_prev].astype(np.int64)
    ids_cur = rng.permutation(pool[60:60 + n_cur]).astype(np.int64)
    L = 50.0
    centre = np.array([1.0, 49.5, 25.0], dtype=cdtype)
    out, inputs = {}, {}
    frames = []
    for tag, ids in (('prev', ids_prev), ('cur', ids_cur)):
        n = len(ids)
        x = (centre + rng.normal(0, 1.5, (n, 3))) % L
        v = rng.normal(0, 100, (n, 3))
        snap = {'coordinates': x.astype(dtype), 'velocities': v.astype(dtype),
                'masses': 1.0, 'box_size': L, 'redshift': 0.25}
        H = ref.utils.hubble_parameter(0.25, 0.07, 0.3, 0.7)
        bulk = np.array([3.0, -2.0, 1.0], dtype=cdtype)
        rh, vr, bv = ref.track_orbits.region_frame(
            snap, (0, n), centre, bulk, H)
        frames.append((rh, vr))
        inputs['in/%s/coordinates' % tag] = snap['coordinates']
        inputs['in/%s/velocities' % tag] = snap['velocities']
        inputs['in/%s/ids' % tag] = ids
        out['/%s/rhat' % tag] = rh
        out['/%s/vr' % tag] = vr
    inputs['in/centre'] = centre
    inputs['in/bulk'] = bulk
    inputs['in/H'] = np.asarray(H)
    inputs['in/box_size'] = np.asarray(L)
    inputs['in/redshift'] = np.asarray(0.25)
    for mode in ('pericentric', 'apocentric'):
        d = ref.track_orbits.compare_radial_velocities(
            ids_cur, ids_prev, frames[1][1], frames[0][1], frames[1][0],
            frames[0][0], mode)
        angles_prev = rng.uniform(0, 6, n_prev).astype(np.float16)
        ang, eang = ref.track_orbits.calc_angles(n_cur, angles_prev, d)
        inputs['in/angles_prev'] = angles_prev
        for k, v in d.items():
            out['/%s/%s' % (mode, k)] = v
        out['/%s/angles' % mode] = ang
        out['/%s/apsis_angles' % mode] = eang
        # every later mode draws a fresh angles_prev; keep them apart
        inputs['in/angles_prev_' + mode] = angles_prev
    out['/myin1d'] = ref.utils.myin1d(
        np.array([7, 5, 3, 1]), np.array([3, 7]))
    wrap_in = np.array([[50, -50, 49.9999], [50.0001, -50.0001, 151]])
    out['/recenter'] = ref.utils.recenter_coordinates(wrap_in.copy(), 100.0)
    inputs['in/recenter'] = wrap_in
    a, ea = ref.track_orbits.calc_angles(
        5, np.array([1, 9, 2, 3], dtype=np.float16),
        {'inds_departed': np.array([1]),
         'angle_changes': np.array([.3, np.nan, .2]),
         'apsis_inds': np.array([2]), 'inds_match': np.array([3, 0, 1])})
    out['/kat_angles'] = a
    out['/kat_apsis_angles'] = ea
    path = save_fixture('kernels_' + name, dict(kind='kernels'), inputs, out)
    print('%-28s %8.1f KB' % (
        os.path.basename(path), os.path.getsize(path) / 1024))


def run_regions(name, dtype, cdtype, periodic, seed):
    """The selection loop of the reference's example loader
    (``example_script.py:50-58``), executed with the REFERENCE's own
    ``utils.vector_norm`` / ``utils.recenter_coordinates``."""
    rng = np.random.default_rng(seed)
    L = 40.0
    n, n_h = 6000, 9
    centres = (rng.uniform(0, L, (n_h, 3))).astype(cdtype)
    centres[0] = [0.3, 39.8, 20.0]                 # on a box face
    centres[1] = centres[0] + np.array([1.0, -0.5, 0.7], dtype=cdtype)  # overlap
    radii = rng.uniform(1.0, 6.0, n_h).astype(cdtype)
    radii[2] = 0.01                                # an empty region
    host = rng.integers(0, n_h, n)
    x = (centres[host].astype(np.float64) +
         rng.normal(0, 2.5, (n, 3))) % L
    x[:50] = centres[3].astype(np.float64) + radii[3] * np.array([1.0, 0, 0]) \
        * (1 + rng.normal(0, 1e-7, (50, 1)))       # on the edge of a region
    coordinates = x.astype(dtype)
    box_size = L if periodic else None
    region_inds = []
    for position, radius in zip(np.atleast_2d(centres), np.atleast_1d(radii)):
        d = coordinates - position
        if periodic:
            d = ref.utils.recenter_coordinates(d, box_size)
        r = ref.utils.vector_norm(d)
        region_inds.append(np.argwhere(r < radius).flatten())
    region_lens = [len(inds) for inds in region_inds]
    out = {'/region_offsets': np.cumsum([0] + region_lens)[:-1],
           '/region_inds': np.hstack(region_inds).astype(int)}
    inputs = {'in/coordinates': coordinates, 'in/centres': centres,
              'in/radii': radii, 'in/box_size': np.asarray(L if periodic else -1.0)}
    path = save_fixture('regions_' + name, dict(kind='regions'), inputs, out)
    print('%-28s %8.1f KB' % (
        os.path.basename(path), os.path.getsize(path) / 1024))


def _jsonable(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, (np.dtype, type)):
            v = np.dtype(v).name
        elif isinstance(v, np.ndarray):
            v = v.tolist()
        out[k] = v
    return out


def main(only=None):
    f32, f64 = np.float32, np.float64
    if only == 'regions':
        run_regions('f32_c64', f32, f64, True, 11)
        run_regions('f32_c32', f32, f32, True, 12)
        run_regions('f64_c64_open', f64, f64, False, 13)
        return
    base = dict(n_particles=1600, n_halos=5, n_snap=7)
    sf, _ = run_track('peri_f32_cat64',
                      dict(base, late_halos=0.6, dtype=f32,
                           catalogue_dtype=f64),
                      checkpoint=True)
    run_postprocessing('peri_default', sf, save_final_counts=True)
    run_postprocessing('peri_cut_subset', sf, angle_cut=1.0,
                       halo_ids=np.array(storage.tree(sf)[
                           '/snapshot_016/halo_IDs'][[3, 0]]))
    run_track('apo_f32_cat32',
              dict(base, dtype=f32, catalogue_dtype=f32, box_vector=True),
              mode='apocentric')
    run_track('peri_f64_hubble',
              dict(base, dtype=f64, catalogue_dtype=f64, hubble=True))
    run_track('peri_f32_hubble',
              dict(base, dtype=f32, catalogue_dtype=f64, hubble=True))
    run_track('peri_f32_nobulk',
              dict(base, dtype=f32, catalogue_dtype=f64,
                   catalogue_bulk=False))
    run_track('apo_f64_massarr',
              dict(base, dtype=f64, catalogue_dtype=f64,
                   catalogue_bulk=False, mass_array=True), mode='apocentric')
    run_track('peri_f32_nonperiodic',
              dict(base, dtype=f32, catalogue_dtype=f64, periodic=False))
    run_track('peri_nfw_1halo_reversed',
              dict(n_particles=2500, n_halos=1, n_snap=8, nfw=True,
                   dtype=f32, catalogue_dtype=f64),
              reverse=True, squeeze=True)
    run_track('peri_leading_empty',
              dict(base, dtype=f32, catalogue_dtype=f64, late_halos=0.5,
                   seed=77), skip_rows=(0, 1, 4))
    run_track('peri_resume',
              dict(base, dtype=f32, catalogue_dtype=f64), resume_at=4)
    run_onthefly('peri_f32', dict(base, dtype=f32, catalogue_dtype=f64), t=3)
    run_onthefly('apo_f64_missing',
                 dict(base, dtype=f64, catalogue_dtype=f64), t=4,
                 mode='apocentric', drop_prev=(1,), drop_now=(3,))
    run_onthefly('peri_f32_cat32_massarr',
                 dict(base, dtype=f32, catalogue_dtype=f32, mass_array=True),
                 t=2)
    run_progenitors('n20', dict(n_particles=3000, n_halos=12, n_snap=4), t=2,
                    n=20)
    run_kernels('f32_c64', f32, f64, 1)
    run_kernels('f32_c32', f32, f32, 2)
    run_kernels('f64_c64', f64, f64, 3)
    run_regions('f32_c64', f32, f64, True, 11)
    run_regions('f32_c32', f32, f32, True, 12)
    run_regions('f64_c64_open', f64, f64, False, 13)


if __name__ == '__main__':
    # python tests/golden/make_golden.py [regions]: everything, or one family
    main(sys.argv[1] if len(sys.argv) > 1 else None)
