"""GPU parity tests of the tracking path (run on the B200 box, ``-m gpu``).

Every test drives the CUDA path through the C ABI (``liborbit_b200.so`` via
``nbody_orbit_analysis_b200._lib``) and compares with

* the golden fixtures written by the UNMODIFIED reference
  (``tests/golden/*.npz``), and
* the CPU oracle (``oracle/orbit_oracle.py``) on seeded synthetic inputs.

Acceptance (BASELINE.json north_star / SURVEY.md 8(c)): bit-exact match
indices, event IDs and order, offsets; floats within rel 1e-6 (fp64) / 1e-5
(fp32); float16 angles equal or within 1 f16 ulp.
"""
import numpy as np
import pytest

from fixture_io import Replay, expected_tree, list_fixtures, load_fixture
from parity import RTOL, assert_same_array, f16_ulps

pytestmark = pytest.mark.gpu


def _imports():
    from nbody_orbit_analysis_b200 import h5shim, storage, track_orbits
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    return h5shim, storage, track_orbits, OrbitTracker, SynthSim, oracle


def compare_track_trees(got, exp, data_f64, derived_bulk=False):
    """File-level parity of a track_orbits result."""
    assert sorted(got) == sorted(exp), (
        sorted(set(got) - set(exp)), sorted(set(exp) - set(got)))
    n_ang = n_ang_off = 0
    worst = 0.0
    for k in sorted(exp):
        g, e = np.asarray(got[k]), np.asarray(exp[k])
        if e.dtype.kind in 'US':
            assert str(g) == str(e), k
        elif k.endswith('/angles'):
            assert g.dtype == e.dtype == np.float16, k
            assert g.shape == e.shape, (k, g.shape, e.shape)
            u = f16_ulps(g, e)
            n_ang += u.size
            n_ang_off += int(np.sum(u > 0))
            if u.size:
                worst = max(worst, float(u.max()))
        elif k.endswith('/bulk_velocities') and derived_bulk:
            # derived on the device in numpy's summation order: bit-identical
            assert g.dtype == e.dtype and g.shape == e.shape, k
            assert np.array_equal(g, e, equal_nan=True), k
        elif e.dtype.kind == 'f':
            assert_same_array(k, g, e, exact_float=False,
                              rtol=RTOL['float64' if data_f64 else 'float32'])
        else:
            assert_same_array(k, g, e)
    # float16 angles: identical up to arccos rounding -- at most 1 ulp apart
    # except downstream of a 1-ulp difference in an accumulated angle
    assert worst <= 2, 'float16 angles differ by %g ulp' % worst
    assert n_ang_off <= max(2, 0.02 * n_ang), \
        '%d of %d float16 angles differ' % (n_ang_off, n_ang)


@pytest.mark.parametrize('name', list_fixtures('track_'))
def test_track_orbits_matches_reference_fixture(name, tmp_path):
    h5shim, storage, track_orbits, _, _, _ = _imports()
    fx = load_fixture(name)
    meta = fx['meta']
    rp = Replay(fx)
    savefile = str(tmp_path / 'g.h5')
    snaps, mb = fx['in/snapshot_numbers'], fx['in/main_branches']
    run = track_orbits.track_orbits
    if meta['resume_at'] is None:
        run(snaps, mb, rp.regions, rp.load_snapshot_data, savefile,
            mode=meta['mode'], checkpoint=meta['checkpoint'], verbose=False)
    else:
        k = meta['resume_at']
        run(snaps[:k], mb[:k], rp.regions, rp.load_snapshot_data, savefile,
            mode=meta['mode'], checkpoint=True, verbose=False)
        run(snaps, mb, rp.regions, rp.load_snapshot_data, savefile,
            mode=meta['mode'], checkpoint=True, resume=True, verbose=False)
    got = storage.tree(savefile)
    if meta['checkpoint'] or meta['resume_at'] is not None:
        for k, v in storage.tree(savefile + '.checkpoint').items():
            got['/__checkpoint__' + k] = v
    sim = meta['sim']
    compare_track_trees(
        got, expected_tree(fx), data_f64=sim['dtype'] == 'float64',
        derived_bulk=not sim.get('catalogue_bulk', True))


@pytest.mark.parametrize('name', list_fixtures('kernels_'))
def test_frame_and_match_against_reference_vectors(name):
    """region_frame / compare_radial_velocities / calc_angles vectors."""
    _, _, _, OrbitTracker, _, _ = _imports()
    fx = load_fixture(name)
    data_f64 = fx['in/cur/coordinates'].dtype == np.float64
    rtol = RTOL['float64' if data_f64 else 'float32']
    for mode in ('pericentric', 'apocentric'):
        trk = OrbitTracker(mode=mode)
        out = {}
        for tag in ('prev', 'cur'):
            n = len(fx['in/%s/ids' % tag])
            snap = {'coordinates': fx['in/%s/coordinates' % tag],
                    'velocities': fx['in/%s/velocities' % tag],
                    'ids': fx['in/%s/ids' % tag], 'masses': 1.0,
                    'region_offsets': np.array([0]),
                    'box_size': float(fx['in/box_size']),
                    'redshift': float(fx['in/redshift'])}
            if tag == 'cur':
                trk.load_angles(fx['in/angles_prev_' + mode])
            res = trk.step(snap, np.array([0]), fx['in/centre'][None, :],
                           fx['in/bulk'][None, :], H=fx['in/H'][()],
                           want_angles=True, diagnostics=True)
            rhat = res.diag['rhat'].cpu().numpy().reshape(n, 3)
            vr = res.diag['vr'].cpu().numpy()
            assert rhat.dtype == fx['out/%s/rhat' % tag].dtype
            assert vr.dtype == fx['out/%s/vr' % tag].dtype
            assert np.allclose(rhat, fx['out/%s/rhat' % tag], rtol=rtol,
                               atol=1e-7 if not data_f64 else 1e-15)
            assert np.allclose(vr, fx['out/%s/vr' % tag], rtol=rtol,
                               atol=rtol * np.abs(fx['out/%s/vr' % tag]).max())
            out[tag] = res
        res = out['cur']
        exp = {k: fx['out/%s/%s' % (mode, k)] for k in (
            'apsis_inds', 'apsis_ids', 'inds_match', 'inds_departed',
            'angles', 'apsis_angles')}
        match = res.diag['match'].cpu().numpy()
        n_prev = len(fx['in/prev/ids'])
        survivors = np.setdiff1d(np.arange(n_prev), exp['inds_departed'])
        # prev particle survivors[k] must be matched to current inds_match[k]
        assert np.array_equal(match[exp['inds_match']], survivors)
        assert np.sum(match >= 0) == len(survivors)
        assert np.array_equal(res.apsis_ids, exp['apsis_ids'])
        assert np.array_equal(res.apsis_prev_index.cpu().numpy(),
                              survivors[exp['apsis_inds']])
        assert np.array_equal(res.apsis_offsets, [0, len(exp['apsis_ids'])])
        assert f16_ulps(res.apsis_angles, exp['apsis_angles']).max() <= 1
        assert f16_ulps(res.angles, exp['angles']).max() <= 1


CASES = [
    # (particles, halos, snaps, data dtype, catalogue dtype, kwargs)
    (60000, 37, 6, np.float32, np.float32, {}),
    (60000, 37, 6, np.float32, np.float64, {'late_halos': 0.3}),
    (40000, 11, 5, np.float64, np.float64, {'hubble': True}),
    (40000, 300, 5, np.float32, np.float32, {'hubble': True}),
    (30000, 5, 5, np.float32, np.float64, {'catalogue_bulk': False}),
    (30000, 5, 5, np.float64, np.float64,
     {'catalogue_bulk': False, 'mass_array': True}),
    (30000, 1, 5, np.float32, np.float32, {'nfw': True, 'periodic': False}),
    (5000, 2000, 4, np.float32, np.float32, {}),     # many tiny blocks
]


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
@pytest.mark.parametrize('case', CASES, ids=[
    'f32c32', 'f32c64_late', 'f64_hubble', 'f32_hubble_300h', 'f32_nobulk',
    'f64_massarr', 'nfw_nonperiodic', 'tiny_blocks'])
def test_track_orbits_matches_oracle(case, mode, tmp_path):
    h5shim, storage, track_orbits, _, SynthSim, oracle = _imports()
    n, nh, ns, dt, cdt, kw = case
    sim = SynthSim(n, nh, ns, dtype=dt, catalogue_dtype=cdt, **kw)
    f_gpu, f_cpu = str(tmp_path / 'gpu.h5'), str(tmp_path / 'cpu.h5')
    args = (sim.snapshot_numbers, sim.main_branches, sim.regions,
            sim.load_snapshot_data)
    track_orbits.track_orbits(*args, f_gpu, mode=mode, checkpoint=True,
                              verbose=False)
    oracle.track_orbits(*args, f_cpu, mode=mode, checkpoint=True,
                        storage=storage)
    got, exp = storage.tree(f_gpu), storage.tree(f_cpu)
    got['/ckpt'] = storage.tree(f_gpu + '.checkpoint')['/angles']
    exp['/ckpt'] = storage.tree(f_cpu + '.checkpoint')['/angles']
    n_events = sum(len(v) for k, v in exp.items() if k.endswith('er_IDs'))
    assert n_events > 0
    # the checkpoint is the whole per-particle float16 accumulator
    u = f16_ulps(got.pop('/ckpt'), exp.pop('/ckpt'))
    assert u.max() <= 2 and np.mean(u > 0) < 0.02
    compare_track_trees(got, exp, data_f64=dt == np.float64,
                        derived_bulk=not kw.get('catalogue_bulk', True))


def test_empty_block_and_vanishing_halo(tmp_path):
    """A halo may disappear and come back; a block may be empty."""
    h5shim, storage, track_orbits, _, SynthSim, oracle = _imports()
    sim = SynthSim(20000, 6, 6, dtype=np.float32, catalogue_dtype=np.float32)
    mb = sim.main_branches.copy()
    mb[2, 4] = -1          # halo 4 vanishes at t=2 and returns at t=3
    base_load = sim.load_snapshot_data

    def load(snap_no, pos, rad):
        s = base_load(snap_no, pos, rad)
        # empty the second block at every snapshot
        offs = np.append(s['region_offsets'], len(s['ids']))
        if len(offs) > 3:
            lo, hi = offs[1], offs[2]
            keep = np.ones(len(s['ids']), dtype=bool)
            keep[lo:hi] = False
            for k in ('ids', 'coordinates', 'velocities'):
                s[k] = s[k][keep]
            offs[2:] -= hi - lo
            s['region_offsets'] = offs[:-1]
        return s

    for f, fn, kw in ((str(tmp_path / 'g.h5'), track_orbits.track_orbits, {}),
                      (str(tmp_path / 'c.h5'), oracle.track_orbits,
                       {'storage': storage})):
        fn(sim.snapshot_numbers, mb, sim.regions, load, f, verbose=False, **kw)
    compare_track_trees(storage.tree(str(tmp_path / 'g.h5')),
                        storage.tree(str(tmp_path / 'c.h5')), data_f64=False)


def test_argument_errors_come_first():
    _, _, track_orbits, _, _, _ = _imports()
    with pytest.raises(ValueError):
        track_orbits.track_orbits([1, 2], [[1]], None, None, 'x', verbose=False)
    with pytest.raises(ValueError):
        track_orbits.track_orbits([1], [[1]], None, None, 'x', mode='both',
                                  verbose=False)


@pytest.mark.parametrize('world', [2, 3])
def test_id_sharding_emulated_on_one_gpu(world, tmp_path):
    """The multi-GPU decomposition without NCCL: `world` trackers on one GPU,
    each fed the particles with id mod world == rank; their local event lists
    (left in HBM) are concatenated in rank order and merged by order key with
    oa_merge_event_lists -- the result must equal the unsharded oracle run."""
    import ctypes as C
    import torch
    from nbody_orbit_analysis_b200 import sharded, _lib
    _, _, _, OrbitTracker, SynthSim, oracle = _imports()
    lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
    sim = SynthSim(50000, 17, 5, dtype=np.float32, catalogue_dtype=np.float32)
    exists = np.arange(17)
    trks = [OrbitTracker() for _ in range(world)]
    for trk in trks:
        trk.events_on_device = True
    prev = None
    dev = torch.device('cuda')
    n_events = 0
    for t, snap_no in enumerate(sim.snapshot_numbers):
        pos, rad, bulk = sim.regions(snap_no, sim.main_branches[t])
        full = sim.load_snapshot_data(snap_no, pos, rad)
        state, exp = oracle.track_snapshot(full, exists, pos, bulk, 0.0,
                                           'pericentric', prev)
        prev = state
        results = []
        for r, trk in enumerate(trks):
            local, gpos = sharded.shard_snapshot(full, r, world)
            results.append(trk.step(local, exists, pos, bulk, 0.0, gpos=gpos))
        if t == 0:
            continue
        keys, ids, angs, sizes = [], [], [], []
        counts = np.zeros(17, dtype=np.int64)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for res in results:
            E = res.n_events
            k = torch.empty(max(E, 1), dtype=torch.int64, device=dev)
            check(lib.oa_gather_i64(ptr(res.prev_gen.gpos),
                                    ptr(res.apsis_prev_index), E, None, ptr(k),
                                    st))
            keys.append(k[:E]); ids.append(res.d_ids); angs.append(res.d_ang)
            sizes.append(E)
            counts += np.diff(res.apsis_offsets)
        k_all, i_all, a_all = torch.cat(keys), torch.cat(ids), torch.cat(angs)
        total = int(k_all.numel())
        list_off = torch.tensor(np.concatenate(([0], np.cumsum(sizes))),
                                dtype=torch.int64, device=dev)
        ids_o = torch.empty(total, dtype=torch.int64, device=dev)
        ang_o = torch.empty(total, dtype=torch.int16, device=dev)
        check(lib.oa_merge_event_lists(ptr(k_all), ptr(i_all), ptr(a_all),
                                       total, ptr(list_off), world, ptr(ids_o),
                                       ptr(ang_o), st))
        assert np.array_equal(ids_o.cpu().numpy(), exp['apsis_ids'])
        assert np.array_equal(np.concatenate(([0], np.cumsum(counts))),
                              exp['apsis_offsets'])
        u = f16_ulps(ang_o.cpu().numpy().view(np.float16), exp['apsis_angles'])
        assert u.size == 0 or u.max() <= 2
        n_events += total

        # the same through the sync-free path: packed send buffers (the
        # all-gather is emulated by concatenation) + oa_merge_gathered
        n_seg = len(results[0].apsis_offsets) - 1
        for cap, overflow in ((max(sizes) + 5, 0), (max(max(sizes) // 2, 1), 1)):
            nbytes = lib.oa_exchange_bytes(n_seg, cap)
            sends = []
            for res in results:
                send = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
                check(lib.oa_pack_events(
                    ptr(res.prev_gen.gpos), ptr(res.d_sel), ptr(res.d_ids_buf),
                    ptr(res.d_ang_buf), ptr(res.d_small), n_seg, cap, ptr(send),
                    st))
                sends.append(send)
            recv = torch.cat(sends)
            ids2 = torch.empty(world * cap, dtype=torch.int64, device=dev)
            ang2 = torch.empty(world * cap, dtype=torch.int16, device=dev)
            info = torch.empty(n_seg + 3 + world, dtype=torch.int64, device=dev)
            check(lib.oa_merge_gathered(ptr(recv), world, n_seg, cap, ptr(ids2),
                                        ptr(ang2), ptr(info), st))
            info = info.cpu().numpy()
            assert int(info[2 + n_seg + world]) == overflow
            assert list(info[2 + n_seg:2 + n_seg + world]) == sizes
            assert np.array_equal(info[1:2 + n_seg], exp['apsis_offsets'])
            if not overflow:
                assert int(info[0]) == total
                assert np.array_equal(ids2[:total].cpu().numpy(),
                                      exp['apsis_ids'])
                assert np.array_equal(ang2[:total].cpu().numpy(),
                                      ang_o.cpu().numpy())

        # key-range partitioned path: quantile proposals (all-gather emulated by
        # concatenation), per-rank send blocks, all-to-all emulated by a block
        # transpose, per-rank merge; the rank slices concatenated in rank order
        # are the global lists
        cap = max(sizes) // world + 64
        blk = lib.oa_exchange_bytes(0, cap)
        props, sends, cnts = [], [], []
        for res in results:
            prop = torch.zeros(max(world - 1, 1), dtype=torch.int64, device=dev)
            check(lib.oa_split_quantiles(ptr(res.prev_gen.gpos), ptr(res.d_sel),
                                         ptr(res.d_small), n_seg, world,
                                         ptr(prop), st))
            props.append(prop)
        prop_all = torch.cat(props)
        for res in results:
            send = torch.zeros(world * blk, dtype=torch.uint8, device=dev)
            cnt = torch.zeros(n_seg, dtype=torch.int64, device=dev)
            bnd = torch.zeros(world + 1, dtype=torch.int64, device=dev)
            check(lib.oa_pack_split(
                ptr(res.prev_gen.gpos), ptr(res.d_sel), ptr(res.d_ids_buf),
                ptr(res.d_ang_buf), ptr(res.d_small), n_seg, ptr(prop_all),
                world, cap, ptr(bnd), ptr(send), ptr(cnt), st))
            sends.append(send)
            cnts.append(cnt)
            assert int(bnd[-1]) == res.n_events
        got_ids, got_ang, slice_sizes = [], [], []
        for q in range(world):
            recv = torch.cat([sd[q * blk:(q + 1) * blk] for sd in sends])
            ids3 = torch.empty(world * cap, dtype=torch.int64, device=dev)
            ang3 = torch.empty(world * cap, dtype=torch.int16, device=dev)
            info3 = torch.empty(2, dtype=torch.int64, device=dev)
            check(lib.oa_merge_blocks(ptr(recv), world, cap, ptr(ids3),
                                      ptr(ang3), ptr(info3), st))
            sz, largest = (int(v) for v in info3.cpu().tolist())
            assert 0 <= largest <= cap          # no block was truncated
            slice_sizes.append(sz)
            got_ids.append(ids3[:sz])
            got_ang.append(ang3[:sz])
        assert sum(slice_sizes) == total
        # the key ranges are balanced (local quantiles are global quantiles)
        assert max(slice_sizes) <= 1.5 * total / world + 64
        assert np.array_equal(torch.cat(got_ids).cpu().numpy(), exp['apsis_ids'])
        assert np.array_equal(torch.cat(got_ang).cpu().numpy(),
                              ang_o.cpu().numpy())
        assert np.array_equal(
            np.concatenate(([0], np.cumsum(sum(c.cpu().numpy() for c in cnts)))),
            exp['apsis_offsets'])
    assert n_events > 0


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
def test_device_state_checkpoint_resume(mode, tmp_path):
    """``checkpoint='state'`` saves the carried device state
    (records + ID table) next to the reference's ``angles`` dataset; a resumed
    run continues at the NEXT snapshot -- its loader is never asked for the
    snapshots already saved -- and writes the same file as an uninterrupted
    run.  A plain ``checkpoint=True`` file still resumes the reference way."""
    _, storage, track_orbits, _, SynthSim, _ = _imports()
    sim = SynthSim(30000, 6, 7, dtype=np.float32, catalogue_dtype=np.float64,
                   late_halos=0.3)
    f_a, f_b = str(tmp_path / 'a.h5'), str(tmp_path / 'b.h5')
    args = (sim.regions, sim.load_snapshot_data)
    track_orbits.track_orbits(sim.snapshot_numbers, sim.main_branches, *args,
                              f_a, mode=mode, verbose=False)
    k = 4

    class Crash(Exception):
        pass

    def crashing(sn, pos, rad):           # the run dies while loading snapshot k
        if int(sn) == int(sim.snapshot_numbers[k]):
            raise Crash()
        return sim.load_snapshot_data(sn, pos, rad)
    with pytest.raises(Crash):
        track_orbits.track_orbits(sim.snapshot_numbers, sim.main_branches,
                                  sim.regions, crashing, f_b, mode=mode,
                                  checkpoint='state', verbose=False)
    saved = [int(key.split('/')[1].split('_')[1]) for key in storage.tree(f_b)
             if key.endswith('/halo_IDs')]
    last = max(saved)
    assert last < int(sim.snapshot_numbers[k])
    ck = storage.tree(f_b + '.checkpoint')
    assert '/angles' in ck and '/b200_state/rec' in ck
    asked = []

    def loader(sn, pos, rad):
        asked.append(int(sn))
        return sim.load_snapshot_data(sn, pos, rad)
    track_orbits.track_orbits(sim.snapshot_numbers, sim.main_branches,
                              sim.regions, loader, f_b, mode=mode,
                              checkpoint='state', resume=True, verbose=False)
    assert asked == [int(s) for s in sim.snapshot_numbers if int(s) > last]
    got, exp = storage.tree(f_b), storage.tree(f_a)
    assert set(got) == set(exp)
    for key in exp:
        assert np.array_equal(np.asarray(got[key]), np.asarray(exp[key]),
                              equal_nan=np.asarray(exp[key]).dtype.kind == 'f'), key


def test_two_slot_ring_memory_plan(tmp_path, monkeypatch):
    """``ring=2`` (previous + current generation only: the memory plan for
    snapshots that fill HBM, 158 instead of 237 bytes per region-particle) gives
    the same file through the pipelined driver."""
    _, storage, track_orbits, _, SynthSim, oracle = _imports()
    monkeypatch.setenv('OA_TRACKER_RING', '2')
    sim = SynthSim(40000, 8, 6, dtype=np.float32, catalogue_dtype=np.float32,
                   late_halos=0.3)
    f_gpu, f_cpu = str(tmp_path / 'gpu.h5'), str(tmp_path / 'cpu.h5')
    args = (sim.snapshot_numbers, sim.main_branches, sim.regions,
            sim.load_snapshot_data)
    track_orbits.track_orbits(*args, f_gpu, verbose=False)
    oracle.track_orbits(*args, f_cpu, storage=storage)
    compare_track_trees(storage.tree(f_gpu), storage.tree(f_cpu), data_f64=False)
