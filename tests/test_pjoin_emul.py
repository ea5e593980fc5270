"""Partitioned hash join (csrc/oa_pjoin_core.cuh) on the CPU.

The stage code that nvcc compiles into ``oa_pjoin_kernel`` is compiled here by
g++ with ``-DPJ_HOST_EMUL`` (tests/pjoin_emul/pjoin_emul.cpp): a CTA is 512 real
threads + a barrier, several CTAs run concurrently, tickets / dependency
counters / shared-memory phases are the device's.  Driven over the same
synthetic snapshots as the other parity tests, its events must equal the
oracle's (``track_orbits.py:147-217``) bit for bit, its carried state the
oracle's state.  The plan (partition bits, work items, ticket order) is
``nbody_orbit_analysis_b200/pjoin.py`` -- the code the GPU path uses.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from nbody_orbit_analysis_b200 import pjoin
from nbody_orbit_analysis_b200.synth import SynthSim
from oracle import orbit_oracle as oracle
from parity import f16_ulps

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'pjoin_emul', 'pjoin_emul.cpp')
CORE = os.path.join(os.path.dirname(HERE), 'nbody_orbit_analysis_b200', 'csrc',
                    'oa_pjoin_core.cuh')
HDR = os.path.join(os.path.dirname(HERE), 'include', 'orbit_b200.h')
LIB = os.path.join(HERE, 'pjoin_emul', 'libpjoin_emul.so')

REGION_DTYPE = np.dtype([
    ('centre', np.float64, (3,)), ('bulk', np.float64, (3,)),
    ('prev_begin', np.int64), ('prev_count', np.int64),
    ('prev_bucket', np.int64), ('cur_bucket', np.int64),
    ('cur_begin', np.int64), ('cur_count', np.int64),
    ('centre_f', np.float32, (3,)), ('bulk_f', np.float32, (3,)),
    ('reserved', np.int64)])
REC_DTYPE = np.dtype([('id', np.int64), ('rhat', np.float32, (3,)),
                      ('vr', np.float32), ('pos', np.uint32),
                      ('angle', np.uint16), ('flags', np.uint16)])
NO_EVENT = 0x8000


SMALL_SHAPE = ['-DOA_PJOIN_THREADS=256', '-DOA_PJOIN_MIN_CTAS=4',
               '-DOA_PJOIN_TILE=1024', '-DOA_PJOIN_REC_CAP=1408',
               '-DOA_PJOIN_TARGET=1152', '-DOA_PJOIN_TMA=1']


def build_emul(flags=(), tag=''):
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    out = LIB.replace('.so', tag + '.so')
    deps = [SRC, CORE, HDR]
    if not os.path.exists(out) or any(
            os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run([gxx, '-O1', '-ffp-contract=off', '-pthread', '-shared',
                        '-fPIC', '-std=c++17', '-DPJ_HOST_EMUL', *flags,
                        '-o', out, SRC], check=True)
    lib = C.CDLL(out)
    lib.pj_emul_step.argtypes = [C.POINTER(pjoin.PJoinArgs), C.c_int]
    lib.pj_emul_half_bits.argtypes = [C.c_float]
    lib.pj_emul_half_bits.restype = C.c_uint
    cfg = (C.c_int32 * 8)()
    assert lib.pj_emul_config(cfg) == C.sizeof(pjoin.PJoinArgs), \
        'PJoinArgs does not mirror oa_pjoin_args'
    lib.shape = list(cfg)
    return lib


def use_shape(monkeypatch, lib):
    """The plan must be made for the shape the (emulated) kernel was built with."""
    for name, v in zip(('THREADS', None, 'TILE', 'CTILE', 'REC_CAP', 'TARGET',
                        'MAX_BITS'), lib.shape):
        if name:
            monkeypatch.setattr(pjoin, name, v)


@pytest.fixture(scope='module')
def emul():
    lib = build_emul()
    assert lib.shape[:7] == [pjoin.THREADS, 2, pjoin.TILE, pjoin.CTILE,
                             pjoin.REC_CAP, pjoin.TARGET, pjoin.MAX_BITS]
    return lib


def _ptr(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


class Gen:
    pass


class EmulTracker:
    """Host twin of the partitioned-join branch of ``OrbitTracker``."""

    def __init__(self, lib, mode='pericentric', n_ctas=3):
        self.lib, self.mode, self.n_ctas = lib, mode, n_ctas
        self.prev = None

    def step(self, snap, halo_exists, centres, bulk, H=0.0, target=None,
             lag=None):
        ids = np.ascontiguousarray(snap['ids'], dtype=np.int64)
        pos = np.ascontiguousarray(snap['coordinates'], dtype=np.float32)
        vel = np.ascontiguousarray(snap['velocities'], dtype=np.float32)
        n, n_h = len(ids), len(halo_exists)
        offsets = np.append(np.asarray(snap['region_offsets'], np.int64), n)
        prev = self.prev
        rows = np.zeros(n_h, dtype=REGION_DTYPE)
        rows['centre'] = np.asarray(centres, np.float64).reshape(n_h, 3)
        rows['centre_f'] = rows['centre']
        rows['bulk'] = np.asarray(bulk, np.float64).reshape(n_h, 3)
        rows['bulk_f'] = rows['bulk']
        rows['cur_begin'], rows['cur_count'] = offsets[:-1], np.diff(offsets)
        rows['prev_begin'] = -1
        prev_bits = np.full(n_h, -1, dtype=np.int32)
        prev_pb = np.zeros(n_h, dtype=np.int64)
        prev_counts = np.zeros(n_h, dtype=np.int64)
        matched = np.zeros(n_h, dtype=bool)
        if prev is not None and n_h and len(prev.halo_exists):
            k = np.minimum(np.searchsorted(prev.halo_exists, halo_exists),
                           len(prev.halo_exists) - 1)
            matched = prev.halo_exists[k] == halo_exists
            km = k[matched]
            rows['prev_begin'][matched] = prev.offsets[km]
            rows['prev_count'][matched] = prev.offsets[km + 1] - prev.offsets[km]
            prev_bits[matched] = prev.bits[km]
            prev_pb[matched] = prev.pb[km]
            prev_counts[matched] = rows['prev_count'][matched]
        plan = pjoin.make_plan(offsets, prev_bits, prev_pb, prev_counts, target, lag)

        g = Gen()
        g.n, g.ids, g.offsets, g.halo_exists = n, ids, offsets, np.asarray(halo_exists)
        g.bits, g.pb = plan.bits, plan.pb
        g.rec = np.full(max(n, 1), 0x5A, dtype=np.uint8).repeat(32).view(REC_DTYPE)
        g.part_off = np.full(max(plan.n_entries, 1), 0xDEADBEEF, dtype=np.uint32)
        g.mark = np.full(max(n, 1), 0x1234, dtype=np.uint16)
        ws = np.full(2 * plan.total + 4 + 3 * n_h + plan.n_entries, 0x77,
                     dtype=np.uint32)

        a = pjoin.PJoinArgs()
        a.pos, a.vel, a.ids, a.n_cur = _ptr(pos), _ptr(vel), _ptr(ids), n
        a.regions, a.plan = _ptr(rows), _ptr(plan.rows)
        a.group_first, a.range_start = _ptr(plan.group_first), _ptr(plan.range_start)
        a.n_regions, a.n_groups, a.n_ranges = n_h, plan.n_groups, plan.n_ranges
        a.centre_f32 = int(np.asarray(centres).dtype == np.float32)
        a.bulk_f32 = int(np.asarray(bulk).dtype == np.float32)
        box = snap.get('box_size')
        a.periodic = int(box is not None)
        if box is not None:
            for q, v in enumerate(np.broadcast_to(np.asarray(box, np.float64), (3,))):
                a.box[q] = float(v)
        a.mode = {'pericentric': 0, 'apocentric': 1}[self.mode]
        a.hubble_on, a.hubble = int(H != 0.0), float(H)
        a.one_plus_z = 1 + float(snap.get('redshift', 0.0))
        if prev is not None:
            a.rec_prev, a.part_off_prev = _ptr(prev.rec), _ptr(prev.part_off)
            a.mark_prev, a.n_prev = _ptr(prev.mark), prev.n
        a.rec_cur, a.part_off_cur, a.mark_cur = _ptr(g.rec), _ptr(g.part_off), _ptr(g.mark)
        a.workspace, a.workspace_bytes = _ptr(ws), ws.nbytes
        a.n_part_entries, a.total_tickets = plan.n_entries, plan.total
        assert self.lib.pj_emul_step(C.byref(a), self.n_ctas) == 0
        # the item list written by the expand pre-kernel is the plan's ticket order
        items = ws[:2 * plan.total].view(np.uint64)
        for tkt in range(0, plan.total, max(1, plan.total // 200)):
            st_, j_, i_ = pjoin.decode(plan, tkt)
            assert int(items[tkt]) == (st_ << 62) | (j_ << 32) | i_

        out = None
        if prev is not None:
            sel = np.flatnonzero(prev.mark[:prev.n] != NO_EVENT)
            seg = rows['prev_begin'][matched]
            out = {'apsis_ids': prev.ids[sel],
                   'apsis_angles': prev.mark[sel].view(np.float16),
                   'apsis_offsets': np.append(np.searchsorted(sel, seg), len(sel))}
        self.prev, self.plan, self.rows = g, plan, rows
        return g, out


def check_state(g, plan, state, mode_rows):
    """Carried records against the oracle's state (block order via rec.pos)."""
    n = g.n
    rec = g.rec[:n]
    # every block position exactly once, inside its own region's range and partition
    assert np.array_equal(np.sort(rec['pos']), np.arange(n, dtype=np.uint32))
    where = np.empty(n, dtype=np.int64)
    where[rec['pos']] = np.arange(n)
    for j in range(len(g.offsets) - 1):
        lo, hi = g.offsets[j], g.offsets[j + 1]
        w = where[lo:hi]
        assert np.all((w >= lo) & (w < hi))
        b, pb = int(plan.bits[j]), int(plan.pb[j])
        po = g.part_off[pb:pb + (1 << b) + 1].astype(np.int64)
        assert po[0] == lo and po[-1] == hi and np.all(np.diff(po) >= 0)
    by_pos = rec[where]
    assert np.array_equal(by_pos['id'], g.ids)
    assert np.array_equal(by_pos['rhat'], state.rhats.astype(np.float32))
    assert np.array_equal(np.sign(by_pos['vr']),
                          np.sign(state.radial_vels).astype(np.float32))
    assert np.all(g.mark[:n] == NO_EVENT)
    ulps = f16_ulps(by_pos['angle'].view(np.float16), state.angles)
    assert ulps.max() <= 2 and np.mean(ulps > 0) < 0.02


def run_case(lib, sim, mode='pericentric', targets=None, n_ctas=3, lag=1 << 12,
             hubble=False):
    trk = EmulTracker(lib, mode, n_ctas)
    prev_state = None
    n_events = 0
    stats = {'bits': [], 'tickets': [], 'maxlen': [], 'packs': []}
    for t, sn in enumerate(sim.snapshot_numbers):
        exists = np.flatnonzero(np.asarray(sim.main_branches[t]) >= 0) \
            if hasattr(sim, 'main_branches') else np.arange(sim.n_halos)
        halo_ids = np.asarray(sim.main_branches[t])[exists]
        pos, rad, bulk = sim.regions(sn, halo_ids)
        snap = sim.load_snapshot_data(sn, pos, rad)
        H = 0.0
        if hubble:
            from nbody_orbit_analysis_b200.utils import hubble_parameter
            H = hubble_parameter(snap['redshift'], snap['H0'], snap['Omega_m'],
                                 snap['Omega_L'], snap.get('Omega_k', 0))
        with np.errstate(all='ignore'):
            state, exp = oracle.track_snapshot(snap, exists, pos, bulk, H, mode,
                                               prev_state)
        target = targets[t % len(targets)] if targets else None
        g, out = trk.step(snap, exists, pos, bulk, H, target, lag)
        check_state(g, trk.plan, state, trk.rows)
        if exp is not None:
            assert np.array_equal(out['apsis_ids'], exp['apsis_ids'])
            assert np.array_equal(out['apsis_offsets'], exp['apsis_offsets'])
            ulps = f16_ulps(out['apsis_angles'], exp['apsis_angles'])
            assert ulps.size == 0 or (ulps.max() <= 2 and np.mean(ulps > 0) < 0.02)
            n_events += len(exp['apsis_ids'])
        stats['bits'].append(int(trk.plan.bits.max()) if len(trk.plan.bits) else 0)
        stats['tickets'].append(trk.plan.total)
        stats['maxlen'].append(int(np.diff(g.offsets).max()))
        stats['packs'].append(int(trk.plan.rows['pack_len'].max()))
        prev_state = state
    assert n_events > 0
    return stats


def test_half_conversions_match_numpy(emul):
    rng = np.random.default_rng(0)
    x = np.concatenate((rng.standard_normal(2000) * 3, [0.0, 65504.0, 1e-8, 7e4,
                                                        np.pi, 6.1e-5, 5.9e-8]))
    for v in x.astype(np.float32):
        assert emul.pj_emul_half_bits(float(v)) == int(
            np.float32(v).astype(np.float16).view(np.uint16))


def test_plan_decode_covers_every_item_once():
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 9000, 60)
    lens[7] = 0
    lens[20] = 70000
    offsets = np.concatenate(([0], np.cumsum(lens)))
    prev_bits = rng.integers(-1, 3, 60).astype(np.int32)
    prev_counts = np.where(prev_bits >= 0, rng.integers(0, 3000, 60), 0)
    plan = pjoin.make_plan(offsets, prev_bits, np.zeros(60, np.int64), prev_counts,
                           2304, 1 << 14)
    seen = {}
    last_of = {}
    for tkt in range(plan.total):
        st, j, i = pjoin.decode(plan, tkt)
        assert (st, j, i) not in seen
        seen[(st, j, i)] = tkt
        last_of[(st, j)] = tkt
    bits = plan.bits
    for j in range(60):
        tiles = -(-int(lens[j]) // pjoin.TILE) if bits[j] > 0 else 0
        ctiles = -(-int(lens[j]) // pjoin.CTILE) if bits[j] > 0 else 0
        joins = (1 << max(int(prev_bits[j]), 0) if prev_bits[j] >= 0 else 0) \
            if bits[j] > 0 else int(plan.rows['pack_len'][j] > 0)
        assert all((pjoin.COUNT, j, i) in seen for i in range(ctiles))
        assert all((pjoin.SCATTER, j, i) in seen for i in range(tiles))
        assert ((pjoin.SCAN, j, 0) in seen) == (bits[j] > 0)
        assert all((pjoin.JOIN, j, i) in seen for i in range(joins))
        # an item only waits for items with smaller tickets
        if tiles:
            assert last_of[(pjoin.COUNT, j)] < seen[(pjoin.SCAN, j, 0)]
            assert seen[(pjoin.SCAN, j, 0)] < seen[(pjoin.SCATTER, j, 0)]
            if joins:
                assert last_of[(pjoin.SCATTER, j)] < seen[(pjoin.JOIN, j, 0)]
    n_items = sum(1 for _ in seen)
    assert n_items == plan.total


def test_c_plan_equals_numpy_plan():
    """``oa_pjoin_plan_host`` (what the tracker calls) against ``make_plan``."""
    from nbody_orbit_analysis_b200 import _lib
    rng = np.random.default_rng(5)
    planner = pjoin.Planner(_lib.lib)
    for trial in range(120):
        nh = int(rng.integers(0, 300))
        lens = rng.integers(0, 30000, nh)
        if nh > 3:
            lens[rng.integers(0, nh, 2)] = [0, 900000]
        offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        prev_bits = rng.integers(-1, 6, nh).astype(np.int32)
        prev_pb = rng.integers(0, 10 ** 6, nh).astype(np.int64)
        target = int(rng.choice([300, 2304, 5000]))
        lag = int(rng.choice([1 << 12, 1 << 19]))
        prev_counts = rng.integers(0, 4000, nh).astype(np.int64)
        if trial % 4 == 0:
            lens = rng.integers(0, 400, nh)          # many tiny regions: packs
            offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
            prev_counts = rng.integers(0, 400, nh).astype(np.int64)
        a = pjoin.make_plan(offsets, prev_bits, prev_pb, prev_counts, target, lag)
        b = planner(offsets, prev_bits, prev_pb, prev_counts, target, lag)
        if trial % 4 == 0 and nh > 40:
            assert a.rows['pack_len'].max() > 1
        assert np.array_equal(a.rows, b.rows)
        assert np.array_equal(a.group_first, b.group_first)
        assert np.array_equal(a.range_start, b.range_start)
        assert np.array_equal(a.bits, b.bits) and np.array_equal(a.pb, b.pb)
        assert (a.n_entries, a.n_groups, a.n_ranges, a.total) == \
            (b.n_entries, b.n_groups, b.n_ranges, b.total)


def test_small_regions_direct_join(emul):
    """Every region is one partition (bits 0): frame + join in one item."""
    sim = SynthSim(12000, 12, 5, dtype=np.float32, catalogue_dtype=np.float32)
    st = run_case(emul, sim)
    assert max(st['bits']) == 0 and max(st['maxlen']) < pjoin.REC_CAP


@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
def test_partitioned_regions(emul, mode):
    """Partition target lowered to 300 particles: COUNT / SCAN / SCATTER / JOIN
    with up to 2^5 partitions per region, small regions mixed in."""
    sim = SynthSim(40000, 9, 4, dtype=np.float32, catalogue_dtype=np.float32)
    st = run_case(emul, sim, mode, targets=[300])
    assert max(st['bits']) >= 3


def test_partition_count_grows_between_snapshots(emul):
    """bits_cur > bits_prev: a previous partition joins 2 or 4 current ones."""
    sim = SynthSim(50000, 6, 6, dtype=np.float32, catalogue_dtype=np.float32)
    st = run_case(emul, sim, targets=[4000, 4000, 1000, 1000, 250, 250])
    assert st['bits'][0] < st['bits'][2] < st['bits'][4]


def test_previous_partition_larger_than_the_table(emul):
    """A previous block of > REC_CAP records in ONE partition is joined in
    several table batches (matched records are skipped by later batches)."""
    sim = SynthSim(40000, 3, 4, dtype=np.float32, catalogue_dtype=np.float32)
    st = run_case(emul, sim, targets=[1 << 20])
    assert max(st['bits']) == 0 and max(st['maxlen']) > 2 * pjoin.REC_CAP


def test_many_ctas_and_one_cta(emul):
    sim = SynthSim(20000, 7, 4, dtype=np.float32, catalogue_dtype=np.float32)
    run_case(emul, sim, targets=[500], n_ctas=1)
    run_case(emul, sim, targets=[500], n_ctas=6, lag=1 << 10)


def test_no_data_race_under_tsan(emul, tmp_path):
    """The stage code under ThreadSanitizer (partitioned, growing partition
    count, multi-batch join; 2-3 concurrent CTAs of 512 threads): no report."""
    import sys
    gcc = shutil.which('gcc')
    tsan = subprocess.run([gcc, '-print-file-name=libtsan.so'],
                          capture_output=True, text=True).stdout.strip() \
        if gcc else ''
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip('libtsan not available')
    lib = str(tmp_path / 'libpjoin_emul_tsan.so')
    subprocess.run([shutil.which('g++'), '-O1', '-g', '-ffp-contract=off',
                    '-pthread', '-shared', '-fPIC', '-std=c++17',
                    '-DPJ_HOST_EMUL', '-fsanitize=thread', '-o', lib, SRC],
                   check=True)
    env = dict(os.environ, LD_PRELOAD=tsan,
               TSAN_OPTIONS='report_signal_unsafe=0 exitcode=0 halt_on_error=0')
    run = subprocess.run(
        [sys.executable, os.path.join(HERE, 'pjoin_emul', 'tsan_run.py'), lib],
        capture_output=True, text=True, env=env, timeout=1500)
    out = run.stdout + run.stderr
    if 'tsan-run-complete' not in out and 'ThreadSanitizer' not in out:
        pytest.skip('the interpreter does not run under libtsan here: %s'
                    % out[-300:])
    assert 'WARNING: ThreadSanitizer' not in out, out[-3000:]
    assert 'tsan-run-complete' in out, out[-3000:]


def test_small_cta_shape(monkeypatch):
    """The same stage code built as 256-thread CTAs with 1408-record tables
    (4 CTAs per SM on the device): a tuning build must not change results."""
    lib = build_emul(SMALL_SHAPE, '_small')
    assert lib.shape[:7] == [256, 4, 1024, 8192, 1408, 1152, 12]
    use_shape(monkeypatch, lib)
    sim = SynthSim(60000, 9, 4, dtype=np.float32, catalogue_dtype=np.float32)
    st = run_case(lib, sim, targets=[None], n_ctas=4, lag=1 << 13)
    assert max(st['bits']) >= 3 and max(st['maxlen']) > 2 * 1408
    sim = SynthSim(30000, 3, 3, dtype=np.float32, catalogue_dtype=np.float32)
    run_case(lib, sim, targets=[1 << 20], n_ctas=2)        # multi-batch joins


def test_packs_of_tiny_regions(emul):
    """Hundreds of halos with a few dozen particles each: consecutive small
    regions share one work item (one tile of inputs, one table holding the
    previous blocks of all members) -- with a particle allowed in two halos."""
    sim = SynthSim(9000, 300, 4, dtype=np.float32, catalogue_dtype=np.float32,
                   late_halos=0.2)
    st = run_case(emul, sim, lag=1 << 11)
    assert max(st['bits']) == 0
    assert max(st['packs']) >= 8 and max(st['tickets']) < 120
    # tiny and partitioned regions mixed
    sim = SynthSim(30000, 60, 4, dtype=np.float32, catalogue_dtype=np.float32)
    st = run_case(emul, sim, targets=[300], lag=1 << 12)
    assert max(st['bits']) >= 2 and max(st['packs']) >= 2


def test_pack_with_the_same_particles_in_two_halos(emul):
    """Overlapping regions: the same IDs sit in two blocks of one pack, so the
    shared table holds every ID twice -- the match must stay inside the halo."""
    sim = SynthSim(1400, 2, 4, dtype=np.float32, catalogue_dtype=np.float32)
    trk = EmulTracker(emul)
    prev_state = None
    n_events = 0
    for t, sn in enumerate(sim.snapshot_numbers):
        pos, rad, bulk = sim.regions(sn, sim.main_branches[t])
        s = sim.load_snapshot_data(sn, pos, rad)
        n0 = int(s['region_offsets'][1])
        # blocks: [halo 0 | halo 1 | halo 0 again, seen from a shifted centre and
        # in reversed order]
        dup = slice(0, n0)
        snap = dict(s)
        for key in ('ids', 'coordinates', 'velocities'):
            snap[key] = np.concatenate((s[key], s[key][dup][::-1]))
        snap['region_offsets'] = np.append(s['region_offsets'], len(s['ids']))
        pos3 = np.vstack((pos, pos[0] + np.float32(0.01)))
        bulk3 = np.vstack((bulk, bulk[0] * np.float32(0.5)))
        exists = np.arange(3)
        with np.errstate(all='ignore'):
            state, exp = oracle.track_snapshot(snap, exists, pos3, bulk3, 0.0,
                                               'pericentric', prev_state)
        g, out = trk.step(snap, exists, pos3, bulk3, 0.0, lag=1 << 16)
        assert trk.plan.rows['pack_len'][0] == 3
        check_state(g, trk.plan, state, trk.rows)
        if exp is not None:
            assert np.array_equal(out['apsis_ids'], exp['apsis_ids'])
            assert np.array_equal(out['apsis_offsets'], exp['apsis_offsets'])
            n_events += len(exp['apsis_ids'])
        prev_state = state
    assert n_events > 0
