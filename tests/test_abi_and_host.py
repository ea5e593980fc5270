"""CPU-only tests: the C-ABI library loads and exports every declared symbol,
host-side logic (storage container, synthetic workloads, argument checks)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'orbit_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(oa_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from nbody_orbit_analysis_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), 'missing export ' + name
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib.oa_abi_version() == _lib.ABI_VERSION
    assert _lib.lib.oa_track_args_size() == C.sizeof(_lib.TrackArgs)
    assert _lib.lib.oa_record_bytes(0) == 32
    assert _lib.lib.oa_record_bytes(1) == 64
    # index bits hold every block-local index 0 .. n-1
    for n in (0, 1, 2, 3, 7, 8, 9, 1000, 2**20, 2**20 + 1):
        b = _lib.lib.oa_index_bits(n)
        assert (1 << b) >= n and (b == 1 or (1 << (b - 1)) < n)
    # bucket ranges of consecutive blocks never overlap
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 50, 2000)
    off = np.concatenate(([0], np.cumsum(lens)))
    begin = np.array([_lib.lib.oa_table_bucket_begin(int(off[j]), j)
                      for j in range(len(lens))])
    nb = lens // _lib.BUCKET_LOAD + 1
    assert np.all(begin[1:] >= begin[:-1] + nb[:-1])
    assert np.array_equal(begin, off[:-1] // _lib.BUCKET_LOAD + np.arange(
        len(lens)))
    n_buckets = _lib.lib.oa_table_buckets(int(off[-1]), len(lens))
    assert begin[-1] + nb[-1] <= n_buckets
    assert _lib.lib.oa_table_slots(int(off[-1]), len(lens)) >= 9 * n_buckets


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from nbody_orbit_analysis_b200 import _lib
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    from nbody_orbit_analysis_b200.track_orbits import track_orbits
    with pytest.raises(_lib.OrbitB200Error):
        OrbitTracker()
    with pytest.raises(_lib.OrbitB200Error):
        track_orbits([1], [[5]], None, None, 'x', verbose=False)
    sm = C.c_int()
    assert _lib.lib.oa_device_info(C.byref(sm), None, None, None, None) < 0
    assert _lib.lib.oa_last_error()


def test_argument_validation_precedes_everything():
    from nbody_orbit_analysis_b200.track_orbits import track_orbits
    with pytest.raises(ValueError):
        track_orbits([1, 2], [[1]], None, None, 'x', verbose=False)
    with pytest.raises(ValueError):
        track_orbits([1], [[1]], None, None, 'x', mode='both', verbose=False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'nbody_orbit_analysis_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, f)).read()
                assert 'import oracle' not in text and \
                    'from oracle' not in text, f


def test_shim_container_roundtrip(tmp_path):
    from nbody_orbit_analysis_b200 import h5shim
    f = str(tmp_path / 'a.h5')
    with h5shim.File(f, 'w') as hf:
        hf.attrs['mode'] = 'pericentric'
        hf.attrs['box_size'] = 100.0
    with h5shim.File(f, 'r+') as hf:
        g = hf.create_group('snapshot_012')
        g.create_dataset('angles', data=np.arange(5, dtype=np.float16))
        g.create_dataset('empty', data=np.array([], dtype=np.int64))
        g.create_dataset('pos', data=np.arange(12.).reshape(4, 3))
        with pytest.raises(ValueError):
            g.create_dataset('pos', data=np.zeros(1))
    with h5shim.File(f, 'a') as hf:
        hf.create_group('snapshot_003')
    with h5shim.File(f, 'r') as hf:
        assert list(hf.keys()) == ['snapshot_003', 'snapshot_012']
        assert hf.attrs['mode'] == 'pericentric'
        assert 'box_size' in hf.attrs and hf.attrs['box_size'] == 100.0
        g = hf['snapshot_012']
        assert g['angles'].dtype == np.float16 and len(g['angles']) == 5
        assert g['angles'][1:3].tolist() == [1.0, 2.0]
        assert g['pos'][:].shape == (4, 3) and g['pos'][2:4][0, 0] == 6.0
        assert len(g['empty']) == 0 and g['empty'][:].dtype == np.int64
        with pytest.raises(KeyError):
            hf['nope']
    with pytest.raises(FileNotFoundError):
        h5shim.File(str(tmp_path / 'missing.h5'), 'r')


def test_synthetic_workload_is_reproducible_and_has_apsides():
    from nbody_orbit_analysis_b200.synth import SynthSim
    a = SynthSim(5000, 7, 4, late_halos=0.4)
    b = SynthSim(5000, 7, 4, late_halos=0.4)
    assert np.array_equal(a.main_branches, b.main_branches)
    halo_ids = a.main_branches[2][a.main_branches[2] != -1]
    pa = a.regions(a.snapshot_numbers[2], halo_ids)
    sa = a.load_snapshot_data(a.snapshot_numbers[2], pa[0], pa[1])
    sb = b.load_snapshot_data(b.snapshot_numbers[2], pa[0], pa[1])
    for k in ('ids', 'coordinates', 'velocities', 'region_offsets'):
        assert np.array_equal(sa[k], sb[k])
    assert len(np.unique(sa['ids'])) == len(sa['ids'])
    assert sa['coordinates'].min() >= 0 and sa['coordinates'].max() <= a.box
    # block order differs from snapshot to snapshot (matching is non-trivial)
    pc = a.regions(a.snapshot_numbers[3], a.main_branches[3][
        a.main_branches[3] != -1])
    sc = a.load_snapshot_data(a.snapshot_numbers[3], pc[0], pc[1])
    common = np.intersect1d(sa['ids'][:200], sc['ids'])
    assert len(common) > 20
    assert not np.array_equal(
        sa['ids'][np.isin(sa['ids'], common)],
        sc['ids'][np.isin(sc['ids'], common)])


def test_region_rows_host_equals_numpy_assembly():
    """``oa_region_rows_host`` (what ``OrbitTracker.submit_device`` calls)
    against the numpy assembly of the region table it replaced: centres and
    bulk velocities in both float dtypes, matched / new / vanished halos
    (reference ``track_orbits.py:155-165``)."""
    from nbody_orbit_analysis_b200 import _lib
    rng = np.random.default_rng(9)
    lib = _lib.lib
    for trial in range(62):
        n_h = int(rng.integers(0, 50))
        n_p = int(rng.integers(0, 50))
        pool = 200
        if trial >= 60:               # large catalogues: the threaded row fill
            n_h, n_p, pool = 40000 + trial, 39000, 60000
        ids_cur = np.sort(rng.choice(pool, n_h, replace=False)).astype(np.int64)
        ids_prev = np.sort(rng.choice(pool, n_p, replace=False)).astype(np.int64)
        offsets = np.concatenate(([0], np.cumsum(rng.integers(0, 5000, n_h)))
                                 ).astype(np.int64)
        p_off = np.concatenate(([0], np.cumsum(rng.integers(0, 5000, n_p)))
                               ).astype(np.int64)
        p_buckets = (p_off[:-1] // _lib.BUCKET_LOAD + np.arange(n_p)).astype(np.int64)
        cdt = rng.choice([np.float32, np.float64])
        bdt = rng.choice([np.float32, np.float64])
        cen = rng.standard_normal((n_h, 3)).astype(cdt) * 50
        bulk = rng.standard_normal((n_h, 3)).astype(bdt) if trial % 3 else None

        # numpy assembly (the code this function replaced)
        rows = np.zeros(n_h, dtype=_lib.REGION_DTYPE)
        rows['centre'] = cen.astype(np.float64)
        rows['centre_f'] = rows['centre']
        rows['cur_begin'] = offsets[:-1]
        rows['cur_count'] = np.diff(offsets)
        rows['prev_begin'] = -1
        matched = np.zeros(n_h, dtype=bool)
        if n_h and n_p:
            pos_c = np.minimum(np.searchsorted(ids_prev, ids_cur), n_p - 1)
            matched = ids_prev[pos_c] == ids_cur
            k = pos_c[matched]
            rows['prev_begin'][matched] = p_off[k]
            rows['prev_count'][matched] = p_off[k + 1] - p_off[k]
            rows['prev_bucket'][matched] = p_buckets[k]
        buckets = offsets[:-1] // _lib.BUCKET_LOAD + np.arange(n_h)
        rows['cur_bucket'] = buckets
        if bulk is not None:
            rows['bulk'] = bulk.astype(np.float64)
            rows['bulk_f'] = rows['bulk']

        got = np.full(n_h, 0x33, dtype=np.uint8).repeat(128).view(_lib.REGION_DTYPE)
        g_buckets = np.empty(n_h, dtype=np.int64)
        g_matched = np.empty(n_h, dtype=np.bool_)
        g_prev = np.empty(n_h, dtype=np.int32)
        g_seg = np.empty(max(n_h, 1), dtype=np.int64)
        n_m = C.c_int(-1)
        rc = lib.oa_region_rows_host(
            n_h, offsets.ctypes.data, cen.ctypes.data, _lib.dtype_code(cdt),
            bulk.ctypes.data if bulk is not None else None,
            _lib.dtype_code(bdt), ids_cur.ctypes.data,
            ids_prev.ctypes.data if n_p else None, n_p,
            p_off.ctypes.data if n_p else None,
            p_buckets.ctypes.data if n_p else None, got.ctypes.data,
            g_buckets.ctypes.data, g_matched.ctypes.data, g_prev.ctypes.data,
            g_seg.ctypes.data, C.byref(n_m))
        assert rc == 0
        assert got.tobytes() == rows.tobytes()
        assert np.array_equal(g_buckets, buckets)
        assert np.array_equal(g_matched, matched)
        assert n_m.value == int(matched.sum())
        assert np.array_equal(g_seg[:n_m.value], rows['prev_begin'][matched])
        exp_prev = np.full(n_h, -1, dtype=np.int32)
        if n_h and n_p:
            exp_prev[matched] = pos_c[matched]
        assert np.array_equal(g_prev, exp_prev)


def test_pairwise_sum_equals_numpy():
    """The mass sum of the derived bulk velocity (``np.sum`` of a contiguous
    1-D slice, ``track_orbits.py:270-274``) is numpy's pairwise sum: the routine
    the device kernel uses (csrc/oa_bulk.cu, compiled for the host too) must
    return numpy's value bit for bit at every length class of the algorithm."""
    import ctypes as C
    from nbody_orbit_analysis_b200 import _lib
    rng = np.random.default_rng(3)
    sizes = list(range(0, 280)) + [1000, 1023, 1024, 1025, 4095, 8191, 8192,
                                   8193, 16385, 70001, 200000, 1234567]
    for dt in (np.float32, np.float64):
        for n in sizes:
            a = (rng.uniform(0.5, 1.5, n) * rng.choice([-1, 1], n)).astype(dt)
            out = C.c_double()
            _lib.check(_lib.lib.oa_pairwise_sum_host(
                a.ctypes.data, _lib.dtype_code(dt), n, C.byref(out)))
            ref = np.sum(a) if n else dt(0)
            assert dt(out.value) == ref, (dt, n)


def test_axis0_reduction_is_sequential():
    """The other half of the derived bulk velocity: numpy reduces axis 0 of an
    (n, 3) array row after row in the array's dtype (what the device kernel's
    three chains do) -- pinned here so that a numpy that changes its order is
    noticed."""
    rng = np.random.default_rng(4)
    for dt in (np.float32, np.float64):
        for n in (1, 9, 129, 8193, 50000):
            v = (rng.normal(0, 100, (n + 3, 3)) + 30).astype(dt)[3:]
            acc = v[0].copy()
            for row in v[1:]:
                acc = acc + row
            assert np.array_equal(np.add.reduce(v, axis=0), acc)
            assert np.array_equal(np.mean(v, axis=0), acc / dt(n))


def test_pj2_plan_invariants():
    """``oa_pj2_plan_host`` (host side of the second partitioned join): partition
    counts are powers of two that never shrink for a halo, capacities cover the
    mean fill with the Poisson head room, record slots / fill entries are laid
    out back to back, one JOIN item per current partition of a halo that has a
    previous block, and the ticket ranges hold every tile exactly once."""
    from nbody_orbit_analysis_b200 import pj2, _lib
    plan = pj2.Planner(_lib.lib)
    rng = np.random.default_rng(2)
    n_h = 300
    lens = rng.integers(0, 40000, n_h)
    lens[[3, 17]] = 0
    lens[5] = 900000
    prev = [np.zeros(n_h, dtype=np.uint32) for _ in range(4)]
    for step in range(4):
        off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        p = plan(off, *prev, group_particles=1 << 17)
        P, cap = p.P.astype(np.int64), p.cap.astype(np.int64)
        assert ((P & (P - 1)) == 0).all() and (P >= 1).all()
        assert (P >= prev[0]).all()                               # never shrinks
        mean = -(-lens // P)
        assert (mean <= pj2.TARGET).all() or (P[mean > pj2.TARGET] ==
                                              prev[0][mean > pj2.TARGET]).all()
        assert (cap >= np.minimum(mean + 16, pj2.CAP)).all() and (cap <= pj2.CAP).all()
        assert np.array_equal(p.pb, np.concatenate(([0], np.cumsum(P)))[:-1])
        assert np.array_equal(p.base, np.concatenate(([0], np.cumsum(P * cap)))[:-1])
        assert p.n_entries == P.sum() and p.n_slots == (P * cap).sum()
        rows = p.rows
        has_prev = prev[0] > 0
        joins = np.where(has_prev & (lens > 0), P, 0)
        assert np.array_equal(rows['join_first'],
                              np.concatenate(([0], np.cumsum(joins))))
        assert (rows['P_cur'][:n_h][has_prev] ==
                prev[0][has_prev].astype(np.int64) << rows['shift'][:n_h][has_prev]).all()
        # groups partition the regions; ticket ranges: tiles once, joins once
        gf, go, rs = p.group_first, p.group_off, p.range_start
        assert gf[0] == 0 and gf[-1] == n_h and (np.diff(gf.astype(np.int64)) > 0).all()
        assert np.array_equal(go, off[gf])
        n_tiles = -(-int(off[-1]) // pj2.TILE)
        tiles = sum(int(rs[2 * s + 1]) - int(rs[2 * s]) for s in range(p.n_groups))
        assert tiles == n_tiles
        assert p.total == n_tiles + int(joins.sum())
        assert p.n_ranges == 2 * (p.n_groups + pj2.LAG) and rs[-1] == p.total
        prev = [p.P.copy(), p.cap.copy(), p.base.copy(), p.pb.copy()]
        lens = (lens * rng.uniform(0.7, 1.6, n_h)).astype(np.int64)


@pytest.mark.parametrize('nbytes', [0, 1, 4095, (8 << 20) - 1, (16 << 20) + 17,
                                    (24 << 20) + 4097, (70 << 20) + 3])
@pytest.mark.parametrize('threads', [0, 1, 3, 8, 100])
def test_host_copy_is_a_memcpy(nbytes, threads):
    """``oa_host_copy`` (the staging copy of pageable loader arrays into the
    pinned ingest ring): every byte, for part sizes that do and do not divide
    the length, unaligned ends, more threads than parts; bytes around the
    destination untouched."""
    from nbody_orbit_analysis_b200._lib import lib
    rng = np.random.default_rng(nbytes + threads)
    src = rng.integers(0, 256, nbytes + 64, dtype=np.uint8)
    dst = np.full(nbytes + 128, 0xA5, dtype=np.uint8)
    rc = lib.oa_host_copy(dst.ctypes.data + 32, src.ctypes.data + 5, nbytes, threads)
    assert rc == 0
    assert np.array_equal(dst[32:32 + nbytes], src[5:5 + nbytes])
    assert (dst[:32] == 0xA5).all() and (dst[32 + nbytes:] == 0xA5).all()


def test_stage_copy_paths(monkeypatch):
    """tracker._stage_copy: small copies through torch, large ones through
    ``oa_host_copy`` with the thread count of ``stage_threads()``."""
    import torch
    from nbody_orbit_analysis_b200 import tracker
    monkeypatch.setattr(tracker, 'STAGE_MIN_BYTES', 1 << 16)
    monkeypatch.setattr(tracker, '_stage_threads', None)
    monkeypatch.setenv('OA_STAGE_THREADS', '3')
    assert tracker.stage_threads() == 3
    for n in (100, 1 << 14, (1 << 20) + 13):
        src = torch.arange(n, dtype=torch.int64) * 7 - 3
        dst = torch.zeros(n, dtype=torch.int64)
        tracker._stage_copy(dst, src)
        assert torch.equal(dst, src)
    monkeypatch.setattr(tracker, '_stage_threads', None)
    monkeypatch.delenv('OA_STAGE_THREADS')
    monkeypatch.setenv('LOCAL_WORLD_SIZE', '1000')
    assert tracker.stage_threads() == 1
    monkeypatch.setattr(tracker, '_stage_threads', None)


def test_writer_thread_can_be_switched_off(monkeypatch):
    import threading
    from nbody_orbit_analysis_b200.track_orbits import _Writer
    monkeypatch.setenv('OA_WRITER_THREAD', '0')
    seen = []
    w = _Writer()
    w.submit(lambda: seen.append(threading.current_thread() is
                                 threading.main_thread()))
    assert seen == [True]
    with pytest.raises(OSError):
        w.submit(lambda: (_ for _ in ()).throw(OSError('x')))
    w.wait()


def test_writer_thread_order_and_errors():
    """track_orbits._Writer: writes run one at a time in submission order; an
    exception raised by a write surfaces on the caller's thread at the next
    submit / wait; ``wait(swallow=True)`` (used while another error unwinds)
    joins without raising."""
    import threading
    import time
    from nbody_orbit_analysis_b200.track_orbits import _Writer
    w, log = _Writer(), []

    def job(k, delay):
        assert threading.current_thread().name == 'orbit-b200-writer'
        time.sleep(delay)
        log.append(k)
    for k, d in enumerate((0.05, 0.0, 0.02, 0.0)):
        w.submit(job, k, d)
    w.wait()
    assert log == [0, 1, 2, 3]

    def bad():
        raise OSError('disk full')
    w.submit(bad)
    with pytest.raises(OSError, match='disk full'):
        w.submit(job, 9, 0.0)            # surfaces before the next write starts
    assert log == [0, 1, 2, 3]
    w.submit(bad)
    w.wait(swallow=True)
    w.wait()                             # nothing left, nothing raised
