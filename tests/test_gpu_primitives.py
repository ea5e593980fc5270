"""GPU tests of the building blocks behind the C ABI: radix sort, ordered
selection, segmented bulk velocity, hash-table matching at awkward sizes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _env():
    import torch
    from nbody_orbit_analysis_b200 import _lib
    return torch, _lib, _lib.lib, _lib.ptr, _lib.check


def _dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


@pytest.mark.parametrize('n', [1, 31, 2047, 2048, 2049, 70001, 1 << 20])
@pytest.mark.parametrize('bits', [(0, 64), (0, 8), (3, 29), (40, 41), (7, 7)])
def test_radix_sort_pairs_is_a_stable_sort(n, bits):
    torch, _lib, lib, ptr, check = _env()
    rng = np.random.default_rng(n * 131 + bits[0])
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, n, dtype=np.uint64)
    if bits == (0, 8):
        keys = keys & np.uint64(0xFF)
    vals = np.arange(n, dtype=np.uint64)
    k_in, v_in = _dev(torch, keys.view(np.int64)), _dev(torch, vals.view(np.int64))
    k_out, v_out = torch.empty_like(k_in), torch.empty_like(v_in)
    ws_bytes = lib.oa_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    check(lib.oa_sort_pairs_u64(ptr(k_in), ptr(v_in), ptr(k_out), ptr(v_out), n,
                                bits[0], bits[1], ptr(ws), ws_bytes, None))
    torch.cuda.synchronize()
    width = bits[1] - bits[0]
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(2**64 - 1)
    digit = (keys >> np.uint64(bits[0])) & mask
    order = np.argsort(digit, kind='stable')
    assert np.array_equal(k_out.cpu().numpy().view(np.uint64), keys[order])
    assert np.array_equal(v_out.cpu().numpy().view(np.uint64), vals[order])
    # keys only
    k2 = torch.empty_like(k_in)
    check(lib.oa_sort_pairs_u64(ptr(k_in), None, ptr(k2), None, n, bits[0],
                                bits[1], ptr(ws), ws_bytes, None))
    torch.cuda.synchronize()
    assert np.array_equal(k2.cpu().numpy().view(np.uint64), keys[order])


def test_minmax():
    torch, _lib, lib, ptr, check = _env()
    rng = np.random.default_rng(5)
    x = rng.integers(-2**62, 2**62, 1234567, dtype=np.int64)
    out = torch.empty(2, dtype=torch.int64, device='cuda')
    check(lib.oa_minmax_i64(ptr(_dev(torch, x)), len(x), ptr(out), None))
    assert out.cpu().tolist() == [int(x.min()), int(x.max())]


@pytest.mark.parametrize('n', [0, 1, 7, 8, 2047, 2048, 2049, 300000, 3000017])
@pytest.mark.parametrize('density', [0.0, 0.03, 0.5, 1.0])
def test_ordered_select(n, density):
    torch, _lib, lib, ptr, check = _env()
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    trk = OrbitTracker()
    rng = np.random.default_rng(n + int(density * 100))
    marks = np.full(n + 8, 0x8000, dtype=np.uint16)
    hit = rng.random(n) < density
    marks[:n][hit] = rng.integers(0, 0x7C00, int(hit.sum()), dtype=np.uint16)
    d = _dev(torch, marks.view(np.int16))
    sel, total = trk.select(d, n, _lib.OA_SEL_NE, 0x8000)
    assert total == int(hit.sum())
    assert np.array_equal(sel[:total].cpu().numpy(), np.flatnonzero(hit))
    sel, total = trk.select(d, n, _lib.OA_SEL_EQ, 0x8000)
    assert np.array_equal(sel[:total].cpu().numpy(), np.flatnonzero(~hit))
    if n:
        begins = np.sort(rng.integers(0, n + 1, 17))
        offs = trk.segment_offsets(sel, total, begins)
        assert np.array_equal(
            offs, np.searchsorted(np.flatnonzero(~hit), begins, side='left'))


@pytest.mark.parametrize('dtype,mdtype', [(np.float32, np.float32),
                                          (np.float64, np.float64),
                                          (np.float32, np.float64),
                                          (np.float64, np.float32)])
@pytest.mark.parametrize('weighted', [False, True])
def test_bulk_velocity(dtype, mdtype, weighted):
    """Derived bulk velocity = the reference's own numpy expressions
    (track_orbits.py:270-280), BIT for bit: the kernel adds in numpy's order
    (axis-0 reduction row after row, mass sum pairwise)."""
    torch, _lib, lib, ptr, check = _env()
    rng = np.random.default_rng(11)
    lens = np.concatenate([[0, 1, 2, 5000, 0, 33, 9000, 4096, 4097, 1, 0, 7, 8,
                            9, 127, 128, 129, 300000],
                           rng.integers(0, 50, 400)])
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    n, nh = int(off[-1]), len(lens)
    vel = (rng.normal(0, 200, (n, 3)) + 40).astype(dtype)
    mass = rng.uniform(0.5, 2, n).astype(mdtype) if weighted else None
    rows = np.zeros(nh, dtype=_lib.REGION_DTYPE)
    d_rows = _dev(torch, rows.view(np.uint8))
    d_out = torch.empty(3 * nh, dtype=torch.float64, device='cuda')
    ws_bytes = lib.oa_bulk_workspace_bytes(n, nh)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    d_vel, d_off = _dev(torch, vel.reshape(-1)), _dev(torch, off)
    d_mass = _dev(torch, mass) if weighted else None
    code = 1 if dtype == np.float64 else 0
    mcode = 1 if mdtype == np.float64 else 0
    check(lib.oa_bulk_velocity(ptr(d_vel), code, ptr(d_mass), mcode, ptr(d_off),
                               nh, n, 0, ptr(d_rows), ptr(d_out), ptr(ws),
                               ws_bytes, None))
    got = d_out.cpu().numpy().reshape(nh, 3)
    rows_back = d_rows.cpu().numpy().view(_lib.REGION_DTYPE)
    with np.errstate(all='ignore'):
        for j in range(nh):
            v = vel[off[j]:off[j + 1]]
            if weighted:
                m = mass[off[j]:off[j + 1]]
                exp = np.sum(m[:, np.newaxis] * v, axis=0) / np.sum(m)
            else:
                exp = np.mean(v, axis=0)
            assert np.array_equal(got[j], exp.astype(np.float64),
                                  equal_nan=True), (j, got[j], exp)
            assert np.array_equal(rows_back['bulk'][j], got[j], equal_nan=True)
    # deterministic: a second run gives the same bits
    d_out2 = torch.empty_like(d_out)
    check(lib.oa_bulk_velocity(ptr(d_vel), code, ptr(d_mass), mcode, ptr(d_off),
                               nh, n, 0, ptr(d_rows), ptr(d_out2), ptr(ws),
                               ws_bytes, None))
    assert np.array_equal(d_out2.cpu().numpy(), d_out.cpu().numpy(),
                          equal_nan=True)


@pytest.mark.parametrize('id_kind', ['dense', 'sparse63', 'negative'])
def test_matching_is_exact_for_awkward_ids(id_kind):
    """Membership / match indices are bit-exact whatever the ID values are
    (hash collisions, fingerprints, 63-bit IDs, negative IDs)."""
    torch, _lib, lib, ptr, check = _env()
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    rng = np.random.default_rng(3)
    lens_prev = np.array([1, 0, 3, 100000, 17, 2, 0, 4000])
    lens_cur = np.array([1, 5, 0, 110000, 17, 3, 0, 3000])
    nh = len(lens_prev)

    def make_ids(total):
        if id_kind == 'dense':
            return rng.permutation(total * 2)[:total].astype(np.int64)
        if id_kind == 'sparse63':
            return np.unique(rng.integers(0, 2**63 - 1, total * 2,
                                          dtype=np.int64))[:total]
        return (rng.permutation(total * 2)[:total] - total).astype(np.int64)

    pool = make_ids(int(max(lens_prev.sum(), lens_cur.sum()) * 2))
    rng.shuffle(pool)
    blocks_prev, blocks_cur = [], []
    at = 0
    for lp, lc in zip(lens_prev, lens_cur):
        span = pool[at:at + max(lp, lc) + 50]
        at += len(span)
        blocks_prev.append(rng.permutation(span)[:lp])
        blocks_cur.append(rng.permutation(span)[:lc])
    trk = OrbitTracker()
    centres = rng.uniform(10, 90, (nh, 3)).astype(np.float32)
    bulk = np.zeros((nh, 3), dtype=np.float32)
    exists = np.arange(nh)
    results = []
    for blocks in (blocks_prev, blocks_cur):
        ids = np.concatenate(blocks)
        lens = np.array([len(b) for b in blocks])
        n = len(ids)
        snap = {'ids': ids, 'masses': 1.0,
                'coordinates': rng.uniform(0, 100, (n, 3)).astype(np.float32),
                'velocities': rng.normal(0, 1, (n, 3)).astype(np.float32),
                'region_offsets': np.concatenate(([0], np.cumsum(lens)[:-1])),
                'box_size': 100.0, 'redshift': 0.0}
        results.append((trk.step(snap, exists, centres, bulk, H=0.0,
                                 diagnostics=True), ids, lens))
    res, ids_cur, lens_c = results[1]
    _, ids_prev, lens_p = results[0]
    match = res.diag['match'].cpu().numpy()
    offp = np.concatenate(([0], np.cumsum(lens_p)))
    offc = np.concatenate(([0], np.cumsum(lens_c)))
    exp = np.full(len(ids_cur), -1, dtype=np.int64)
    for j in range(nh):
        prev_pos = {int(v): offp[j] + k
                    for k, v in enumerate(ids_prev[offp[j]:offp[j + 1]])}
        for c in range(offc[j], offc[j + 1]):
            exp[c] = prev_pos.get(int(ids_cur[c]), -1)
    assert np.array_equal(match, exp)
    assert (exp >= 0).sum() > 50000
