"""Kernels of the multi-GPU event exchange (csrc/oa_segment.cu: oa_pack_events,
oa_merge_gathered, oa_split_quantiles, oa_pack_split, oa_merge_blocks) on ONE
device: ``world`` virtual ranks, the collectives replaced by slicing.

The GPU tests compare the kernels with the brute-force expectation (all events
sorted by their position in the unsharded previous snapshot, reference order
``track_orbits.py:315-316``) and with the numpy restatements of
``tests/exchange_emul.py``; the CPU test runs the same harness on the numpy
restatements, which pins the restatements (and the harness) themselves.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import exchange_emul as emul


class _Backend:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type == 'cuda':
            from nbody_orbit_analysis_b200 import _lib
            self.lib, self.ptr, self.check = _lib.lib, _lib.ptr, _lib.check
            self.st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        else:
            self.lib, self.ptr, self.check = emul.EmulLib(), (lambda t: t), \
                (lambda rc: None)
            self.st = None

    def dev(self, arr):
        # (an empty list is a zero-element tensor, data_ptr() == 0: the C ABI
        # accepts NULL arrays for empty lists, like Tracker._to_device produces)
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def empty(self, n, dtype):
        # poisoned, not zeroed: a kernel must write everything that is read
        t = torch.empty(max(int(n), 1), dtype=dtype, device=self.device)
        t.view(torch.uint8).fill_(0xA5)
        return t


def make_world(world, n_prev, n_events, n_halos, seed):
    """Global event list + the share of every virtual rank (id mod world)."""
    rng = np.random.default_rng(seed)
    starts = np.concatenate(([0], np.sort(rng.choice(
        np.arange(1, n_prev), n_halos - 1, replace=False))))
    pid = rng.permutation(n_prev).astype(np.int64) + 7 * 10 ** 11
    ev_pos = np.sort(rng.choice(n_prev, n_events, replace=False))
    ang = rng.standard_normal(n_prev).astype(np.float16)
    glob = {'ids': pid[ev_pos], 'ang': ang[ev_pos],
            'offsets': np.append(np.searchsorted(ev_pos, starts), n_events)}
    ranks = []
    for r in range(world):
        mine = np.flatnonzero(pid % world == r).astype(np.int64)
        sel = np.flatnonzero(np.isin(mine, ev_pos)).astype(np.int64)
        seg = np.searchsorted(mine, starts)
        offs = np.searchsorted(sel, seg).astype(np.int64)
        ranks.append({'gpos': mine, 'sel': sel, 'ids': pid[mine][sel],
                      'ang': ang[mine][sel].view(np.int16),
                      'small': np.append(offs, len(sel))})
    return glob, ranks


def run_split(be, ranks, world, n_seg, cap):
    """Quantiles -> (all-gather) -> pack -> (all-to-all) -> merge."""
    lib, ptr, check, st = be.lib, be.ptr, be.check, be.st
    d = [{k: be.dev(v) for k, v in r.items()} for r in ranks]
    props = []
    for r in range(world):
        p = be.empty(max(world - 1, 1), torch.int64)
        check(lib.oa_split_quantiles(ptr(d[r]['gpos']), ptr(d[r]['sel']),
                                     ptr(d[r]['small']), n_seg, world, ptr(p),
                                     st))
        props.append(p[:max(world - 1, 1)])
    prop_all = torch.cat(props)
    blk = lib.oa_exchange_bytes(0, cap)
    sends, counts, bnds = [], [], []
    for r in range(world):
        send = be.empty(world * blk, torch.uint8)
        cnt = be.empty(n_seg, torch.int64)
        bnd = be.empty(world + 1, torch.int64)
        check(lib.oa_pack_split(
            ptr(d[r]['gpos']), ptr(d[r]['sel']), ptr(d[r]['ids']),
            ptr(d[r]['ang']), ptr(d[r]['small']), n_seg, ptr(prop_all), world,
            cap, ptr(bnd), ptr(send), ptr(cnt), st))
        sends.append(send)
        counts.append(cnt[:n_seg].cpu().numpy())
        bnds.append(bnd.cpu().numpy())
    out = []
    for q in range(world):
        recv = torch.cat([s[q * blk:(q + 1) * blk] for s in sends]).contiguous()
        ids = be.empty(world * cap, torch.int64)
        ang = be.empty(world * cap, torch.int16)
        info = be.empty(2, torch.int64)
        check(lib.oa_merge_blocks(ptr(recv), world, cap, ptr(ids), ptr(ang),
                                  ptr(info), st))
        info = info.cpu().numpy()
        n = int(info[0])
        out.append((ids[:n].cpu().numpy(), ang[:n].cpu().numpy(), info))
    return out, sum(counts), bnds, prop_all.cpu().numpy()


def run_gather(be, ranks, world, n_seg, cap):
    """Pack -> (all-gather) -> merge on every rank (rank 0's copy checked)."""
    lib, ptr, check, st = be.lib, be.ptr, be.check, be.st
    nbytes = lib.oa_exchange_bytes(n_seg, cap)
    sends = []
    for r in ranks:
        d = {k: be.dev(v) for k, v in r.items()}
        send = be.empty(nbytes, torch.uint8)
        check(lib.oa_pack_events(ptr(d['gpos']), ptr(d['sel']), ptr(d['ids']),
                                 ptr(d['ang']), ptr(d['small']), n_seg, cap,
                                 ptr(send), st))
        sends.append(send[:nbytes])
    recv = torch.cat(sends).contiguous()
    ids = be.empty(world * cap, torch.int64)
    ang = be.empty(world * cap, torch.int16)
    info = be.empty(n_seg + 3 + world, torch.int64)
    check(lib.oa_merge_gathered(ptr(recv), world, n_seg, cap, ptr(ids),
                                ptr(ang), ptr(info), st))
    info = info.cpu().numpy()
    n = int(info[0])
    return ids[:n].cpu().numpy(), ang[:n].cpu().numpy(), info


CASES = [
    # world, n_prev, n_events, n_halos
    (2, 5000, 700, 5),
    (4, 40000, 9000, 40),
    (8, 200000, 60000, 300),
    (3, 3000, 3000, 7),           # every particle has an event
    (8, 5000, 3, 4),              # fewer events than ranks
    (4, 5000, 0, 6),              # no events at all
    (1, 4000, 900, 3),
]


def check_split(be, world, n_prev, n_events, n_halos, seed=5):
    glob, ranks = make_world(world, n_prev, n_events, n_halos, seed)
    largest_local = max(len(r['sel']) for r in ranks)
    cap = (largest_local // max(world, 1) + 64) * 2
    out, counts, bnds, _ = run_split(be, ranks, world, n_halos, cap)
    ids = np.concatenate([o[0] for o in out])
    ang = np.concatenate([o[1] for o in out])
    assert np.array_equal(ids, glob['ids'])
    assert np.array_equal(ang, glob['ang'].view(np.int16))
    assert np.array_equal(np.concatenate(([0], np.cumsum(counts))),
                          glob['offsets'])
    for r, b in enumerate(bnds):
        assert b[0] == 0 and b[world] == len(ranks[r]['sel'])
        assert np.all(np.diff(b[:world + 1]) >= 0)
    for o in out:
        assert o[2][1] <= cap                      # nothing was truncated
    if n_events >= 64 * world:
        # quantile splitters balance the slices (uniform sample per rank)
        sizes = np.array([len(o[0]) for o in out])
        assert sizes.max() <= 1.5 * n_events / world + 64
    return glob, ranks, out


def check_split_overflow(be, world, n_prev, n_events, n_halos, seed=6):
    """Too small a capacity: the true block sizes are reported (so the host
    can size the repeat), nothing is written out of bounds, what fits is kept
    in key order."""
    glob, ranks = make_world(world, n_prev, n_events, n_halos, seed)
    cap = 16
    out, counts, bnds, _ = run_split(be, ranks, world, n_halos, cap)
    true_blocks = np.array([np.diff(b[:world + 1]) for b in bnds])  # [src, dst]
    for q, o in enumerate(out):
        assert o[2][1] == true_blocks[:, q].max()
        assert o[2][0] == np.minimum(true_blocks[:, q], cap).sum()
    assert max(o[2][1] for o in out) > cap
    # the same exchange with room for the largest block is complete
    cap2 = int(max(o[2][1] for o in out))
    out2, _, _, _ = run_split(be, ranks, world, n_halos, cap2)
    assert np.array_equal(np.concatenate([o[0] for o in out2]), glob['ids'])


def check_gather(be, world, n_prev, n_events, n_halos, seed=7):
    glob, ranks = make_world(world, n_prev, n_events, n_halos, seed)
    sizes = np.array([len(r['sel']) for r in ranks])
    cap = int(sizes.max()) + 5
    ids, ang, info = run_gather(be, ranks, world, n_halos, cap)
    assert info[0] == n_events
    assert np.array_equal(info[1:2 + n_halos], glob['offsets'])
    assert np.array_equal(info[2 + n_halos:2 + n_halos + world], sizes)
    assert info[2 + n_halos + world] == 0
    assert np.array_equal(ids, glob['ids'])
    assert np.array_equal(ang, glob['ang'].view(np.int16))
    if sizes.max() > 8:                            # overflow: flag + true sizes
        cap = int(sizes.max()) - 3
        ids, ang, info = run_gather(be, ranks, world, n_halos, cap)
        assert info[2 + n_halos + world] == 1
        assert np.array_equal(info[2 + n_halos:2 + n_halos + world], sizes)
        assert info[0] == np.minimum(sizes, cap).sum()


# -- CPU: the numpy restatements against the brute-force expectation ---------
@pytest.mark.parametrize('case', CASES)
def test_exchange_emulation_split(case):
    check_split(_Backend('cpu'), *case)


@pytest.mark.parametrize('case', CASES)
def test_exchange_emulation_gather(case):
    check_gather(_Backend('cpu'), *case)


def test_exchange_emulation_overflow():
    check_split_overflow(_Backend('cpu'), 4, 40000, 9000, 40)


# -- GPU: the kernels -----------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_exchange_kernels_split(case):
    world, n_prev, n_events, n_halos = case
    glob, ranks, out = check_split(_Backend('cuda'), *case)
    # identical to the numpy restatement, slice by slice
    cap = (max(len(r['sel']) for r in ranks) // max(world, 1) + 64) * 2
    ref, _, _, _ = run_split(_Backend('cpu'), ranks, world, n_halos, cap)
    for a, b in zip(out, ref):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[2], b[2])


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_exchange_kernels_gather(case):
    check_gather(_Backend('cuda'), *case)


@pytest.mark.gpu
def test_exchange_kernels_overflow():
    check_split_overflow(_Backend('cuda'), 4, 40000, 9000, 40)
    check_split_overflow(_Backend('cuda'), 8, 200000, 60000, 300)


# -- staging of several snapshots for one exchange (oa_stage_events) -----------
def check_stage(be):
    """Three snapshots staged back to back; the staged arrays exchanged as ONE
    snapshot give every snapshot's list in order (keys tagged by slot)."""
    lib, ptr, check, st = be.lib, be.ptr, be.check, be.st
    world, n_prev = 3, 20000
    specs = [(1500, 6), (0, 4), (4200, 9)]            # (events, halos) per snapshot
    globs, staged = [], []
    for r in range(world):
        staged.append({'keys': be.empty(9000, torch.int64),
                       'ids': be.empty(9000, torch.int64),
                       'ang': be.empty(9000, torch.int16),
                       'small': be.empty(32, torch.int64), 'n': 0, 'seg': 0})
    for slot, (m, n_halos) in enumerate(specs):
        glob, ranks = make_world(world, n_prev, m, n_halos, seed=40 + slot)
        globs.append(glob)
        for r, rk in enumerate(ranks):
            d = {k: be.dev(v) for k, v in rk.items()}
            s = staged[r]
            n_local = len(rk['sel'])
            small_out = s['small'][s['seg']:]
            check(lib.oa_stage_events(
                ptr(d['gpos']), ptr(d['sel']), ptr(d['ids']), ptr(d['ang']),
                ptr(d['small']), n_halos, n_local, slot << 58, s['n'],
                ptr(s['keys']), ptr(s['ids']), ptr(s['ang']), ptr(small_out), st))
            s['n'] += n_local
            s['seg'] += n_halos
    # the batch as one snapshot per rank: identity selection over the staged keys
    n_seg = sum(h for _, h in specs)
    batch_ranks = []
    for s in staged:
        n = s['n']
        small = s['small'][:n_seg + 1].cpu().numpy()
        assert small[-1] == n and np.all(np.diff(small) >= 0)
        batch_ranks.append({'gpos': s['keys'][:n].cpu().numpy(),
                            'sel': np.arange(n, dtype=np.int64),
                            'ids': s['ids'][:n].cpu().numpy(),
                            'ang': s['ang'][:n].cpu().numpy(), 'small': small})
    cap = max(r_['small'][-1] for r_ in batch_ranks) // world * 2 + 64
    out, counts, _, _ = run_split(be, batch_ranks, world, n_seg, int(cap))
    ids = np.concatenate([o[0] for o in out])
    ang = np.concatenate([o[1] for o in out])
    assert np.array_equal(ids, np.concatenate([g['ids'] for g in globs]))
    assert np.array_equal(ang, np.concatenate(
        [g['ang'].view(np.int16) for g in globs]))
    exp_counts = np.concatenate([np.diff(g['offsets']) for g in globs])
    assert np.array_equal(counts, exp_counts)


def test_stage_events_emulation():
    check_stage(_Backend('cpu'))


@pytest.mark.gpu
def test_stage_events_kernel():
    check_stage(_Backend('cuda'))
