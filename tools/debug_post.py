import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from fixture_io import load_fixture, expected_tree
from nbody_orbit_analysis_b200 import _lib
from nbody_orbit_analysis_b200._device import DeviceContext
from nbody_orbit_analysis_b200._lib import lib, check, ptr
ctx = DeviceContext(); st = ctx.stream()
rng = np.random.default_rng(0)
# 1. expand_segments
offs = np.array([0, 5, 5, 9, 20], dtype=np.int64); table = np.array([7, 3, 1, 0], dtype=np.int32)
out = ctx.empty(20, torch.int32)
check(lib.oa_expand_segments(ptr(ctx.upload(offs)), 4, ptr(ctx.upload(table)), 20, ptr(out), st))
print('expand', out.cpu().numpy(), np.repeat(table, np.diff(offs)))
# 2. sort by (seg, id) + run heads + segment offsets
P = 300
ids = rng.integers(100, 140, P).astype(np.int64); seg = rng.integers(0, 5, P).astype(np.int64)
d_ids, d_seg = ctx.upload(ids), ctx.upload(seg)
_, order, _ = ctx.argsort_values(d_ids, P)
seg_sorted = ctx.gather_i64(d_seg, order, P)
seg_sorted, order = ctx.sort_pairs(seg_sorted, order, P, 3)
ids_sorted = ctx.gather_i64(d_ids, order, P)
o = np.lexsort((ids, seg))
print('sort ok', np.array_equal(ids_sorted[:P].cpu().numpy(), ids[o]), np.array_equal(seg_sorted[:P].cpu().numpy(), seg[o]))
head = ctx.empty(P + 8, torch.int16)
check(lib.oa_run_heads(ptr(seg_sorted), ptr(ids_sorted), P, ptr(head), st))
starts, n_runs = ctx.select(head, P, _lib.OA_SEL_EQ, 1)
si, ss = ids[o], seg[o]
h = np.ones(P, bool); h[1:] = (si[1:] != si[:-1]) | (ss[1:] != ss[:-1])
print('heads ok', n_runs == h.sum(), np.array_equal(starts[:n_runs].cpu().numpy(), np.flatnonzero(h)))
u_seg = ctx.gather_i64(seg_sorted, starts, n_runs)
d_keys = ctx.upload(np.arange(6, dtype=np.int64)); d_poff = ctx.empty(6, torch.int64)
check(lib.oa_segment_offsets(ptr(u_seg), n_runs, None, ptr(d_keys), 6, ptr(d_poff), st))
print('poff', d_poff.cpu().numpy(), np.searchsorted(ss[h], np.arange(6)))
counts = ctx.empty(n_runs, torch.int64)
check(lib.oa_run_lengths(ptr(starts), n_runs, P, ptr(counts), st))
print('counts ok', np.array_equal(counts[:n_runs].cpu().numpy(), np.diff(np.append(np.flatnonzero(h), P))))
# 3. angle cut + select
ang = rng.random(1000).astype(np.float16); cut = float(np.float16(np.pi / 4))
d_ang = ctx.upload(ang.view(np.int16)); marks = ctx.empty(1008, torch.int16)
check(lib.oa_angle_cut(ptr(d_ang), 1000, cut, ptr(marks), st))
sel, n_sel = ctx.select(marks, 1000, _lib.OA_SEL_EQ, 1)
print('cut ok', np.array_equal(sel[:n_sel].cpu().numpy(), np.flatnonzero(ang > np.pi / 4)))
