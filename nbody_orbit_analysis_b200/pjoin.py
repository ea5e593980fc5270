"""Host side of the partitioned hash join (``oa_pjoin_step``,
``csrc/oa_pjoin_core.cuh``): the per-snapshot plan -- partition bits of every
region, the layout of the partition-offset array, the work items of the four
stages and their ticket order -- built with vectorised numpy from the block
offsets.  Nothing here touches particle data.

Reference: the plan only reorganises how ``track(j)`` (``track_orbits.py:
147-185``) is evaluated; results are independent of it (tested against the
oracle for several partition sizes in ``tests/test_pjoin_emul.py``).
"""
import ctypes as C

import numpy as np

# Shape of the kernel as compiled (include/orbit_b200.h); `configure` replaces
# these defaults with what the loaded library reports (tuning builds differ).
THREADS = 512
TILE = 2048            # OA_PJOIN_TILE  (particles per SCATTER item)
CTILE = 8192           # OA_PJOIN_CTILE (particles per COUNT item)
REC_CAP = 2944         # OA_PJOIN_REC_CAP
TARGET = 2304          # OA_PJOIN_TARGET: particles per partition at most (mean)
MAX_BITS = 12          # OA_PJOIN_MAX_BITS
PACK_MAX = 32          # OA_PJOIN_PACK_MAX: small regions per work item
# particles per group of regions: the pipeline granularity (a group's COUNT,
# SCAN, SCATTER and JOIN items are one superstep apart; three groups of new
# records wait in L2).  OA_PJOIN_LAG overrides it (tuning).
import os as _os
LAG_PARTICLES = int(_os.environ.get('OA_PJOIN_LAG', 1 << 19))

JOIN, SCATTER, SCAN, COUNT = 0, 1, 2, 3
_LAG = (3, 2, 1, 0)    # superstep s holds stage `st` of group s - _LAG[st]

PLAN_DTYPE = np.dtype([
    ('pb_cur', np.uint32), ('pb_prev', np.uint32), ('bits_cur', np.int32),
    ('bits_prev', np.int32), ('tile_first', np.uint32),
    ('join_first', np.uint32), ('scan_first', np.uint32),
    ('count_first', np.uint32), ('pack_len', np.uint32),
    ('reserved', np.uint32, (3,))])
assert PLAN_DTYPE.itemsize == 48

_vp, _i64, _i32, _u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32


class PJoinArgs(C.Structure):
    """``oa_pjoin_args`` -- keep in sync with include/orbit_b200.h."""
    _fields_ = [
        ('pos', _vp), ('vel', _vp), ('ids', _vp), ('n_cur', _i64),
        ('regions', _vp), ('plan', _vp), ('group_first', _vp),
        ('range_start', _vp),
        ('n_regions', _i32), ('n_groups', _i32), ('n_ranges', _i32),
        ('centre_f32', _i32), ('bulk_f32', _i32), ('periodic', _i32),
        ('mode', _i32), ('hubble_on', _i32),
        ('box', C.c_double * 3), ('hubble', C.c_double),
        ('one_plus_z', C.c_double),
        ('rec_prev', _vp), ('part_off_prev', _vp), ('mark_prev', _vp),
        ('n_prev', _i64),
        ('rec_cur', _vp), ('part_off_cur', _vp), ('mark_cur', _vp),
        ('workspace', _vp), ('workspace_bytes', C.c_size_t),
        ('n_part_entries', _i64), ('sm_reserve', _i32),
        ('total_tickets', _u32),
    ]


def configure(lib):
    """Take the kernel's compile-time shape from the library."""
    global THREADS, TILE, CTILE, REC_CAP, TARGET, MAX_BITS
    out = (C.c_int32 * 8)()
    lib.oa_pjoin_config(out)
    THREADS, _, TILE, CTILE, REC_CAP, TARGET, MAX_BITS, _ = list(out)
    return list(out)


class PlanInfo(C.Structure):
    """``oa_pjoin_plan_info``."""
    _fields_ = [('n_part_entries', _i64), ('total_tickets', _u32),
                ('n_groups', _i32), ('n_ranges', _i32), ('max_bits', _i32)]


class Plan:
    """Plan of one snapshot.  ``rows`` / ``group_first`` / ``range_start`` go to
    the device; ``bits`` / ``pb`` are kept on the host for the next snapshot."""
    __slots__ = ('rows', 'group_first', 'range_start', 'bits', 'pb',
                 'n_entries', 'n_groups', 'n_ranges', 'total')


def need_bits(lens, target=None):
    """Smallest b with ``len <= target << b`` (at most MAX_BITS)."""
    target = TARGET if target is None else target
    thr = np.int64(target) << np.arange(MAX_BITS + 1, dtype=np.int64)
    return np.minimum(np.searchsorted(thr, lens, side='left'),
                      MAX_BITS).astype(np.int32)


def make_plan(offsets, prev_bits, prev_pb, prev_counts=None, target=None,
              lag_particles=None):
    """``offsets``: (n_regions + 1,) block starts + n.  ``prev_bits`` /
    ``prev_pb`` / ``prev_counts``: per region of THIS snapshot, the partition
    bits, the first partition-offset entry and the length of the same halo's
    previous block (bits -1: none).  numpy restatement of
    ``oa_pjoin_plan_host`` (the tests compare the two)."""
    target = TARGET if target is None else target
    lag_particles = LAG_PARTICLES if lag_particles is None else lag_particles
    offsets = np.asarray(offsets, dtype=np.int64)
    n_h = len(offsets) - 1
    lens = np.diff(offsets)
    prev_bits = np.asarray(prev_bits, dtype=np.int32)
    has_prev = prev_bits >= 0
    bits = np.maximum(need_bits(lens, target), np.where(has_prev, prev_bits, 0))
    big = bits > 0
    entries = (np.int64(1) << bits) + 1
    tiles = np.where(big, -(-lens // TILE), 0)
    ctiles = np.where(big, -(-lens // CTILE), 0)
    joins = np.where(big, np.where(has_prev,
                                   np.int64(1) << np.maximum(prev_bits, 0), 0), 1)
    # groups of consecutive regions (by the position of their first particle)
    gid = offsets[:-1] // lag_particles
    # packs of consecutive small regions of one group (greedy, like the C loop)
    prev_counts = np.zeros(n_h, dtype=np.int64) if prev_counts is None else \
        np.where(has_prev, np.asarray(prev_counts, dtype=np.int64), 0)
    pack_len = np.zeros(n_h, dtype=np.int64)
    leader, pc, pp = -1, 0, 0
    for j in range(n_h):
        if big[j]:
            leader = -1
            continue
        new_group = j == 0 or gid[j] != gid[j - 1]
        if (leader >= 0 and not new_group and pack_len[leader] < PACK_MAX and
                pc + lens[j] <= TILE and pp + prev_counts[j] <= REC_CAP):
            pack_len[leader] += 1
            pc += lens[j]
            pp += prev_counts[j]
            joins[j] = 0
        else:
            leader, pc, pp = j, int(lens[j]), int(prev_counts[j])
            pack_len[j] = 1

    p = Plan()
    rows = np.zeros(n_h + 1, dtype=PLAN_DTYPE)
    for name, per_region in (('pb_cur', entries), ('tile_first', tiles),
                             ('count_first', ctiles),
                             ('join_first', joins), ('scan_first', big)):
        pref = np.zeros(n_h + 1, dtype=np.int64)
        np.cumsum(per_region, out=pref[1:])
        assert pref[-1] < 2 ** 32
        rows[name] = pref
    rows['bits_cur'][:n_h] = bits
    rows['bits_prev'][:n_h] = prev_bits
    rows['pb_prev'][:n_h] = np.where(has_prev, prev_pb, 0)
    rows['pack_len'][:n_h] = pack_len
    p.rows = rows
    p.bits = bits
    p.pb = rows['pb_cur'][:n_h].astype(np.int64)
    p.n_entries = int(rows['pb_cur'][n_h])

    if n_h:
        cut = np.flatnonzero(gid[1:] != gid[:-1]) + 1
        group_first = np.concatenate(([0], cut, [n_h]))
    else:
        group_first = np.zeros(1, dtype=np.int64)
    G = len(group_first) - 1
    p.group_first = group_first.astype(np.uint32)
    p.n_groups = G

    # items per (superstep, stage): stage st of group s - lag[st]
    per = np.zeros((G + 3, 4), dtype=np.int64)
    for st, name in ((JOIN, 'join_first'), (SCATTER, 'tile_first'),
                     (SCAN, 'scan_first'), (COUNT, 'count_first')):
        pref = rows[name].astype(np.int64)
        per[_LAG[st]:_LAG[st] + G, st] = np.diff(pref[group_first])
    rs = np.zeros(4 * (G + 3) + 1, dtype=np.int64)
    np.cumsum(per.reshape(-1), out=rs[1:])
    assert rs[-1] < 2 ** 32
    p.range_start = rs.astype(np.uint32)
    p.n_ranges = 4 * (G + 3)
    p.total = int(rs[-1])
    return p


def decode(plan, ticket):
    """ticket -> (stage, region, index): what the expand pre-kernel writes
    into the item list (Python restatement for the tests)."""
    rs = plan.range_start.astype(np.int64)
    r = int(np.searchsorted(rs, ticket, side='right')) - 1
    r = min(r, plan.n_ranges - 1)
    stage = r & 3
    g = (r >> 2) - _LAG[stage]
    name = {JOIN: 'join_first', SCAN: 'scan_first',
            COUNT: 'count_first'}.get(stage, 'tile_first')
    pref = plan.rows[name].astype(np.int64)
    gf = plan.group_first.astype(np.int64)
    target = pref[gf[g]] + (ticket - rs[r])
    j = int(np.searchsorted(pref[gf[g]:gf[g + 1]], target, side='right')) - 1
    j += int(gf[g])
    return stage, j, int(target - pref[j])


class Planner:
    """``make_plan`` through the C ABI (``oa_pjoin_plan_host``: one loop over the
    regions instead of ~40 numpy calls; 0.2 ms -> a few microseconds per
    snapshot at 1000 regions).  Scratch arrays are reused between snapshots."""

    def __init__(self, lib):
        self._fn = lib.oa_pjoin_plan_host
        self._cap = -1

    def _reserve(self, n_h):
        if n_h > self._cap:
            cap = int(n_h * 1.25) + 16
            self._rows = np.zeros(cap + 1, dtype=PLAN_DTYPE)
            self._group = np.zeros(cap + 1, dtype=np.uint32)
            self._range = np.zeros(4 * (cap + 3) + 1, dtype=np.uint32)
            self._cap = cap

    def __call__(self, offsets, prev_bits, prev_pb, prev_counts=None,
                 target=None, lag_particles=None):
        target = TARGET if target is None else target
        lag_particles = LAG_PARTICLES if lag_particles is None else lag_particles
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        prev_bits = np.ascontiguousarray(prev_bits, dtype=np.int32)
        prev_pb = np.ascontiguousarray(prev_pb, dtype=np.int64)
        n_h = len(offsets) - 1
        prev_counts = np.zeros(n_h, dtype=np.int64) if prev_counts is None \
            else np.ascontiguousarray(prev_counts, dtype=np.int64)
        self._reserve(n_h)
        bits = np.empty(n_h, dtype=np.int32)
        pb = np.empty(n_h, dtype=np.int64)
        info = PlanInfo()
        vp = C.c_void_p
        rc = self._fn(offsets.ctypes.data_as(vp), n_h,
                      prev_bits.ctypes.data_as(vp), prev_pb.ctypes.data_as(vp),
                      prev_counts.ctypes.data_as(vp),
                      int(target), int(lag_particles),
                      self._rows.ctypes.data_as(vp), bits.ctypes.data_as(vp),
                      pb.ctypes.data_as(vp), self._group.ctypes.data_as(vp),
                      self._range.ctypes.data_as(vp), C.byref(info))
        if rc != 0:
            raise RuntimeError('oa_pjoin_plan_host failed (%d)' % rc)
        p = Plan()
        # (views into the scratch arrays: valid until the next call)
        p.rows = self._rows[:n_h + 1]
        p.group_first = self._group[:info.n_groups + 1]
        p.range_start = self._range[:info.n_ranges + 1]
        p.bits, p.pb = bits, pb
        p.n_entries = int(info.n_part_entries)
        p.n_groups, p.n_ranges = int(info.n_groups), int(info.n_ranges)
        p.total = int(info.total_tickets)
        return p
