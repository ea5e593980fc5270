"""Host side of the second-generation partitioned join (``oa_pj2_step``,
``csrc/oa_pj2.cu``): ctypes mirrors of its argument structs and the per-snapshot
plan through ``oa_pj2_plan_host`` (partition counts that never shrink for a
halo, fixed partition capacities, record-slot layout, groups of regions and the
ticket ranges of the persistent kernel).  Nothing here touches particle data.

Reference: the plan only reorganises how ``track(j)`` (``track_orbits.py:
147-185``) is evaluated; results are independent of it
(``tests/test_gpu_zzz_pjoin.py`` runs several partition sizes against the
oracle).
"""
import ctypes as C
import os

import numpy as np

# shape of the kernel as compiled; `configure` reads it back from the library
THREADS, MIN_CTAS, TILE, CAP, TARGET, SIGMAS, LAG, SMEM = \
    480, 2, 1408, 704, 416, 8, 1, 0
# particles per group of regions (pipeline granularity: a group's records wait in
# L2 between its SCATTER and its JOIN).  OA_PJ2_GROUP overrides it (tuning).
GROUP_PARTICLES = int(os.environ.get('OA_PJ2_GROUP', 1 << 19))

PLAN_DTYPE = np.dtype([
    ('P_cur', np.uint32), ('cap_cur', np.uint32), ('base_cur', np.uint32),
    ('pb_cur', np.uint32), ('P_prev', np.uint32), ('cap_prev', np.uint32),
    ('base_prev', np.uint32), ('pb_prev', np.uint32), ('shift', np.uint32),
    ('group', np.uint32), ('join_first', np.uint32), ('reserved', np.uint32)])
assert PLAN_DTYPE.itemsize == 48

_vp, _i64, _i32, _u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32


class PJ2Args(C.Structure):
    """``oa_pj2_args`` -- keep in sync with include/orbit_b200.h."""
    _fields_ = [
        ('pos', _vp), ('vel', _vp), ('ids', _vp), ('n_cur', _i64),
        ('regions', _vp), ('plan', _vp), ('group_first', _vp),
        ('group_off', _vp), ('range_start', _vp),
        ('n_regions', _i32), ('n_groups', _i32), ('n_ranges', _i32),
        ('centre_f32', _i32), ('bulk_f32', _i32), ('periodic', _i32),
        ('mode', _i32), ('hubble_on', _i32),
        ('box', C.c_double * 3), ('hubble', C.c_double),
        ('one_plus_z', C.c_double),
        ('rec_prev', _vp), ('fill_prev', _vp), ('mark_prev', _vp),
        ('n_prev', _i64),
        ('rec_cur', _vp), ('fill_cur', _vp), ('mark_cur', _vp),
        ('workspace', _vp), ('workspace_bytes', C.c_size_t),
        ('n_part_entries', _i64), ('n_rec_slots', _i64),
        ('sm_reserve', _i32), ('total_tickets', _u32),
        ('overflow', _vp),
    ]


class PlanInfo(C.Structure):
    """``oa_pj2_plan_info``."""
    _fields_ = [('n_part_entries', _i64), ('n_rec_slots', _i64),
                ('total_tickets', _u32), ('n_groups', _i32),
                ('n_ranges', _i32), ('max_P', _u32)]


def configure(lib):
    """Take the kernel's compile-time shape from the library."""
    global THREADS, MIN_CTAS, TILE, CAP, TARGET, SIGMAS, LAG, SMEM
    out = (C.c_int32 * 8)()
    lib.oa_pj2_config(out)
    THREADS, MIN_CTAS, TILE, CAP, TARGET, SIGMAS, LAG, SMEM = list(out)
    return list(out)


class Plan:
    """Plan of one snapshot.  ``rows`` / ``group_first`` / ``group_off`` /
    ``range_start`` go to the device; ``P`` / ``cap`` / ``base`` / ``pb`` (per
    region) are kept on the host for the next snapshot."""
    __slots__ = ('rows', 'group_first', 'group_off', 'range_start', 'P', 'cap',
                 'base', 'pb', 'n_entries', 'n_slots', 'n_groups', 'n_ranges',
                 'total', 'max_P')


class Planner:
    """``oa_pj2_plan_host`` with scratch arrays reused between snapshots."""

    def __init__(self, lib):
        self._fn = lib.oa_pj2_plan_host
        self._cap = -1

    def _reserve(self, n_h):
        if n_h > self._cap:
            cap = int(n_h * 1.25) + 16
            self._rows = np.zeros(cap + 1, dtype=PLAN_DTYPE)
            self._group = np.zeros(cap + 1, dtype=np.uint32)
            self._goff = np.zeros(cap + 1, dtype=np.uint32)
            self._range = np.zeros(2 * (cap + LAG) + 1, dtype=np.uint32)
            self._cap = cap

    def __call__(self, offsets, prev_P, prev_cap, prev_base, prev_pb,
                 target=None, group_particles=None):
        target = TARGET if target is None else target
        group_particles = GROUP_PARTICLES if group_particles is None \
            else group_particles
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n_h = len(offsets) - 1
        prev = [np.ascontiguousarray(v, dtype=np.uint32)
                for v in (prev_P, prev_cap, prev_base, prev_pb)]
        self._reserve(n_h)
        out = [np.empty(n_h, dtype=np.uint32) for _ in range(4)]
        info = PlanInfo()
        vp = C.c_void_p
        rc = self._fn(offsets.ctypes.data_as(vp), n_h,
                      *[v.ctypes.data_as(vp) for v in prev],
                      int(group_particles), int(target),
                      self._rows.ctypes.data_as(vp),
                      *[v.ctypes.data_as(vp) for v in out],
                      self._group.ctypes.data_as(vp),
                      self._goff.ctypes.data_as(vp),
                      self._range.ctypes.data_as(vp), C.byref(info))
        if rc != 0:
            from ._lib import check
            check(rc)
        p = Plan()
        # (views into the scratch arrays: valid until the next call)
        p.rows = self._rows[:n_h + 1]
        p.group_first = self._group[:info.n_groups + 1]
        p.group_off = self._goff[:info.n_groups + 1]
        p.range_start = self._range[:info.n_ranges + 1]
        p.P, p.cap, p.base, p.pb = out
        p.n_entries = int(info.n_part_entries)
        p.n_slots = int(info.n_rec_slots)
        p.n_groups, p.n_ranges = int(info.n_groups), int(info.n_ranges)
        p.total = int(info.total_tickets)
        p.max_P = int(info.max_P)
        return p
