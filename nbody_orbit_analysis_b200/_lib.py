"""ctypes binding of ``liborbit_b200.so`` (the C ABI in ``include/orbit_b200.h``).

There is no CPU fallback: if the library cannot be loaded this module raises,
and every compute entry point returns an error without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

from . import _build

OA_F32, OA_F64 = 0, 1
OA_MODE = {'pericentric': 0, 'apocentric': 1}
OA_SEL_NE, OA_SEL_EQ = 0, 1
OA_NO_EVENT = 0x8000
ABI_VERSION = 18
BUCKET_LOAD = 3          # OA_BUCKET_LOAD


class OrbitB200Error(RuntimeError):
    pass


def _load():
    path = os.environ.get('OA_LIB_PATH') or _build.LIB    # (tuning builds)
    if path == _build.LIB and (not os.path.exists(path) or _build.needs_build()):
        try:
            _build.build()
        except Exception as exc:    # noqa: BLE001
            if not os.path.exists(path):
                raise ImportError(
                    "liborbit_b200.so is missing and could not be built (%s). "
                    "This package has no CPU fallback: build it with "
                    "`python -m nbody_orbit_analysis_b200._build`." % (exc,))
    lib = C.CDLL(path)
    lib.oa_abi_version.restype = C.c_int
    if lib.oa_abi_version() != ABI_VERSION:
        raise ImportError(
            "liborbit_b200.so has ABI %d, the Python host expects %d; rebuild "
            "with `python -m nbody_orbit_analysis_b200._build --force`"
            % (lib.oa_abi_version(), ABI_VERSION))
    return lib


lib = _load()
LIB_PATH = _build.LIB

# numpy view of `oa_region` (128 bytes)
REGION_DTYPE = np.dtype([
    ('centre', np.float64, (3,)), ('bulk', np.float64, (3,)),
    ('prev_begin', np.int64), ('prev_count', np.int64),
    ('prev_bucket', np.int64), ('cur_bucket', np.int64),
    ('cur_begin', np.int64), ('cur_count', np.int64),
    ('centre_f', np.float32, (3,)), ('bulk_f', np.float32, (3,)),
    ('reserved', np.int64)])
assert REGION_DTYPE.itemsize == 128

_vp, _i64, _i32, _sz, _u16 = (C.c_void_p, C.c_int64, C.c_int32, C.c_size_t,
                              C.c_uint16)


class TrackArgs(C.Structure):
    """``oa_track_args`` -- keep in sync with include/orbit_b200.h."""
    _fields_ = [
        ('pos', _vp), ('vel', _vp), ('ids', _vp), ('n_cur', _i64),
        ('cur_off', _vp), ('regions', _vp),
        ('n_regions', _i32), ('data_dtype', _i32), ('frame_dtype', _i32),
        ('centre_f32', _i32), ('bulk_f32', _i32), ('periodic', _i32),
        ('onthefly', _i32), ('mode', _i32),
        ('box', C.c_double * 3), ('hubble', C.c_double),
        ('one_plus_z', C.c_double),
        ('rec_prev', _vp), ('tab_prev', _vp), ('tab_prev_buckets', _i64),
        ('n_prev', _i64),
        ('prev_index_bits', _i32), ('cur_index_bits', _i32),
        ('mark_prev', _vp),
        ('rec_cur', _vp), ('tab_cur', _vp), ('tab_cur_buckets', _i64),
        ('mark_cur', _vp), ('workspace', _vp), ('workspace_bytes', _sz),
        ('sm_reserve', _i32), ('reserved0', _i32),
        ('out_rhat', _vp), ('out_vr', _vp), ('out_r', _vp),
        ('out_angle', _vp), ('out_match', _vp), ('dangle_prev', _vp),
    ]


class SynthParams(C.Structure):
    """``oa_synth_params`` -- keep in sync with include/orbit_b200.h."""
    _fields_ = [
        ('seed', C.c_uint64), ('n_universe', _i64), ('id_stride', _i64),
        ('id_offset', _i64), ('halo_start', _vp), ('halo_radius', _vp),
        ('halo_c0', _vp), ('halo_vh', _vp), ('box', C.c_double),
        ('t', C.c_double), ('n_halos', _i32), ('periodic', _i32),
    ]


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


# every symbol declared in include/orbit_b200.h
_sig('oa_last_error', C.c_char_p)
_sig('oa_device_info', C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
     C.POINTER(C.c_int), C.POINTER(_i64), C.POINTER(_i64))
_sig('oa_record_bytes', _sz, C.c_int)
_sig('oa_table_slots', _i64, _i64, _i64)
_sig('oa_table_buckets', _i64, _i64, _i64)
_sig('oa_track_workspace_bytes', _sz, _i64)
_sig('oa_table_bucket_begin', _i64, _i64, _i64)
_sig('oa_index_bits', C.c_int, _i64)
_sig('oa_bulk_workspace_bytes', _sz, _i64, C.c_int)
_sig('oa_bulk_velocity', C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int,
     _i64, C.c_int, _vp, _vp, _vp, _sz, _vp)
_sig('oa_track_fused', C.c_int, C.POINTER(TrackArgs), _vp)
_sig('oa_track_args_size', _sz)
_sig('oa_table_clear', C.c_int, _vp, _i64, _i64, _vp)
if lib.oa_track_args_size() != C.sizeof(TrackArgs):
    raise ImportError("oa_track_args layout mismatch: C %d bytes, ctypes %d"
                      % (lib.oa_track_args_size(), C.sizeof(TrackArgs)))
_sig('oa_select_workspace_bytes', _sz, _i64)
_sig('oa_select_count', C.c_int, _vp, _i64, C.c_int, _u16, _vp, _sz, _vp, _vp)
_sig('oa_select_gather', C.c_int, _vp, _i64, C.c_int, _u16, _vp, _vp, _vp)
_sig('oa_select_gather_events', C.c_int, _vp, _i64, _vp, _vp, C.c_int, _vp, _vp,
     _vp, _vp)
_sig('oa_segment_offsets', C.c_int, _vp, _i64, _vp, _vp, C.c_int, _vp, _vp)
_sig('oa_gather_record_ids', C.c_int, _vp, C.c_int, _vp, _i64, _vp, _vp, _vp)
_sig('oa_gather_u16', C.c_int, _vp, _vp, _i64, _vp, _vp, _vp)
_sig('oa_gather_i64', C.c_int, _vp, _vp, _i64, _vp, _vp, _vp)
_sig('oa_gather_f', C.c_int, _vp, C.c_int, _vp, _i64, _vp, _vp, _vp)
_sig('oa_mark_unmatched', C.c_int, _vp, _i64, _vp, _vp)
_sig('oa_fill_u16', C.c_int, _vp, _i64, _u16, _vp)
_sig('oa_set_record_angles', C.c_int, _vp, C.c_int, _vp, _i64, _vp)
_sig('oa_sort_workspace_bytes', _sz, _i64)
_sig('oa_sort_pairs_u64', C.c_int, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int,
     _vp, _sz, _vp)
_sig('oa_minmax_i64', C.c_int, _vp, _i64, _vp, _vp)
_sig('oa_segment_sort_keys', C.c_int, _vp, _i64, _vp, C.c_int, _vp, _vp, _vp,
     _vp, _vp, _vp)
_sig('oa_merge_event_lists', C.c_int, _vp, _vp, _vp, _i64, _vp, C.c_int, _vp,
     _vp, _vp)
_sig('oa_exchange_bytes', _sz, C.c_int, _i64)
_sig('oa_pack_events', C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _i64, _vp, _vp)
_sig('oa_merge_gathered', C.c_int, _vp, C.c_int, C.c_int, _i64, _vp, _vp, _vp, _vp)
_sig('oa_split_quantiles', C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp)
_sig('oa_pack_split', C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int,
     _i64, _vp, _vp, _vp, _vp)
_sig('oa_merge_blocks', C.c_int, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp)
_sig('oa_run_heads', C.c_int, _vp, _vp, _i64, _vp, _vp)
_sig('oa_run_lengths', C.c_int, _vp, _i64, _i64, _vp, _vp)
_sig('oa_central_radii', C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int,
     C.c_int, C.POINTER(C.c_double), _i64, _vp, _vp)
_sig('oa_segment_heads', C.c_int, _vp, _vp, _vp, _vp, C.c_int, _i64, _vp, _vp)
_sig('oa_scatter_flags', C.c_int, _vp, _vp, _i64, _vp, _vp)
_sig('oa_lookup_sorted', C.c_int, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp,
     _i64, _vp, _vp)
_sig('oa_vote_keys', C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, _i64, _vp, _vp)
_sig('oa_vote_reduce', C.c_int, _vp, _i64, C.c_int, _vp, _vp, _vp)
_sig('oa_angle_cut', C.c_int, _vp, _i64, C.c_double, _vp, _vp)
_sig('oa_expand_segments', C.c_int, _vp, C.c_int, _vp, _i64, _vp, _vp)
_sig('oa_synth_keys', C.c_int, C.POINTER(SynthParams), _vp, _vp, _vp, _vp)
_sig('oa_synth_fill', C.c_int, C.POINTER(SynthParams), _vp, _i64, C.c_int,
     _vp, _vp, _vp, _vp)
_sig('oa_synth_params_size', _sz)
if lib.oa_synth_params_size() != C.sizeof(SynthParams):
    raise ImportError("oa_synth_params layout mismatch")
_sig('oa_stage_events', C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _i64, _i64, _i64,
     _vp, _vp, _vp, _vp, _vp)
_sig('oa_region_rows_host', C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, _vp,
     _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int))
_sig('oa_select_gather_events_ids', C.c_int, _vp, _i64, _vp, _vp, _vp, _vp, _vp,
     _vp)
_sig('oa_pjoin_workspace_bytes', _sz, C.c_int, _i64, C.c_uint32)
_sig('oa_pjoin_args_size', _sz)
_sig('oa_pjoin_plan_host', C.c_int, _vp, C.c_int, _vp, _vp, _vp, _i64, _i64, _vp,
     _vp, _vp, _vp, _vp, _vp)
_sig('oa_pjoin_step', C.c_int, _vp, _vp)
from . import pjoin as _pjoin        # noqa: E402  (struct mirror of oa_pjoin_args)
if lib.oa_pjoin_args_size() != C.sizeof(_pjoin.PJoinArgs):
    raise ImportError("oa_pjoin_args layout mismatch")
lib.oa_pjoin_config.restype = None
lib.oa_pjoin_config.argtypes = [C.POINTER(C.c_int32)]
_pjoin.configure(lib)
_sig('oa_pjoin_stats', C.c_int, _vp, C.c_int)
_sig('oa_copy_async', C.c_int, _vp, _vp, _sz, _vp)
_sig('oa_copy_small', C.c_int, _vp, _vp, _sz, _vp)
_sig('oa_host_copy', C.c_int, _vp, _vp, _sz, C.c_int)
_sig('oa_pairwise_sum_host', C.c_int, _vp, C.c_int, _i64, C.POINTER(C.c_double))
_sig('oa_region_pairs', C.c_int, _vp, C.c_int, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp,
     _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp)
_sig('oa_gather_by_key', C.c_int, _vp, C.c_int, C.c_int, _vp, _i64, _vp, _vp)
_sig('oa_merge_find', C.c_int, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp)
_sig('oa_merge_place', C.c_int, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp,
     _vp, _vp)
_sig('oa_pj2_workspace_bytes', _sz, C.c_int, C.c_uint32)
_sig('oa_pj2_args_size', _sz)
_sig('oa_pj2_plan_host', C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _i64, C.c_int32,
     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp)
_sig('oa_pj2_step', C.c_int, _vp, _vp)
_sig('oa_pj2_stats', C.c_int, _vp, C.c_int)
from . import pj2 as _pj2            # noqa: E402  (struct mirror of oa_pj2_args)
if lib.oa_pj2_args_size() != C.sizeof(_pj2.PJ2Args):
    raise ImportError("oa_pj2_args layout mismatch")
lib.oa_pj2_config.restype = None
lib.oa_pj2_config.argtypes = [C.POINTER(C.c_int32)]
_pj2.configure(lib)

EXPORTS = [
    'oa_abi_version', 'oa_last_error', 'oa_device_info', 'oa_record_bytes',
    'oa_table_slots', 'oa_table_buckets', 'oa_track_workspace_bytes',
    'oa_table_bucket_begin', 'oa_index_bits', 'oa_bulk_workspace_bytes',
    'oa_bulk_velocity', 'oa_track_fused', 'oa_track_args_size', 'oa_table_clear',
    'oa_select_workspace_bytes',
    'oa_select_count', 'oa_select_gather', 'oa_segment_offsets',
    'oa_gather_record_ids', 'oa_gather_u16', 'oa_gather_i64', 'oa_gather_f',
    'oa_mark_unmatched', 'oa_fill_u16', 'oa_set_record_angles',
    'oa_sort_workspace_bytes',
    'oa_sort_pairs_u64', 'oa_minmax_i64', 'oa_synth_keys', 'oa_synth_fill',
    'oa_synth_params_size', 'oa_segment_sort_keys', 'oa_run_heads',
    'oa_run_lengths', 'oa_merge_event_lists', 'oa_central_radii',
    'oa_segment_heads', 'oa_scatter_flags', 'oa_lookup_sorted', 'oa_vote_keys',
    'oa_vote_reduce', 'oa_angle_cut', 'oa_expand_segments', 'oa_exchange_bytes',
    'oa_pack_events', 'oa_merge_gathered', 'oa_select_gather_events',
    'oa_split_quantiles', 'oa_pack_split', 'oa_merge_blocks',
    'oa_select_gather_events_ids', 'oa_pjoin_workspace_bytes',
    'oa_pjoin_args_size', 'oa_pjoin_step', 'oa_pjoin_plan_host',
    'oa_region_rows_host', 'oa_pjoin_config', 'oa_stage_events',
    'oa_pjoin_stats', 'oa_pj2_workspace_bytes', 'oa_pj2_args_size',
    'oa_pj2_plan_host', 'oa_pj2_step', 'oa_pj2_stats', 'oa_pj2_config',
    'oa_copy_async', 'oa_merge_find', 'oa_merge_place', 'oa_region_pairs',
    'oa_gather_by_key', 'oa_pairwise_sum_host', 'oa_copy_small', 'oa_host_copy',
]


def check(rc):
    """Raise on a non-zero return code of the C ABI."""
    if rc != 0:
        msg = lib.oa_last_error()
        raise OrbitB200Error(
            "liborbit_b200 error %d: %s" % (rc, msg.decode() if msg else '?'))


def ptr(t):
    """Device (or pinned host) address of a torch tensor, NULL for None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def dtype_code(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float32:
        return OA_F32
    if np_dtype == np.float64:
        return OA_F64
    raise TypeError("unsupported floating dtype %s" % np_dtype)
