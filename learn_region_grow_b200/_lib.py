"""ctypes binding of liblrg_b200.so (include/lrg_b200.h).  There is no CPU fallback: if the library is missing or a
call fails this raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'liblrg_b200.so')


class LrgError(RuntimeError):
    pass


class GrowParams(C.Structure):
    _fields_ = [('resolution', C.c_float), ('cluster_threshold', C.c_int), ('seed', C.c_uint64),
                ('max_slots', C.c_int), ('max_steps_per_region', C.c_int), ('room_id_base', C.c_int),
                ('trace_capacity', C.c_int), ('flags', C.c_int), ('num_restarts', C.c_int), ('beam_width', C.c_int),
                ('search_width', C.c_int), ('spec_lanes', C.c_int), ('spec_top', C.c_int), ('spec_min_idle', C.c_int),
                ('spec_crit', C.c_int)]


class RoomStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('n_points', 'grow_steps', 'regions', 'clusters', 'stop_noneighbor',
                                          'stop_noexpand', 'stop_stuck', 'stop_other', 'spec_wasted_steps', 'spec_restarts',
                                          'spec_dropped', 'reserved')]


class StepTrace(C.Structure):
    _fields_ = [('seed_point', C.c_int32), ('step_in_region', C.c_int32), ('n_inlier', C.c_int32),
                ('n_neighbor', C.c_int32), ('stop_reason', C.c_int32), ('size_after', C.c_int32),
                ('center', C.c_float * 16), ('add_mask', C.c_uint32 * 16), ('remove_mask', C.c_uint32 * 16),
                ('inlier_idx_crc', C.c_uint32), ('neighbor_idx_crc', C.c_uint32), ('log_prob', C.c_float * 2), ('score', C.c_float),
                ('reserved', C.c_int32)]


STEP_TRACE_DTYPE = np.dtype([('seed_point', '<i4'), ('step_in_region', '<i4'), ('n_inlier', '<i4'), ('n_neighbor', '<i4'),
                             ('stop_reason', '<i4'), ('size_after', '<i4'), ('center', '<f4', (16,)),
                             ('add_mask', '<u4', (16,)), ('remove_mask', '<u4', (16,)),
                             ('inlier_idx_crc', '<u4'), ('neighbor_idx_crc', '<u4'), ('log_prob', '<f4', (2,)), ('score', '<f4'),
                             ('reserved', '<i4')])
ROOM_STATS_DTYPE = np.dtype([(n, '<i4') for n, _ in RoomStats._fields_])
ROOM_METRICS_DTYPE = np.dtype([('nmi', '<f8'), ('ami', '<f8'), ('ars', '<f8'), ('prc', '<f8'), ('rcl', '<f8'), ('iou', '<f8'),
                               ('n_points', '<i4'), ('n_classes', '<i4'), ('n_clusters', '<i4'), ('gt_match', '<i4')])
assert STEP_TRACE_DTYPE.itemsize == C.sizeof(StepTrace) and ROOM_STATS_DTYPE.itemsize == C.sizeof(RoomStats)

FORWARD_AUTO, FORWARD_FMA, FORWARD_TENSOR, FORWARD_TENSOR_F16 = 0, 1, 2, 3
FLAG_KERNEL_TIMING = 1
FLAG_NO_GRAPH = 2
FLAG_LOCKSTEP = 4
FLAG_PRIORITY = 8
FLAG_NO_PROJ_SERVERS = 16
FLAG_NO_TILE_SPLIT = 32
FLAG_HEADS_AFTER_PROJ = 64
FLAG_SCORE_ML = 128
FLAG_NO_SPATIAL_INDEX = 256
FLAG_NO_STEP_OVERLAP = 512
FLAG_ROOMS_IN_ORDER = 1024

_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    'lrg_last_error': (C.c_char_p, []),
    'lrg_version': (_I, []),
    'lrg_device_count': (_I, []),
    'lrg_engine_create': (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _I]),
    'lrg_engine_destroy': (_I, [_P]),
    'lrg_engine_weight_count': (C.c_size_t, [_P]),
    'lrg_engine_load_weights': (_I, [_P, _P, C.c_size_t]),
    'lrg_engine_set_forward_mode': (_I, [_P, _I]),
    'lrg_engine_forward_mode': (_I, [_P]),
    'lrg_engine_range_overflow': (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    'lrg_forward_host': (_I, [_P, _I, _P, _P, _P, _P]),
    'lrg_forward_device': (_I, [_P, _I, _P, _P, _P, _P, _P]),
    'lrg_rooms_upload': (_I, [_P, _I, _P, _P, _P, C.c_float]),
    'lrg_rooms_upload_raw': (_I, [_P, _I, _P, _P, _I, C.c_float]),
    'lrg_rooms_upload_raw_device': (_I, [_P, _I, _P, _P, _I, C.c_float]),
    'lrg_last_prepare_ms': (_I, [_P, C.POINTER(C.c_float)]),
    'lrg_last_prepare_launches': (_I, [_P, C.POINTER(_I)]),
    'lrg_rooms_equalized_offsets': (_I, [_P, _P]),
    'lrg_rooms_features_download': (_I, [_P, _P, _P, _P, _P]),
    'lrg_labels_download_raw': (_I, [_P, _P, _I]),
    'lrg_segmentation_metrics': (_I, [_I, _P, _P, _P, _P, _P, _P]),
    'lrg_room_metrics': (_I, [_P, _P, _I, _I, _P, _P]),
    'lrg_segment_resident': (_I, [_P, C.POINTER(GrowParams), _P]),
    'lrg_labels_download': (_I, [_P, _P, _I]),
    'lrg_trace_download': (_I, [_P, _I, _P, _I, C.POINTER(_I)]),
    'lrg_trace_download_lane': (_I, [_P, _I, _I, _P, _I, C.POINTER(_I)]),
    'lrg_segment_rooms_host': (_I, [_P, _I, _P, _P, _P, C.POINTER(GrowParams), _P, _P]),
    'lrg_last_segment_profile': (_I, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    'lrg_last_grow_profile': (_I, [_P, C.POINTER(_I), C.POINTER(C.c_double * 4), C.POINTER(C.c_int64 * 4)]),
    'lrg_engine_set_tile_timing': (_I, [_P, _I]),
    'lrg_tile_timing': (_I, [_P, C.POINTER(C.c_uint64 * 64), _I]),
    'lrg_last_grow_queue_delay': (_I, [_P, C.POINTER(C.c_double * 4)]),
    'lrg_last_kernel_times': (_I, [_P, C.POINTER(C.c_float * 4)]),
    'lrg_labels_device_ptr': (_I, [_P, _I, C.POINTER(_P)]),
    'lrg_farthest_point_sampling': (_I, [_I, _I, _I, _P, _P, _P, _P]),
    'lrg_gather_point': (_I, [_I, _I, _I, _P, _P, _P, _P]),
    'lrg_sample_and_group': (_I, [_I, _I, _I, C.c_float, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'lrg_fps_set_cluster_min': (_I, [_I]),
    'lrg_pairwise_sqdist': (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    'lrg_scatter_add_point': (_I, [_I, _I, _I, _P, _P, _P, _P]),
    'lrg_prob_sample': (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    'lrg_query_ball_point': (_I, [_I, _I, _I, C.c_float, _I, _P, _P, _P, _P, _P]),
    'lrg_selection_sort': (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    'lrg_group_point': (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    'lrg_group_point_grad': (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    'lrg_three_nn': (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    'lrg_three_interpolate': (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'lrg_three_interpolate_grad': (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'lrg_malloc': (_I, [C.POINTER(_P), C.c_size_t]),
    'lrg_free': (_I, [_P]),
    'lrg_memcpy_h2d': (_I, [_P, _P, C.c_size_t]),
    'lrg_memcpy_d2h': (_I, [_P, _P, C.c_size_t]),
    'lrg_memset': (_I, [_P, _I, C.c_size_t]),
    'lrg_device_synchronize': (_I, []),
    'lrg_set_device': (_I, [_I]),
    'lrg_host_alloc': (_I, [C.POINTER(_P), C.c_size_t]),
    'lrg_host_free': (_I, [_P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load the shared library (once).  Raises LrgError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LrgError('%s is missing: run `python -m learn_region_grow_b200.build` (there is no CPU fallback)' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise LrgError('liblrg_b200 error %d: %s' % (rc, lib().lrg_last_error().decode(errors='replace')))


def require_gpu():
    if lib().lrg_device_count() < 1:
        raise LrgError('no CUDA device visible: the LRGNet engine has no CPU fallback')


def ptr(a):
    """Device or host pointer of a numpy array / torch tensor / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, 'data_ptr'):
        return C.c_void_p(a.data_ptr())
    raise TypeError('cannot take a pointer of %r' % type(a))
