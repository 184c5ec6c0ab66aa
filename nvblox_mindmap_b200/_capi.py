"""ctypes binding of libnvbx.so (include/nvbx_c_api.h).

There is NO CPU fallback: if the CUDA library is missing or no Blackwell GPU is visible, every entry
point raises.  The library is built in-tree by `python -m nvblox_mindmap_b200.build`.
"""
import ctypes as C
import os
import threading

from nvblox_mindmap_b200 import build as _build
from nvblox_mindmap_b200.params import NvbxFrameJob, NvbxCounters, NvbxParams

_lock = threading.Lock()
_lib = None

# every symbol include/nvbx_c_api.h declares (tests check the .so exports all of them)
API_SYMBOLS = [
    'nvbx_default_params', 'nvbx_create', 'nvbx_destroy', 'nvbx_num_maps', 'nvbx_feature_channels',
    'nvbx_get_params', 'nvbx_last_error', 'nvbx_integrate_depth', 'nvbx_integrate_features',
    'nvbx_integrate_color', 'nvbx_integrate_frame_host', 'nvbx_set_host_fetch_mode', 'nvbx_set_pipelining',
    'nvbx_pipeline_join', 'nvbx_integrate_features_lowres',
    'nvbx_upsample_features', 'nvbx_integrate_frame_host_lowres', 'nvbx_decay', 'nvbx_clear', 'nvbx_mark_all_dirty',
    'nvbx_update_feature_mesh', 'nvbx_get_feature_mesh', 'nvbx_update_color_mesh', 'nvbx_get_color_mesh',
    'nvbx_export_points', 'nvbx_gather_points', 'nvbx_num_blocks', 'nvbx_num_allocated_blocks',
    'nvbx_num_allocated_bytes', 'nvbx_voxel_size', 'nvbx_get_block_indices', 'nvbx_get_all_blocks', 'nvbx_get_block_ptr',
    'nvbx_allocate_block', 'nvbx_query_tsdf', 'nvbx_query_features', 'nvbx_get_counters',
    'nvbx_reset_counters', 'nvbx_integrate_frames_batch', 'nvbx_set_gather_tuning', 'nvbx_set_kernel_timing', 'nvbx_get_kernel_timing', 'nvbx_kernel_timing_report', 'nvbx_kernel_launch_count', 'nvbx_pipeline_wait_stats', 'nvbx_debug_last_block_list', 'nvbx_debug_profile_stamps',
    'nvbx_debug_last_synthetic_depth', 'nvbx_version',
]


class NvbxError(RuntimeError):
    """A libnvbx call failed (message from nvbx_last_error())."""


def library_path() -> str:
    if os.environ.get('NVBX_PROFILE') == '1':      # tuning aid: in-kernel counters compiled in
        return _build.PROFILE_LIB_PATH
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libnvbx.so (building it with nvcc first if it is not there)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise NvbxError(f'{path} is missing: run `python -m nvblox_mindmap_b200.build` (no CPU fallback)')
            _build.build(profile=path == _build.PROFILE_LIB_PATH)
        L = C.CDLL(path)
        vp, fp, i32p, i64p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.nvbx_default_params.argtypes = [C.POINTER(NvbxParams)]
        L.nvbx_default_params.restype = None
        L.nvbx_create.argtypes = [C.c_int, fp, C.POINTER(NvbxParams), C.c_int, C.c_int, C.POINTER(vp)]
        L.nvbx_destroy.argtypes = [vp]
        L.nvbx_destroy.restype = None
        L.nvbx_num_maps.argtypes = [vp]
        L.nvbx_feature_channels.argtypes = [vp]
        L.nvbx_get_params.argtypes = [vp, C.POINTER(NvbxParams)]
        L.nvbx_last_error.restype = C.c_char_p
        # poses are passed as the address of 16 row-major floats (a CPU tensor's data_ptr())
        frame = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp]
        L.nvbx_integrate_depth.argtypes = frame
        L.nvbx_integrate_color.argtypes = frame
        L.nvbx_integrate_features.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_float,
                                              C.c_float, C.c_float, C.c_float, vp]
        L.nvbx_integrate_frame_host.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                                C.c_float, C.c_float, C.c_float, C.c_float, vp]
        L.nvbx_integrate_features_lowres.argtypes = [vp, C.c_int, vp] + [C.c_int] * 8 + [vp, vp] + [C.c_float] * 4 + [vp]
        L.nvbx_upsample_features.argtypes = [vp, C.c_int, vp] + [C.c_int] * 8 + [vp, vp]
        L.nvbx_integrate_frame_host_lowres.argtypes = [vp, C.c_int, vp, vp] + [C.c_int] * 8 + [vp, vp, vp] + \
            [C.c_float] * 4 + [vp]
        for n in ('nvbx_decay', 'nvbx_clear', 'nvbx_mark_all_dirty', 'nvbx_update_feature_mesh',
                  'nvbx_update_color_mesh'):
            getattr(L, n).argtypes = [vp, C.c_int, vp]
        L.nvbx_get_feature_mesh.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), i64p, i64p]
        L.nvbx_get_color_mesh.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), i64p, i64p]
        L.nvbx_export_points.argtypes = [vp, C.c_int, vp, vp, C.c_int64, C.c_int, fp, fp, C.c_int, C.c_int,
                                         C.POINTER(vp), C.POINTER(vp), vp]
        L.nvbx_export_points.restype = C.c_int64
        L.nvbx_gather_points.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int64, vp, vp, C.c_int, vp]
        for n in ('nvbx_num_blocks', 'nvbx_num_allocated_blocks', 'nvbx_num_allocated_bytes'):
            getattr(L, n).argtypes = [vp, C.c_int, C.c_int, vp]
            getattr(L, n).restype = C.c_int64
        L.nvbx_voxel_size.argtypes = [vp, C.c_int]
        L.nvbx_voxel_size.restype = C.c_float
        L.nvbx_get_block_indices.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64, vp]
        L.nvbx_get_block_indices.restype = C.c_int64
        L.nvbx_get_all_blocks.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int64, i64p, vp]
        L.nvbx_get_all_blocks.restype = C.c_int64
        L.nvbx_get_block_ptr.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), i64p, vp]
        L.nvbx_allocate_block.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.nvbx_query_tsdf.argtypes = [vp, C.c_int, vp, C.c_int64, vp, vp]
        L.nvbx_query_features.argtypes = [vp, C.c_int, vp, C.c_int64, vp, vp]
        L.nvbx_set_host_fetch_mode.argtypes = [vp, C.c_int]
        L.nvbx_set_pipelining.argtypes = [vp, C.c_int]
        L.nvbx_pipeline_join.argtypes = [vp, C.c_int, vp]
        L.nvbx_get_counters.argtypes = [vp, C.c_int, C.POINTER(NvbxCounters), vp]
        L.nvbx_reset_counters.argtypes = [vp, C.c_int, vp]
        L.nvbx_set_gather_tuning.argtypes = [C.c_int, C.c_int, C.c_int]
        L.nvbx_integrate_frames_batch.argtypes = [C.POINTER(NvbxFrameJob), C.c_int, C.c_int]
        L.nvbx_set_kernel_timing.argtypes = [vp, C.c_int]
        L.nvbx_get_kernel_timing.argtypes = [vp, C.c_int, C.POINTER(C.c_double), i64p]
        L.nvbx_kernel_timing_report.argtypes = [vp, C.c_char_p, C.c_int64]
        L.nvbx_kernel_timing_report.restype = C.c_int64
        L.nvbx_kernel_launch_count.restype = C.c_int64
        L.nvbx_pipeline_wait_stats.argtypes = [i64p, i64p]
        L.nvbx_pipeline_wait_stats.restype = None
        L.nvbx_debug_last_block_list.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64, vp]
        L.nvbx_debug_last_block_list.restype = C.c_int64
        L.nvbx_debug_last_synthetic_depth.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int),
                                                      C.POINTER(C.c_int)]
        L.nvbx_debug_profile_stamps.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
        L.nvbx_version.restype = C.c_char_p
        _lib = L
        return _lib


def check(rc: int) -> int:
    """Raise NvbxError for a negative status code."""
    if rc is not None and rc < 0:
        msg = load().nvbx_last_error()
        raise NvbxError(f'libnvbx error {rc}: {msg.decode() if msg else "?"}')
    return rc


def default_params() -> NvbxParams:
    p = NvbxParams()
    load().nvbx_default_params(C.byref(p))
    return p
