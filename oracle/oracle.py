"""ctypes wrapper around oracle/libnvbx_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path never does (see the header of nvbx_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from nvblox_mindmap_b200.params import NvbxCounters, NvbxParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libnvbx_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(['make', '-C', _HERE] + (['-B'] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, u8p, u16p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16),
                               C.POINTER(C.c_int32))
        L.orc_default_params.argtypes = [C.POINTER(NvbxParams)]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_float, C.c_int, C.POINTER(NvbxParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_fused_half.argtypes = [C.c_int]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_max_threads.restype = C.c_int
        frame = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, fp, C.c_float, C.c_float, C.c_float,
                 C.c_float]
        L.orc_integrate_depth.argtypes = frame
        L.orc_integrate_features.argtypes = frame
        L.orc_integrate_color.argtypes = frame
        for n in ('orc_decay', 'orc_clear', 'orc_update_feature_mesh', 'orc_update_color_mesh', 'orc_mark_all_dirty',
                  'orc_reset_counters'):
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_color_mesh_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_color_mesh_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_all_color_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mesh_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_mesh_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_num_blocks.restype = C.c_int64
        L.orc_num_blocks.argtypes = [C.c_void_p, C.c_int]
        L.orc_block_indices.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_get_all_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_set_tsdf_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_query_tsdf.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_query_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_get_counters.argtypes = [C.c_void_p, C.POINTER(NvbxCounters)]
        L.orc_last_distinct_pixels.restype = C.c_int64
        L.orc_last_distinct_pixels.argtypes = [C.c_void_p]
        L.orc_last_feature_voxels.restype = C.c_int64
        L.orc_last_feature_voxels.argtypes = [C.c_void_p]
        for fn in (L.orc_last_trace_steps, L.orc_last_trace_max_steps):
            fn.restype, fn.argtypes = C.c_int64, [C.c_void_p]
        L.orc_last_block_list.restype = C.c_int64
        L.orc_last_block_list.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.orc_last_synthetic_depth.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        # unit hooks
        L.orc_f2h.restype = C.c_uint16
        L.orc_upsample_bilinear.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.orc_upsample_bilinear.restype = None
        L.orc_f2h.argtypes = [C.c_float]
        L.orc_h2f.restype = C.c_float
        L.orc_h2f.argtypes = [C.c_uint16]
        for n in ('orc_hadd', 'orc_hsub', 'orc_hmul'):
            getattr(L, n).restype = C.c_uint16
            getattr(L, n).argtypes = [C.c_uint16, C.c_uint16]
        L.orc_interp_half.restype = C.c_uint16
        L.orc_interp_half.argtypes = [C.c_float, C.c_float] + [C.c_uint16] * 4
        L.orc_interp_float.restype = C.c_float
        L.orc_interp_float.argtypes = [C.c_float] * 6
        L.orc_weighting.restype = C.c_float
        L.orc_weighting.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_raycast.restype = C.c_int
        L.orc_raycast.argtypes = [fp, fp, i32p, C.c_int]
        L.orc_voxel_center.argtypes = [C.c_float, i32p, i32p, fp]
        L.orc_block_and_voxel.argtypes = [C.c_float, fp, i32p, i32p]
        L.orc_poses_close.restype = C.c_int
        L.orc_poses_close.argtypes = [fp, fp, C.c_float, C.c_float]
        L.orc_project.restype = C.c_int
        L.orc_project.argtypes = [C.c_float] * 4 + [C.c_int, C.c_int, fp, fp]
        L.orc_weld_key.restype = C.c_uint64
        L.orc_weld_key.argtypes = [fp]
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def default_params() -> NvbxParams:
    p = NvbxParams()
    lib().orc_default_params(C.byref(p))
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleMapper:
    """One TSDF + feature map integrated on the CPU with the reference's semantics."""

    def __init__(self, voxel_size_m: float, feature_channels: int, params: NvbxParams = None):
        self.L = lib()
        self.params = params if params is not None else default_params()
        self.C = int(feature_channels)
        self.voxel_size = float(np.float32(voxel_size_m))
        self.h = self.L.orc_create(C.c_float(voxel_size_m), self.C, C.byref(self.params))

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.orc_destroy(self.h)
            self.h = None

    # -- frame integration ----------------------------------------------------------------------
    def add_depth_frame(self, depth, t_w_c, intrinsics, mask=None):
        depth = _f32(depth)
        T = _f32(t_w_c).reshape(16)
        K = _f32(intrinsics)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_integrate_depth(self.h, _ptr(depth), depth.shape[0], depth.shape[1], _ptr(m),
                                   T.ctypes.data_as(C.POINTER(C.c_float)), K[0, 0], K[1, 1], K[0, 2], K[1, 2])

    def add_feature_frame(self, feat, t_w_c, intrinsics, mask=None):
        feat = np.ascontiguousarray(feat)
        assert feat.dtype == np.float16 and feat.shape[2] == self.C
        T = _f32(t_w_c).reshape(16)
        K = _f32(intrinsics)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_integrate_features(self.h, _ptr(feat), feat.shape[0], feat.shape[1], _ptr(m),
                                      T.ctypes.data_as(C.POINTER(C.c_float)), K[0, 0], K[1, 1], K[0, 2], K[1, 2])

    def add_color_frame(self, rgb, t_w_c, intrinsics, mask=None):
        rgb = np.ascontiguousarray(rgb)
        assert rgb.dtype == np.uint8 and rgb.shape[2] == 3
        T = _f32(t_w_c).reshape(16)
        K = _f32(intrinsics)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_integrate_color(self.h, _ptr(rgb), rgb.shape[0], rgb.shape[1], _ptr(m),
                                   T.ctypes.data_as(C.POINTER(C.c_float)), K[0, 0], K[1, 1], K[0, 2], K[1, 2])

    def decay(self):
        self.L.orc_decay(self.h)

    def clear(self):
        self.L.orc_clear(self.h)

    # -- mesh -----------------------------------------------------------------------------------
    def update_feature_mesh(self):
        self.L.orc_update_feature_mesh(self.h)

    def get_feature_mesh(self, with_block_index: bool = False):
        nv, nt = C.c_int64(), C.c_int64()
        self.L.orc_mesh_sizes(self.h, C.byref(nv), C.byref(nt))
        verts = np.zeros((nv.value, 3), np.float32)
        feats = np.zeros((nv.value, self.C), np.float16)
        tris = np.zeros((nt.value // 3, 3), np.int32)
        vb = np.zeros((nv.value, 3), np.int32) if with_block_index else None
        self.L.orc_mesh_copy(self.h, _ptr(verts), _ptr(feats), _ptr(tris), _ptr(vb))
        return (verts, feats, tris, vb) if with_block_index else (verts, feats, tris)

    def update_color_mesh(self):
        self.L.orc_update_color_mesh(self.h)

    def get_color_mesh(self):
        nv, nt = C.c_int64(), C.c_int64()
        self.L.orc_color_mesh_sizes(self.h, C.byref(nv), C.byref(nt))
        verts = np.zeros((nv.value, 3), np.float32)
        cols = np.zeros((nv.value, 3), np.uint8)
        tris = np.zeros((nt.value // 3, 3), np.int32)
        self.L.orc_color_mesh_copy(self.h, _ptr(verts), _ptr(cols), _ptr(tris))
        return verts, cols, tris

    def all_color_blocks(self):
        """(indices [N,3] sorted, rgb [N,8,8,8,3] u8, weight [N,8,8,8] f32)."""
        idx = self.block_indices(2)
        rgb = np.zeros((len(idx), 8, 8, 8, 3), np.uint8)
        w = np.zeros((len(idx), 8, 8, 8), np.float32)
        self.L.orc_get_all_color_blocks(self.h, _ptr(rgb), _ptr(w))
        return idx, rgb, w

    # -- layers ---------------------------------------------------------------------------------
    def num_blocks(self, layer: int) -> int:
        return int(self.L.orc_num_blocks(self.h, layer))

    def block_indices(self, layer: int) -> np.ndarray:
        out = np.zeros((self.num_blocks(layer), 3), np.int32)
        self.L.orc_block_indices(self.h, layer, _ptr(out))
        return out

    def all_blocks(self, layer: int):
        """(indices [N,3] sorted, data): TSDF [N,8,8,8,2] f32 / feature [N,8,8,8,C+1] f16."""
        idx = self.block_indices(layer)
        if layer == 0:
            data = np.zeros((len(idx), 8, 8, 8, 2), np.float32)
        else:
            data = np.zeros((len(idx), 8, 8, 8, self.C + 1), np.float16)
        self.L.orc_get_all_blocks(self.h, layer, _ptr(data))
        return idx, data

    def set_tsdf_block(self, index, data):
        data = _f32(data).reshape(512 * 2)
        self.L.orc_set_tsdf_block(self.h, int(index[0]), int(index[1]), int(index[2]), _ptr(data))

    def mark_all_dirty(self):
        self.L.orc_mark_all_dirty(self.h)

    def query_tsdf(self, xyz):
        xyz = _f32(xyz)
        out = np.zeros((len(xyz), 2), np.float32)
        self.L.orc_query_tsdf(self.h, _ptr(xyz), len(xyz), _ptr(out))
        return out

    def query_features(self, xyz):
        xyz = _f32(xyz)
        out = np.zeros((len(xyz), self.C + 1), np.float16)
        self.L.orc_query_features(self.h, _ptr(xyz), len(xyz), _ptr(out))
        return out

    # -- accounting / stage-level hooks ------------------------------------------------------------
    def counters(self) -> dict:
        c = NvbxCounters()
        self.L.orc_get_counters(self.h, C.byref(c))
        d = c.as_dict()
        d['last_distinct_pixels'] = int(self.L.orc_last_distinct_pixels(self.h))
        d['last_feature_voxels'] = int(self.L.orc_last_feature_voxels(self.h))
        d['last_trace_steps'] = int(self.L.orc_last_trace_steps(self.h))
        d['last_trace_max_steps'] = int(self.L.orc_last_trace_max_steps(self.h))
        return d

    def reset_counters(self):
        self.L.orc_reset_counters(self.h)

    def last_block_list(self, which: int) -> np.ndarray:
        n = int(self.L.orc_last_block_list(self.h, which, None, 0))
        out = np.zeros((n, 3), np.int32)
        self.L.orc_last_block_list(self.h, which, _ptr(out), n)
        return out

    def last_synthetic_depth(self) -> np.ndarray:
        r, c = C.c_int(), C.c_int()
        self.L.orc_last_synthetic_depth(self.h, None, C.byref(r), C.byref(c))
        out = np.zeros((r.value, c.value), np.float32)
        if out.size:
            self.L.orc_last_synthetic_depth(self.h, _ptr(out), C.byref(r), C.byref(c))
        return out


def upsample_bilinear(low_hwc: np.ndarray, channels: int, height: int, width: int, mode: int = 0) -> np.ndarray:
    """N4: the [H, W, C] binary16 frame (as uint16 bits) mindmap's FeatureExtractor.compute() + `.to(float16)`
    produce from the backbone's [lh, lw, lc] fp32 feature map (torch CUDA bilinear up-sampling, align_corners=False,
    zero-padded to `channels`).  mode: 0 torch NCHW kernel, 1 torch NHWC fp32 kernel, 2 bf16 tensors."""
    low = np.ascontiguousarray(low_hwc, dtype=np.float32)
    lh, lw, lc = low.shape
    assert lc <= channels
    out = np.empty((height, width, channels), np.uint16)
    lib().orc_upsample_bilinear(_ptr(low), lh, lw, lc, channels, height, width, mode, _ptr(out))
    return out
