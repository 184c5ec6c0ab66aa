"""Shared helpers of the parity tests: run the same frames through the CUDA path and the CPU oracle."""
import ctypes as C

import numpy as np

from oracle import oracle as O
from tests import scenes as S


def make_params(workspace=None, max_dist=5.0, alpha=1.0, raycast_sub=1, decay=0.98, strict=False, trunc_vox=4.0,
                cache=True, weighting='kInverseSquareWeight'):
    """(nvblox_torch MapperParams, NvbxParams for the oracle) with mindmap's settings
    (mindmap/mapping/helpers/nvblox_mapping_helpers.py:40-70)."""
    from nvblox_torch.mapper_params import (BlockMemoryPoolParams, MapperParams, ProjectiveIntegratorParams,
                                            TsdfDecayIntegratorParams, ViewCalculatorParams)
    pi = ProjectiveIntegratorParams()
    pi.projective_integrator_max_integration_distance_m = max_dist
    pi.projective_appearance_integrator_measurement_weight = alpha
    pi.projective_integrator_truncation_distance_vox = trunc_vox
    pi.projective_integrator_weighting_mode = weighting
    td = TsdfDecayIntegratorParams()
    td.tsdf_decay_factor = decay
    vc = ViewCalculatorParams()
    vc.raycast_subsampling_factor = raycast_sub
    if workspace is not None:
        vc.workspace_bounds_type = 'kBoundingBox'
        (vc.workspace_bounds_min_corner_x_m, vc.workspace_bounds_min_corner_y_m, vc.workspace_bounds_min_height_m) = \
            [float(v) for v in workspace[0]]
        (vc.workspace_bounds_max_corner_x_m, vc.workspace_bounds_max_corner_y_m, vc.workspace_bounds_max_height_m) = \
            [float(v) for v in workspace[1]]
    bp = BlockMemoryPoolParams()
    bp.expansion_factor = 1.0
    bp.num_preallocated_blocks = 0
    mp = MapperParams()
    mp.set_projective_integrator_params(pi)
    mp.set_tsdf_decay_integrator_params(td)
    mp.set_view_calculator_params(vc)
    mp.set_block_memory_pool_params(bp)
    mp.strict_blend = strict
    mp.cache_last_viewpoint = cache
    return mp, oracle_params_from(mp)


def oracle_params_from(mp):
    """The same POD the product receives, built WITHOUT loading libnvbx (oracle defaults + fields)."""
    p = O.default_params()
    pi, vc, td = mp._projective_integrator_params, mp._view_calculator_params, mp._tsdf_decay_integrator_params
    from nvblox_mindmap_b200.params import WEIGHTING_MODES, WORKSPACE_BOUNDS_TYPES
    p.max_integration_distance_m = pi.projective_integrator_max_integration_distance_m
    p.truncation_distance_vox = pi.projective_integrator_truncation_distance_vox
    p.weighting_mode = WEIGHTING_MODES.get(pi.projective_integrator_weighting_mode, 4)
    p.max_weight = pi.projective_integrator_max_weight
    p.invalid_depth_decay_factor = pi.projective_tsdf_integrator_invalid_depth_decay_factor
    p.appearance_measurement_weight = pi.projective_appearance_integrator_measurement_weight
    p.appearance_truncation_distance_vox = mp.appearance_truncation_distance_vox
    p.tsdf_decay_factor = td.tsdf_decay_factor
    p.tsdf_decayed_weight_threshold = td.tsdf_decayed_weight_threshold
    p.raycast_subsampling_factor = int(vc.raycast_subsampling_factor)
    p.workspace_bounds_type = WORKSPACE_BOUNDS_TYPES[vc.workspace_bounds_type]
    p.workspace_min[0], p.workspace_min[1], p.workspace_min[2] = (vc.workspace_bounds_min_corner_x_m,
                                                                 vc.workspace_bounds_min_corner_y_m,
                                                                 vc.workspace_bounds_min_height_m)
    p.workspace_max[0], p.workspace_max[1], p.workspace_max[2] = (vc.workspace_bounds_max_corner_x_m,
                                                                 vc.workspace_bounds_max_corner_y_m,
                                                                 vc.workspace_bounds_max_height_m)
    p.cache_last_viewpoint = int(mp.cache_last_viewpoint)
    p.strict_blend = int(mp.strict_blend)
    return p


def sort_rows(a: np.ndarray) -> np.ndarray:
    """Rows of a 2-D array in lexicographic order (canonical form for unordered comparisons)."""
    if a.shape[0] == 0:
        return a
    keys = a.view(np.uint8).reshape(a.shape[0], -1) if a.dtype.kind == 'f' else a.reshape(a.shape[0], -1)
    order = np.lexsort(keys.T[::-1])
    return a[order]


def gpu_blocks(layer_view):
    """(sorted indices [N,3], data [N,8,8,8,E]) of a CUDA layer view, in index order."""
    import torch
    blocks, indices = layer_view.get_all_blocks()
    if not blocks:
        return np.zeros((0, 3), np.int32), None
    idx = torch.stack(indices).numpy().astype(np.int32)
    data = torch.stack([b.contiguous() for b in blocks]).cpu().numpy()
    order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))
    return idx[order], data[order]


def canonical_mesh(verts, feats, tris):
    """Order-independent representation: per-vertex rows (xyz bits + feature bits) sorted, and triangles
    expanded to 9 coordinates and sorted."""
    v = np.ascontiguousarray(verts, np.float32)
    f = np.ascontiguousarray(feats, np.float16)
    rows = np.concatenate([v.view(np.uint32).astype(np.uint64), f.view(np.uint16).astype(np.uint64)], axis=1)
    rows = rows[np.lexsort(rows.T[::-1])] if len(rows) else rows
    t = np.ascontiguousarray(tris, np.int64)
    tri_xyz = v[t.reshape(-1)].reshape(-1, 9).view(np.uint32) if len(t) else np.zeros((0, 9), np.uint32)
    tri_xyz = tri_xyz[np.lexsort(tri_xyz.T[::-1])] if len(tri_xyz) else tri_xyz
    return rows, tri_xyz


class Pair:
    """The CUDA Mapper (through nvblox_torch -> C ABI) and the CPU oracle fed identically."""

    def __init__(self, voxel_size, C_feat, mp_params, orc_params, device=0):
        import torch
        from nvblox_torch.constants import constants
        from nvblox_torch.mapper import Mapper
        constants.set_feature_array_num_elements(C_feat)
        self.torch = torch
        self.dev = f'cuda:{device}'
        self.C = C_feat
        self.gpu = Mapper(voxel_sizes_m=float(voxel_size), mapper_parameters=mp_params)
        self.cpu = O.OracleMapper(voxel_size, C_feat, orc_params)

    def depth(self, depth, T, K, mask=None):
        t = self.torch
        self.gpu.add_depth_frame(t.from_numpy(depth).to(self.dev), t.from_numpy(T), t.from_numpy(K),
                                 None if mask is None else t.from_numpy(mask).to(self.dev))
        self.cpu.add_depth_frame(depth, T, K, mask)

    def features(self, feat, T, K, mask=None):
        t = self.torch
        self.gpu.add_feature_frame(t.from_numpy(feat).to(self.dev), t.from_numpy(T), t.from_numpy(K),
                                   None if mask is None else t.from_numpy(mask).to(self.dev))
        self.cpu.add_feature_frame(feat, T, K, mask)

    def color(self, rgb, T, K, mask=None):
        t = self.torch
        self.gpu.add_color_frame(t.from_numpy(rgb).to(self.dev), t.from_numpy(T), t.from_numpy(K),
                                 None if mask is None else t.from_numpy(mask).to(self.dev))
        self.cpu.add_color_frame(rgb, T, K, mask)

    def decay(self):
        self.gpu.decay()
        self.cpu.decay()

    def clear(self):
        self.gpu.clear()
        self.cpu.clear()

    def last_block_list(self, which):
        from nvblox_mindmap_b200 import _capi
        L = _capi.load()
        n = int(_capi.check(L.nvbx_debug_last_block_list(self.gpu._handle, 0, which, None, 0, self.gpu._stream())))
        out = np.zeros((max(n, 1), 3), np.int32)
        _capi.check(L.nvbx_debug_last_block_list(self.gpu._handle, 0, which, out.ctypes.data_as(C.c_void_p), n,
                                                 self.gpu._stream()))
        return sort_rows(out[:n]), sort_rows(self.cpu.last_block_list(which))

    def synthetic_depth(self):
        from nvblox_mindmap_b200 import _capi
        from nvblox_mindmap_b200.torch_interop import device_view
        L = _capi.load()
        p, r, c = C.c_void_p(), C.c_int(), C.c_int()
        _capi.check(L.nvbx_debug_last_synthetic_depth(self.gpu._handle, 0, C.byref(p), C.byref(r), C.byref(c)))
        g = device_view(p.value, (r.value, c.value), self.torch.float32, self.gpu._device, owner=self.gpu)
        return g.cpu().numpy(), self.cpu.last_synthetic_depth()

    # ---- comparisons ------------------------------------------------------------------------------
    def check_tsdf(self, exact=True):
        gi, gd = gpu_blocks(self.gpu.tsdf_layer_view(0))
        ci, cd = self.cpu.all_blocks(0)
        assert np.array_equal(gi, ci), f'TSDF block sets differ: gpu {len(gi)} vs oracle {len(ci)}'
        if len(gi) == 0:
            return 0
        if exact:
            bad = np.argwhere(gd.view(np.uint32) != cd.view(np.uint32))
            assert len(bad) == 0, f'{len(bad)} TSDF values differ bitwise; first {bad[:3]}'
        else:    # north_star tolerance: 1e-5 relative in fp32
            np.testing.assert_allclose(gd, cd, rtol=1e-5, atol=1e-7)
        return len(gi)

    def check_features(self, max_ulp=0):
        gi, gd = gpu_blocks(self.gpu.feature_layer_view(0))
        ci, cd = self.cpu.all_blocks(1)
        assert np.array_equal(gi, ci), f'feature block sets differ: gpu {len(gi)} vs oracle {len(ci)}'
        if len(gi) == 0:
            return 0
        g16, c16 = gd.view(np.uint16), cd.view(np.uint16)
        # updated-voxel set and weights: exact
        assert np.array_equal(g16[..., -1], c16[..., -1]), 'feature weights differ'
        if max_ulp == 0:
            bad = np.argwhere(g16 != c16)
            assert len(bad) == 0, f'{len(bad)} feature halves differ bitwise; first {bad[:3]}'
        else:
            ulp = np.abs(ordered_half(g16).astype(np.int32) - ordered_half(c16).astype(np.int32))
            assert ulp.max() <= max_ulp, f'max fp16 ulp distance {ulp.max()} > {max_ulp}'
        return len(gi)

    def check_color(self):
        """Colour layer: block set, RGB bytes and float weights bit-exact."""
        import torch
        from nvblox_mindmap_b200.torch_interop import device_view
        layer = self.gpu.color_layer_view(0)
        gi, gd = gpu_blocks(layer)
        ci, crgb, cw = self.cpu.all_color_blocks()
        assert np.array_equal(gi, ci), f'colour block sets differ: gpu {len(gi)} vs oracle {len(ci)}'
        if len(gi) == 0:
            return 0
        assert np.array_equal(gd, crgb), f'{int((gd != crgb).any(-1).sum())} colour voxels differ'
        # the weights sit behind the RGB bytes of each 8-byte ColorVoxel: read them through the raw block view
        for k, row in enumerate(gi):
            blk = layer.get_block_at_index(torch.from_numpy(row))
            raw = device_view(blk.data_ptr(), (8, 8, 8, 8), torch.uint8, self.gpu._device, owner=self.gpu).cpu().numpy()
            w = raw[..., 4:8].copy().view(np.float32)[..., 0]
            assert np.array_equal(w.view(np.uint32), cw[k].view(np.uint32)), f'colour weights differ in block {row}'
            assert not raw[..., 3].any(), 'padding byte of a ColorVoxel is not zero'
        return len(gi)

    def check_color_mesh(self):
        self.gpu.update_color_mesh(0)
        self.cpu.update_color_mesh()
        m = self.gpu.get_color_mesh(0)
        gv, gc, gt = (m.vertices().cpu().numpy(), m.vertex_colors().cpu().numpy(), m.triangles().cpu().numpy())
        cv, cc, ct = self.cpu.get_color_mesh()
        assert gv.shape == cv.shape, f'vertex count: gpu {gv.shape} vs oracle {cv.shape}'
        assert gt.shape == ct.shape, f'triangle count: gpu {gt.shape} vs oracle {ct.shape}'
        # colours ride through canonical_mesh's 16-bit appearance slot
        gr, gx = canonical_mesh(gv, gc.astype(np.uint16).view(np.float16), gt)
        cr, cx = canonical_mesh(cv, cc.astype(np.uint16).view(np.float16), ct)
        assert np.array_equal(gr, cr), 'colour mesh vertices / vertex colours differ'
        assert np.array_equal(gx, cx), 'colour mesh triangles differ'
        return len(gv)

    def check_mesh(self):
        self.gpu.update_feature_mesh(0)
        self.cpu.update_feature_mesh()
        m = self.gpu.get_feature_mesh(0)
        gv, gf, gt = (m.vertices().cpu().numpy(), m.vertex_features().cpu().numpy(), m.triangles().cpu().numpy())
        cv, cf, ct = self.cpu.get_feature_mesh()
        assert gv.shape == cv.shape, f'vertex count: gpu {gv.shape} vs oracle {cv.shape}'
        assert gt.shape == ct.shape, f'triangle count: gpu {gt.shape} vs oracle {ct.shape}'
        gr, gx = canonical_mesh(gv, gf, gt)
        cr, cx = canonical_mesh(cv, cf, ct)
        assert np.array_equal(gr, cr), 'mesh vertices / vertex features differ'
        assert np.array_equal(gx, cx), 'mesh triangles differ'
        return len(gv)


def ordered_half(u16: np.ndarray) -> np.ndarray:
    """Map fp16 bit patterns to integers that are monotone in value (for ulp distances)."""
    u = u16.astype(np.int32)
    return np.where(u & 0x8000, 0x8000 - (u & 0x7fff), 0x8000 + u)


def orbit_frames(n, H, W, C_feat, scene, radius=0.45, height=0.5, center=(0.35, 0.0, 0.1), n_orbit=64, seed0=1000):
    K = S.intrinsics(W, H)
    for i in range(n):
        T = S.orbit_pose(i, n_orbit, radius, height, center)
        depth = S.render_depth(K, H, W, T, **scene)
        feat = S.feature_frame(H, W, C_feat, seed0 + i)
        yield i, T, K, depth, feat
