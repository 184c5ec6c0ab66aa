"""Seeded inputs and frame sequences behind the reference-compiled golden vectors.

Shared by tests/golden/make_ref_vectors.py (which runs them through the REFERENCE's own kernels on a B200 and
writes tests/golden/ref_nvcc_*.npz) and by tests/test_ref_vectors.py / tests/test_gpu_ref_vectors.py (which run
them through the CPU oracle and through the CUDA product and compare bit for bit).  numpy only.
"""
import ctypes as C

import numpy as np

from tests import scenes as S

# -------------------------------------------------------------------------------------------------------------
# function-level probe inputs
# -------------------------------------------------------------------------------------------------------------
N_FN = 8192


def _half_values(rng, n):
    """binary16 values: mostly N(0,1) (what a backbone emits), plus large / tiny / subnormal / exact-zero entries."""
    v = rng.standard_normal(n).astype(np.float32)
    k = n // 16
    v[:k] *= 1000.0                      # large: products overflow the half range in places
    v[k:2 * k] *= 1e-4                   # tiny
    v[2 * k:3 * k] *= 3e-7               # subnormal halves
    v[3 * k:3 * k + 32] = 0.0
    v[3 * k + 32:3 * k + 40] = [65504.0, -65504.0, 6e-8, -6e-8, 1.0, -1.0, 0.5, 2.0]
    return rng.permutation(v).astype(np.float16)


def function_inputs(channels: int) -> dict:
    rng = np.random.default_rng(20261017)
    d = {}
    xy = rng.random((N_FN, 2), dtype=np.float32)
    xy[:64] = rng.integers(0, 2, (64, 2)).astype(np.float32)            # corners
    xy[64:128, 0] = np.float32(0.5)
    xy[128:192] = (rng.integers(0, 2048, (64, 2)) / 2048.0).astype(np.float32)
    d['interp_xy'] = xy
    d['interp_half_f'] = _half_values(rng, N_FN * 4).reshape(N_FN, 4)
    ff = rng.standard_normal((N_FN, 4)).astype(np.float32)
    ff[: N_FN // 2] = (np.float32(0.8) + np.float32(0.4) * rng.random((N_FN // 2, 4), dtype=np.float32))  # depths
    ff[:256] = -1.0                                                     # sphere-trace misses
    ff[256:512, rng.integers(0, 4, 256)] = -1.0
    d['interp_float_f'] = ff
    nb = 192
    d['blend_a'] = _half_values(rng, nb * channels).reshape(nb, channels)
    d['blend_b'] = _half_values(rng, nb * channels).reshape(nb, channels)
    alpha = np.concatenate([np.full(32, 1.0), np.full(32, 0.8), np.full(32, 0.3), np.full(32, 0.5),
                            rng.random(64)]).astype(np.float32)
    d['blend_w'] = np.stack([np.float32(1.0) - alpha, alpha], axis=1).astype(np.float32)
    # marching-cubes vertex interpolation: neighbouring voxel centres of a 2 cm grid and TSDF pairs
    vs = np.float32(0.02)
    base = (rng.integers(-40, 40, (N_FN, 3)).astype(np.float32) + np.float32(0.5)) * vs
    axis = rng.integers(0, 3, N_FN)
    step = np.zeros((N_FN, 3), np.float32)
    step[np.arange(N_FN), axis] = vs
    d['vertex_v1'] = base.astype(np.float32)
    d['vertex_v2'] = (base + step).astype(np.float32)
    sdf = (rng.random((N_FN, 2), dtype=np.float32) * np.float32(0.08)).astype(np.float32)
    sdf[:, 1] *= -1.0
    sdf[:128, 1] = sdf[:128, 0] - np.float32(5e-5)                      # below kMinSdfDifference
    sw = rng.random(N_FN) < 0.5
    sdf[sw] = sdf[sw][:, ::-1]
    d['vertex_sdf'] = np.ascontiguousarray(sdf)
    # positions -> (block, voxel): random, plus points on and next to block / voxel boundaries
    d['block_sizes'] = np.asarray([0.16, 0.08, 0.4], np.float32)
    for k, bs in enumerate(d['block_sizes']):
        p = (rng.random((N_FN, 3), dtype=np.float32) * np.float32(4.0) - np.float32(2.0)).astype(np.float32)
        kk = rng.integers(-200, 200, (2048, 3)).astype(np.float32)
        edge = (kk * np.float32(bs / np.float32(8))).astype(np.float32)
        p[:2048] = edge
        p[2048:3072] = np.nextafter(edge[:1024], np.float32(10))
        p[3072:4096] = np.nextafter(edge[1024:], np.float32(-10))
        d[f'bv_points_{k}'] = p
    # projectThreadVoxel
    poses = [S.orbit_pose(3), S.orbit_pose(17, radius=0.6, height=0.35), S.look_at((-0.2, 0.0, 0.6), (0.35, 0.0, 0.1))]
    cams = []
    for k, T in enumerate(poses):
        W = H = (512, 96, 1024)[k]
        Kc = S.intrinsics(W, H)
        bs = (0.16, 0.08, 0.08)[k]
        cams.append((Kc[0, 0], Kc[1, 1], Kc[0, 2], Kc[1, 2], W, H, bs, 5.0 if k < 2 else 0.9))
        nblk = int(np.ceil(1.5 / bs))
        b = rng.integers(-nblk, nblk, (N_FN, 3))
        v = rng.integers(0, 8, (N_FN, 3))
        d[f'project_bv_{k}'] = np.concatenate([b, v], axis=1).astype(np.int32)
    d['project_poses'] = np.stack([np.linalg.inv(T.astype(np.float64)).astype(np.float32) for T in poses])
    d['project_cams'] = np.asarray(cams, np.float32)
    # UpdateTsdfVoxelFunctor
    meas = (np.float32(0.3) + rng.random(N_FN, dtype=np.float32) * np.float32(1.2)).astype(np.float32)
    vd = (meas + (rng.random(N_FN, dtype=np.float32) - np.float32(0.5)) * np.float32(0.3)).astype(np.float32)
    meas[:256] = 0.0                                                    # invalid depth
    meas[256:300] = -1.0
    d['tsdf_in'] = np.stack([meas, vd], axis=1).astype(np.float32)
    d['tsdf_active'] = (rng.random(N_FN) < 0.85).astype(np.uint8)
    vox = np.stack([(rng.random(N_FN, dtype=np.float32) - np.float32(0.5)) * np.float32(0.16),
                    rng.random(N_FN, dtype=np.float32) * np.float32(5.0)], axis=1).astype(np.float32)
    vox[::5] = 0.0                                                      # unobserved voxels
    d['tsdf_voxels'] = vox
    d['tsdf_params'] = np.asarray([0.08, 5.0, 0.5], np.float32)         # truncation, max weight, invalid decay
    return d


def function_outputs_oracle(inp: dict, channels: int) -> dict:
    """The same probes through the CPU oracle's function-level entry points (orc_fn_*)."""
    from oracle import oracle as O
    L = O.lib()
    vp = C.c_void_p

    def p(a):
        return a.ctypes.data_as(vp)

    for name in ('orc_fn_interp_half', 'orc_fn_interp_float', 'orc_fn_interp_vertex'):
        getattr(L, name).restype = None
    L.orc_fn_interp_half.argtypes = [C.c_int, vp, vp, vp]
    L.orc_fn_interp_float.argtypes = [C.c_int, vp, vp, vp]
    L.orc_fn_blend.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp]
    L.orc_fn_interp_vertex.argtypes = [C.c_int, vp, vp, vp, vp]
    L.orc_fn_block_voxel.argtypes = [C.c_int, C.c_float, vp, vp]
    L.orc_fn_project.argtypes = [C.c_int, vp] + [C.c_float] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, vp]
    L.orc_fn_tsdf_functor.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, vp, vp, vp, vp]
    out = {}
    n = len(inp['interp_xy'])
    xy = np.ascontiguousarray(inp['interp_xy'], np.float32)
    fh = np.ascontiguousarray(inp['interp_half_f']).view(np.uint16)
    oh = np.zeros(n, np.uint16)
    L.orc_fn_interp_half(n, p(xy), p(fh), p(oh))
    out['interp_half_out'] = oh
    ff = np.ascontiguousarray(inp['interp_float_f'], np.float32)
    of = np.zeros(n, np.float32)
    L.orc_fn_interp_float(n, p(xy), p(ff), p(of))
    out['interp_float_out'] = of
    a = np.ascontiguousarray(inp['blend_a']).view(np.uint16)
    b = np.ascontiguousarray(inp['blend_b']).view(np.uint16)
    w = np.ascontiguousarray(inp['blend_w'], np.float32)
    ob = np.zeros_like(a)
    L.orc_fn_blend(a.shape[0], channels, p(a), p(b), p(w), p(ob))
    out['blend_out'] = ob
    v1, v2, sdf = (np.ascontiguousarray(inp[k], np.float32) for k in ('vertex_v1', 'vertex_v2', 'vertex_sdf'))
    ov = np.zeros_like(v1)
    L.orc_fn_interp_vertex(len(v1), p(v1), p(v2), p(sdf), p(ov))
    out['vertex_out'] = ov
    for k, bs in enumerate(inp['block_sizes']):
        pts = np.ascontiguousarray(inp[f'bv_points_{k}'], np.float32)
        o = np.zeros((len(pts), 6), np.int32)
        L.orc_fn_block_voxel(len(pts), C.c_float(bs), p(pts), p(o))
        out[f'bv_out_{k}'] = o
    for k in range(len(inp['project_poses'])):
        T = np.ascontiguousarray(inp['project_poses'][k], np.float32).reshape(16)
        fx, fy, cx, cy, W, H, bs, maxd = inp['project_cams'][k]
        bv = np.ascontiguousarray(inp[f'project_bv_{k}'], np.int32)
        o = np.zeros((len(bv), 6), np.float32)
        ok = np.zeros(len(bv), np.int32)
        L.orc_fn_project(len(bv), p(T), fx, fy, cx, cy, int(W), int(H), bs, maxd, p(bv), p(o), p(ok))
        out[f'project_out_{k}'] = o
        out[f'project_ok_{k}'] = ok
    trunc, maxw, inv = inp['tsdf_params']
    tin = np.ascontiguousarray(inp['tsdf_in'], np.float32)
    act = np.ascontiguousarray(inp['tsdf_active'], np.uint8)
    for mode in range(6):
        vox = np.ascontiguousarray(inp['tsdf_voxels'], np.float32).copy()
        upd = np.zeros(len(tin), np.uint8)
        L.orc_fn_tsdf_functor(len(tin), trunc, maxw, inv, mode, p(tin), p(act), p(vox), p(upd))
        out[f'tsdf_out_{mode}'] = vox
        out[f'tsdf_updated_{mode}'] = upd
    return out


# -------------------------------------------------------------------------------------------------------------
# frame sequences
# -------------------------------------------------------------------------------------------------------------
SCENARIOS = {
    # mindmap's cube-stacking settings (2 cm, workspace box, alpha = 1, raycast subsampling 1, decay 0.98)
    'cube_stacking': dict(voxel_size=0.02, workspace=S.WS_CUBE_STACKING, scene=S.S_TABLE, W=96, H=96, steps=5,
                          cameras=[dict(orbit=dict(stride=7))], alpha=1.0, max_dist=5.0, raycast_sub=1, decay=0.98,
                          weighting='kInverseSquareWeight', color=True, mask_every=0, invalid_decay=-1.0),
    # drill-in-box shape: 1 cm, two cameras per step (fixed head + orbiting wrist), exponential filter (alpha 0.8),
    # masks on odd steps
    'drill_in_box': dict(voxel_size=0.01, workspace=S.WS_DRILL_IN_BOX, scene=S.S_SPHERE_SMALL, W=48, H=48, steps=3,
                         cameras=[dict(fixed=((-0.05, 0.0, 0.45), (0.35, 0.0, 0.1))),
                                  dict(orbit=dict(stride=5, radius=0.4, height=0.4))],
                         alpha=0.8, max_dist=5.0, raycast_sub=1, decay=0.999, weighting='kInverseSquareWeight',
                         color=False, mask_every=2, invalid_decay=-1.0),
    # unbounded map, 5 cm voxels, raycast subsampling 4, depth beyond the integration distance, invalid-depth decay,
    # drop-off weighting, alpha 0.3, a camera that stands still for two frames (viewpoint-cache hit)
    'plane_unbounded': dict(voxel_size=0.05, workspace=None, scene=dict(plane_z=0.0, spheres=[(0.3, 0.1, 0.3, 0.3)]),
                            W=80, H=60, steps=4,
                            cameras=[dict(poses=[((0.0, -0.2, 1.6), (0.3, 0.1, 0.0)), ((0.0, -0.2, 1.6), (0.3, 0.1, 0.0)),
                                                 ((0.3, -0.5, 1.2), (0.3, 0.1, 0.0)), ((2.5, 2.0, 3.5), (0.3, 0.1, 0.0))])],
                            alpha=0.3, max_dist=3.0, raycast_sub=4, decay=0.9, weighting='kInverseSquareDropoffWeight',
                            color=True, mask_every=3, invalid_decay=0.5, holes=True),
    # stick-in-bin shape (MM/mapping/nvblox_mapper_constants.py:72-80): a workspace 4-5 m AWAY from the origin (block
    # indices 23..34: the float arithmetic of block / voxel indexing runs in another binade), 2 cm, orbit + fixed camera
    'stick_in_bin': dict(voxel_size=0.02, workspace=((3.7, 1.5, 0.44), (5.5, 3.2, 1.25)),
                         scene=dict(plane_z=0.5, spheres=[(4.6, 2.35, 0.6, 0.11)]), W=48, H=48, steps=3,
                         cameras=[dict(orbit=dict(stride=9, radius=0.42, height=1.02, center=(4.6, 2.35, 0.58))),
                                  dict(fixed=((4.05, 1.9, 1.15), (4.6, 2.35, 0.55)))],
                         alpha=1.0, max_dist=5.0, raycast_sub=1, decay=0.98, weighting='kInverseSquareWeight',
                         color=True, mask_every=0, invalid_decay=-1.0),
    # mug-in-drawer shape (nvblox_mapper_constants.py:45-53): tall workspace, decay 0.999, non-square frames, raycast
    # subsampling 2, constant weighting, masks every step
    'mug_in_drawer': dict(voxel_size=0.02, workspace=((-0.2, -0.8, -0.2), (0.9, 0.8, 1.0)), scene=S.S_TABLE, W=64, H=48,
                          steps=4, cameras=[dict(orbit=dict(stride=11, radius=0.5, height=0.62))],
                          alpha=1.0, max_dist=5.0, raycast_sub=2, decay=0.999, weighting='kConstantWeight',
                          color=False, mask_every=1, invalid_decay=-1.0),
}


def scenario_params(sc):
    """(nvblox_torch MapperParams, oracle NvbxParams) of a scenario."""
    from tests.parity_utils import make_params, oracle_params_from
    mp, _ = make_params(workspace=sc['workspace'], max_dist=sc['max_dist'], alpha=sc['alpha'],
                        raycast_sub=sc['raycast_sub'], decay=sc['decay'], strict=True, weighting=sc['weighting'])
    mp._projective_integrator_params.projective_tsdf_integrator_invalid_depth_decay_factor = sc['invalid_decay']
    return mp, oracle_params_from(mp)


def scenario_frames(sc, channels):
    """Yields (step, camera, T_W_C, K, depth, mask, features, rgb)."""
    W, H = sc['W'], sc['H']
    K = S.intrinsics(W, H)
    for step in range(sc['steps']):
        for ci, cam in enumerate(sc['cameras']):
            if 'orbit' in cam:
                o = dict(cam['orbit'])
                stride = o.pop('stride')
                T = S.orbit_pose((step * stride) % 64, **o)
            elif 'fixed' in cam:
                T = S.look_at(*cam['fixed'])
            else:
                T = S.look_at(*cam['poses'][step])
            depth = S.render_depth(K, H, W, T, **sc['scene'])
            if sc.get('holes'):
                depth = depth.copy()
                depth[5:15, 10:30] = 0.0        # invalid depth: exercises invalid_depth_decay_factor
            mask = None
            if sc['mask_every'] and step % sc['mask_every'] == sc['mask_every'] - 1:
                mask = S.border_lower_half_mask(H, W)
            seed = 5000 + 100 * step + ci
            feat = S.feature_frame(H, W, channels, seed)
            rgb = S.color_frame(H, W, seed + 50) if sc['color'] else None
            yield step, ci, T, K, depth, mask, feat, rgb


class OracleBackend:
    def __init__(self, sc, channels):
        from oracle import oracle as O
        _, op = scenario_params(sc)
        self.m = O.OracleMapper(sc['voxel_size'], channels, op)

    def depth(self, depth, T, K, mask):
        self.m.add_depth_frame(depth, T, K, mask)

    def features(self, feat, T, K, mask):
        self.m.add_feature_frame(feat, T, K, mask)

    def color(self, rgb, T, K, mask):
        self.m.add_color_frame(rgb, T, K, mask)

    def decay(self):
        self.m.decay()

    def last_block_list(self, which):
        return self.m.last_block_list(which)

    def synthetic_depth(self):
        return self.m.last_synthetic_depth()

    def tsdf(self):
        return self.m.all_blocks(0)

    def feat(self):
        return self.m.all_blocks(1)

    def colour(self):
        return self.m.all_color_blocks()


def _sorted(idx):
    idx = np.asarray(idx, np.int32).reshape(-1, 3)
    return idx[np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))] if len(idx) else idx


def run_scenario(sc, channels, be) -> dict:
    """Feed the scenario to a backend; returns per-frame stage outputs and the final layers as numpy arrays."""
    tl, fl, sd, tl_n, fl_n = [], [], [], [], []
    for step, ci, T, K, depth, mask, feat, rgb in scenario_frames(sc, channels):
        be.depth(depth, T, K, mask)
        a = _sorted(be.last_block_list(0))
        tl.append(a)
        tl_n.append(len(a))
        if rgb is not None:
            be.color(rgb, T, K, mask)
        be.features(feat, T, K, mask)
        b = _sorted(be.last_block_list(1))
        fl.append(b)
        fl_n.append(len(b))
        sd.append(np.asarray(be.synthetic_depth(), np.float32).reshape(-1) if len(b) else np.zeros(0, np.float32))
        if ci == len(sc['cameras']) - 1:
            be.decay()
    res = dict(frame_tsdf_lists=np.concatenate(tl) if tl else np.zeros((0, 3), np.int32),
               frame_tsdf_counts=np.asarray(tl_n, np.int32),
               frame_feat_lists=np.concatenate(fl) if fl else np.zeros((0, 3), np.int32),
               frame_feat_counts=np.asarray(fl_n, np.int32),
               frame_synth_depth=np.concatenate(sd) if sd else np.zeros(0, np.float32),
               frame_synth_sizes=np.asarray([len(x) for x in sd], np.int32))
    res['tsdf_idx'], res['tsdf_data'] = be.tsdf()
    res['feat_idx'], res['feat_data'] = be.feat()
    if sc['color']:
        res['color_idx'], res['color_rgb'], res['color_w'] = be.colour()
    return res


def oracle_mesh_rows(sc, channels, res) -> dict:
    """Un-welded marching-cubes vertices of the final TSDF layer and the voxel the closest-voxel paint picks for each,
    through the oracle: a scratch map receives the TSDF blocks and feature blocks whose channel 0 holds each voxel's
    linear index."""
    from oracle import oracle as O
    L = O.lib()
    L.orc_set_feature_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    _, op = scenario_params(sc)
    op.mesh_weld_vertices = 0
    m = O.OracleMapper(sc['voxel_size'], channels, op)
    idx, tsdf = res['tsdf_idx'], res['tsdf_data']
    fb = np.zeros((512, channels + 1), np.float16)
    fb[:, 0] = np.arange(512, dtype=np.float16)
    for i, b in enumerate(idx):
        m.set_tsdf_block(b, tsdf[i])
        L.orc_set_feature_block(m.h, int(b[0]), int(b[1]), int(b[2]), fb.ctypes.data_as(C.c_void_p))
    m.mark_all_dirty()
    m.update_feature_mesh()
    verts, feats, tris, vb = m.get_feature_mesh(with_block_index=True)
    return mesh_rows_from(idx, verts, feats[:, 0].astype(np.int64), vb)


def mesh_rows_from(idx, verts, voxel, vb) -> dict:
    """Group vertices by block (in the order of idx) and sort each block's rows canonically."""
    rows = np.concatenate([np.ascontiguousarray(verts, np.float32).view(np.uint32).astype(np.int64),
                           np.asarray(voxel, np.int64)[:, None]], axis=1) if len(verts) else np.zeros((0, 4), np.int64)
    out, cnt = [], []
    key = {tuple(b): i for i, b in enumerate(np.asarray(idx).tolist())}
    owner = np.asarray([key[tuple(b)] for b in np.asarray(vb).tolist()], np.int64) if len(verts) else np.zeros(0, np.int64)
    for i in range(len(idx)):
        r = rows[owner == i]
        r = r[np.lexsort(r.T[::-1])] if len(r) else r
        out.append(r)
        cnt.append(len(r))
    allr = np.concatenate(out) if out else np.zeros((0, 4), np.int64)
    return dict(mesh_vertex_bits=allr[:, :3].astype(np.uint32), mesh_vertex_voxel=allr[:, 3].astype(np.int16),
                mesh_counts=np.asarray(cnt, np.int32))


# -------------------------------------------------------------------------------------------------------------
# the CUDA product as a backend (tests/test_gpu_ref_vectors.py)
# -------------------------------------------------------------------------------------------------------------
class GpuBackend:
    def __init__(self, sc, channels, device=0):
        import torch
        from nvblox_torch.constants import constants
        from nvblox_torch.mapper import Mapper
        constants.set_feature_array_num_elements(channels)
        mp, _ = scenario_params(sc)
        self.t = torch
        self.dev = f'cuda:{device}'
        self.C = channels
        self.m = Mapper(voxel_sizes_m=float(sc['voxel_size']), mapper_parameters=mp)

    def _d(self, a):
        return None if a is None else self.t.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def depth(self, depth, T, K, mask):
        self.m.add_depth_frame(self._d(depth), self.t.from_numpy(T), self.t.from_numpy(K), self._d(mask))

    def features(self, feat, T, K, mask):
        self.m.add_feature_frame(self._d(feat), self.t.from_numpy(T), self.t.from_numpy(K), self._d(mask))

    def color(self, rgb, T, K, mask):
        self.m.add_color_frame(self._d(rgb), self.t.from_numpy(T), self.t.from_numpy(K), self._d(mask))

    def decay(self):
        self.m.decay()

    def last_block_list(self, which):
        from nvblox_mindmap_b200 import _capi
        L = _capi.load()
        n = int(_capi.check(L.nvbx_debug_last_block_list(self.m._handle, 0, which, None, 0, self.m._stream())))
        out = np.zeros((max(n, 1), 3), np.int32)
        _capi.check(L.nvbx_debug_last_block_list(self.m._handle, 0, which, out.ctypes.data_as(C.c_void_p), n,
                                                 self.m._stream()))
        return out[:n]

    def synthetic_depth(self):
        from nvblox_mindmap_b200 import _capi
        from nvblox_mindmap_b200.torch_interop import device_view
        L = _capi.load()
        p, r, c = C.c_void_p(), C.c_int(), C.c_int()
        _capi.check(L.nvbx_debug_last_synthetic_depth(self.m._handle, 0, C.byref(p), C.byref(r), C.byref(c)))
        return device_view(p.value, (r.value, c.value), self.t.float32, self.m._device, owner=self.m).cpu().numpy()

    def _layer(self, view, shape, dtype):
        from tests.parity_utils import gpu_blocks
        idx, data = gpu_blocks(view)
        if data is None:
            data = np.zeros((0,) + shape, dtype)
        return idx, data

    def tsdf(self):
        return self._layer(self.m.tsdf_layer_view(0), (8, 8, 8, 2), np.float32)

    def feat(self):
        return self._layer(self.m.feature_layer_view(0), (8, 8, 8, self.C + 1), np.float16)

    def colour(self):
        from nvblox_mindmap_b200.torch_interop import device_view
        layer = self.m.color_layer_view(0)
        idx, rgb = self._layer(layer, (8, 8, 8, 3), np.uint8)
        w = np.zeros((len(idx), 8, 8, 8), np.float32)
        for k, row in enumerate(idx):
            blk = layer.get_block_at_index(self.t.from_numpy(row))
            raw = device_view(blk.data_ptr(), (8, 8, 8, 8), self.t.uint8, self.m._device, owner=self.m).cpu().numpy()
            w[k] = raw[..., 4:8].copy().view(np.float32)[..., 0]
        return idx, rgb, w


def gpu_mesh_rows(sc, channels, tsdf_idx, tsdf_data, device=0) -> np.ndarray:
    """Un-welded marching-cubes vertices + painted voxel id of the given TSDF layer through the CUDA product:
    globally sorted rows (x bits, y bits, z bits, voxel)."""
    import torch
    from nvblox_mindmap_b200 import _capi
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    constants.set_feature_array_num_elements(channels)
    mp, _ = scenario_params(sc)
    mp._mesh_integrator_params.mesh_integrator_weld_vertices = False
    m = Mapper(voxel_sizes_m=float(sc['voxel_size']), mapper_parameters=mp)
    tl, fl = m.tsdf_layer_view(0), m.feature_layer_view(0)
    fb = torch.zeros((8, 8, 8, channels + 1), dtype=torch.float16, device=f'cuda:{device}')
    fb[..., 0] = torch.arange(512, device=f'cuda:{device}').reshape(8, 8, 8).to(torch.float16)
    for i, b in enumerate(tsdf_idx):
        bt = torch.from_numpy(np.ascontiguousarray(b))
        tl.allocate_block_at_index(bt)
        fl.allocate_block_at_index(bt)
    for i, b in enumerate(tsdf_idx):    # views are taken after every allocation (allocation may move nothing, but
        bt = torch.from_numpy(np.ascontiguousarray(b))    # the documented contract is "invalidated by any mutation")
        tl.get_block_at_index(bt).copy_(torch.from_numpy(tsdf_data[i]).to(f'cuda:{device}'))
        fl.get_block_at_index(bt).copy_(fb)
    _capi.check(_capi.load().nvbx_mark_all_dirty(m._handle, 0, m._stream()))
    m.update_feature_mesh(0)
    mesh = m.get_feature_mesh(0)
    v = mesh.vertices().cpu().numpy()
    vox = mesh.vertex_features()[:, 0].cpu().numpy().astype(np.int64)
    rows = np.concatenate([np.ascontiguousarray(v, np.float32).view(np.uint32).astype(np.int64), vox[:, None]], axis=1)
    return rows[np.lexsort(rows.T[::-1])] if len(rows) else rows
