"""Deterministic golden scenarios of the hot path and the digests that summarise their outputs.

A scenario feeds a fixed frame sequence to a *driver* (the CPU oracle or the CUDA Mapper, see
OracleDriver / GpuDriver) and records, after every step, SHA-256 digests of the products the parity
contract names: the TSDF view block list, the TSDF layer, the synthetic depth image, the feature band
block list, the feature layer and the (canonicalised) feature mesh.  tests/golden/core_path.json holds
the digests produced by the oracle when tests/golden/make_golden.py was run; both the oracle (CPU
suite) and the CUDA path (-m gpu) must reproduce them bit for bit.
"""
import hashlib
import os

import numpy as np

from tests import scenes as S
from tests.parity_utils import canonical_mesh, make_params, sort_rows

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()[:32]


class OracleDriver:
    def __init__(self, voxel, C, mp, op):
        from oracle import oracle as O
        self.m = O.OracleMapper(voxel, C, op)

    def depth(self, d, T, K, mask=None):
        self.m.add_depth_frame(d, T, K, mask)

    def features(self, f, T, K, mask=None):
        self.m.add_feature_frame(f, T, K, mask)

    def decay(self):
        self.m.decay()

    def view_blocks(self):
        return sort_rows(self.m.last_block_list(0))

    def band_blocks(self):
        return sort_rows(self.m.last_block_list(1))

    def synth(self):
        return self.m.last_synthetic_depth()

    def layer(self, which):
        return self.m.all_blocks(which)

    def mesh(self):
        self.m.update_feature_mesh()
        return self.m.get_feature_mesh()


class GpuDriver:
    def __init__(self, voxel, C, mp, op):
        from tests.parity_utils import Pair
        self.p = Pair(voxel, C, mp, op)      # the oracle half of the pair stays idle

    def depth(self, d, T, K, mask=None):
        t = self.p.torch
        self.p.gpu.add_depth_frame(t.from_numpy(d).to(self.p.dev), t.from_numpy(T), t.from_numpy(K),
                                   None if mask is None else t.from_numpy(mask).to(self.p.dev))

    def features(self, f, T, K, mask=None):
        t = self.p.torch
        self.p.gpu.add_feature_frame(t.from_numpy(f).to(self.p.dev), t.from_numpy(T), t.from_numpy(K),
                                     None if mask is None else t.from_numpy(mask).to(self.p.dev))

    def decay(self):
        self.p.gpu.decay()

    def _list(self, which):
        import ctypes as C
        from nvblox_mindmap_b200 import _capi
        L, g = _capi.load(), self.p.gpu
        n = int(_capi.check(L.nvbx_debug_last_block_list(g._handle, 0, which, None, 0, g._stream())))
        out = np.zeros((max(n, 1), 3), np.int32)
        _capi.check(L.nvbx_debug_last_block_list(g._handle, 0, which, out.ctypes.data_as(C.c_void_p), n, g._stream()))
        return sort_rows(out[:n])

    def view_blocks(self):
        return self._list(0)

    def band_blocks(self):
        return self._list(1)

    def synth(self):
        return self.p.synthetic_depth()[0]

    def layer(self, which):
        from tests.parity_utils import gpu_blocks
        g = self.p.gpu
        return gpu_blocks(g.tsdf_layer_view(0) if which == 0 else g.feature_layer_view(0))

    def mesh(self):
        g = self.p.gpu
        g.update_feature_mesh(0)
        m = g.get_feature_mesh(0)
        return m.vertices().cpu().numpy(), m.vertex_features().cpu().numpy(), m.triangles().cpu().numpy()


def _record(drv, with_mesh):
    rec = {}
    vb = drv.view_blocks()
    rec['view_blocks'] = [len(vb), _digest(vb)]
    bb = drv.band_blocks()
    rec['band_blocks'] = [len(bb), _digest(bb)]
    rec['synth'] = _digest(drv.synth().view(np.uint32))
    ti, td = drv.layer(0)
    rec['tsdf'] = [len(ti), _digest(ti, td.view(np.uint32)) if len(ti) else '']
    fi, fd = drv.layer(1)
    rec['features'] = [len(fi), _digest(fi, fd.view(np.uint16)) if len(fi) else '']
    if with_mesh:
        v, f, t = drv.mesh()
        rows, tri = canonical_mesh(v, f, t)
        rec['mesh'] = [int(len(v)), int(len(t)), _digest(rows, tri)]
    return rec


def load_threedmatch():
    """3 frames of nvblox's own 3DMatch test sequence (NB/tests/data/3dmatch/seq-01, frames 0-2), depth
    nearest-downsampled 4x to 160x120 by make_golden.py; poses and intrinsics as in the fixture."""
    z = np.load(os.path.join(GOLDEN_DIR, 'threedmatch_seq01_160x120.npz'))
    return z['depth_mm'].astype(np.float32) / 1000.0, z['poses'].astype(np.float32), z['K'].astype(np.float32)


def scenario_table_orbit(drv_cls):
    """Cube-stacking style replay (mindmap parameters, strict blend so features are bit-exact):
    4 frames, decay every step, lower-half mask on frame 2, mesh after frames 1 and 3."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, strict=True)
    drv = drv_cls(0.02, 32, mp, op)
    out = []
    K = S.intrinsics(96, 96)
    for i in range(4):
        T = S.orbit_pose(i)
        if i:
            drv.decay()
        drv.depth(S.render_depth(K, 96, 96, T, **S.S_TABLE), T, K)
        drv.features(S.feature_frame(96, 96, 32, 2000 + i), T, K,
                     S.border_lower_half_mask(96, 96) if i == 2 else None)
        out.append(_record(drv, with_mesh=i % 2 == 1))
    return out


def scenario_blend_alpha03(drv_cls):
    """Exponential filter with alpha = 0.3 (the reference default is 0.8; test_feature_integrator uses 0.3)."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=0.3)
    drv = drv_cls(0.02, 16, mp, op)
    out = []
    K = S.intrinsics(80, 64)
    for i in range(3):
        T = S.orbit_pose(2 * i, 64, 0.5, 0.45)
        drv.depth(S.render_depth(K, 64, 80, T, **S.S_SPHERE_SMALL), T, K)
        drv.features(S.feature_frame(64, 80, 16, 3000 + i), T, K)
        out.append(_record(drv, with_mesh=i == 2))
    return out


def scenario_threedmatch(drv_cls):
    """Real depth: nvblox's 3DMatch fixture, 5 cm voxels, reference default parameters (no workspace
    box -> overflow-hash index, raycast subsampling 4, 7 m range, alpha 0.8)."""
    mp, op = make_params(workspace=None, max_dist=7.0, raycast_sub=4, alpha=0.8, decay=0.95)
    drv = drv_cls(0.05, 16, mp, op)
    depth, poses, K = load_threedmatch()
    out = []
    for i in range(len(depth)):
        drv.depth(depth[i], poses[i], K)
        drv.features(S.feature_frame(depth.shape[1], depth.shape[2], 16, 4000 + i), poses[i], K)
        out.append(_record(drv, with_mesh=i == len(depth) - 1))
    return out


def scenario_two_cameras_1cm_c1024(drv_cls):
    """Drill-in-box / stress shape: head (static: viewpoint-cache hit on step 1) + wrist camera per step, 1 cm
    voxels, 1024 channels (the 4 x 512-byte chunk path), drill-in-box workspace, decay 0.999."""
    mp, op = make_params(workspace=S.WS_DRILL_IN_BOX, decay=0.999, strict=True)
    drv = drv_cls(0.01, 1024, mp, op)
    out = []
    K = S.intrinsics(80, 80)
    T_head = S.look_at((-0.2, 0.0, 0.6), (0.4, 0.0, 0.05))
    for i in range(2):
        if i:
            drv.decay()
        for cam, T in enumerate((T_head, S.orbit_pose(i, 64, 0.4, 0.45))):
            drv.depth(S.render_depth(K, 80, 80, T, **S.S_TABLE), T, K)
            drv.features(S.feature_frame(80, 80, 1024, 5000 + 2 * i + cam), T, K)
            out.append(_record(drv, with_mesh=(i == 1 and cam == 1)))
    return out


SCENARIOS = {
    'table_orbit': scenario_table_orbit,
    'blend_alpha03': scenario_blend_alpha03,
    'threedmatch': scenario_threedmatch,
    'two_cameras_1cm_c1024': scenario_two_cameras_1cm_c1024,
}
