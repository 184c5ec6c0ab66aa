"""GPU parity tests: the CUDA path (through the nvblox_torch surface -> C ABI) against the CPU oracle on
identical seeded inputs.  Bars (BASELINE.json north_star): allocated block sets and extracted voxel
indices bit-exact; TSDF distance/weight within 1e-5 relative (we assert bit-exact, which is stronger and
holds because both sides use the uncontracted IEEE operation sequence); features within 1 fp16 ulp (we
assert bit-exact for the default non-fused arithmetic)."""
import numpy as np
import pytest

from tests import scenes as S
from tests.parity_utils import Pair, make_params, orbit_frames

pytestmark = pytest.mark.gpu


def test_single_frame_plumbing():
    """BASELINE configs[0] scaled down: one synthetic depth + feature frame into a 2 cm map."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 64, mp, op)
    K = S.intrinsics(128, 128)
    T = S.look_at((0.35, 0.0, 0.9), (0.35, 0.0, 0.0), up=(1, 0, 0))
    depth = S.render_depth(K, 128, 128, T, **S.S_TABLE)
    pair.depth(depth, T, K)
    g, c = pair.last_block_list(0)
    assert np.array_equal(g, c) and len(g) > 0
    assert pair.check_tsdf() > 0
    feat = S.feature_frame(128, 128, 64, 7)
    pair.features(feat, T, K)
    g, c = pair.last_block_list(1)
    assert np.array_equal(g, c) and len(g) > 0
    gs, cs = pair.synthetic_depth()
    assert np.array_equal(gs.view(np.uint32), cs.view(np.uint32))
    assert pair.check_features() > 0
    assert pair.check_mesh() > 0
    gc, cc = pair.gpu.counters(0), pair.cpu.counters()
    for k in ('tsdf_voxels_updated', 'feature_voxels_updated', 'feature_band_blocks', 'feature_candidate_blocks',
              'tsdf_blocks_allocated', 'feature_blocks_allocated', 'mesh_vertices'):
        assert gc[k] == cc[k], (k, gc[k], cc[k])


@pytest.mark.parametrize('alpha,strict', [(1.0, False), (1.0, True), (0.3, False)])
def test_orbit_sequence(alpha, strict):
    """Cube-stacking style replay: moving wrist camera, decay every step, mesh every second step."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=alpha, strict=strict)
    pair = Pair(0.02, 32, mp, op)
    for i, T, K, depth, feat in orbit_frames(6, 96, 96, 32, S.S_TABLE):
        if i:
            pair.decay()
        pair.depth(depth, T, K)
        pair.features(feat, T, K, mask=S.border_lower_half_mask(96, 96) if i % 3 == 2 else None)
        pair.check_tsdf()
        # alpha == 1 fast path differs from the reference arithmetic only in the sign of zero
        pair.check_features(max_ulp=0 if (strict or alpha != 1.0) else 1)
        if i % 2:
            pair.check_mesh()
    assert pair.check_mesh() > 0


def test_c768_features_bit_exact():
    """The benchmark feature length (vectorised 3-vectors-per-lane path)."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 768, mp, op)
    for i, T, K, depth, feat in orbit_frames(2, 64, 64, 768, S.S_SPHERE_SMALL):
        pair.depth(depth, T, K)
        pair.features(feat, T, K)
    pair.check_tsdf()
    assert pair.check_features(max_ulp=1) > 0
    assert pair.check_mesh() > 0


@pytest.mark.parametrize('variant,permille,ticket', [(0, 0, 1), (1, 300, 2), (4, 300, 2), (4, 1000, 1), (4, 0, 3), (5, 500, 4),
                                                    (6, 150, 2), (7, 0, 1), (8, 0, 1), (7, -1, 1), (4, -1, 1), (10, 0, 1)])
@pytest.mark.parametrize('C,alpha', [(768, 1.0), (40, 0.3)])
def test_gather_schedules_are_result_identical(variant, permille, ticket, C, alpha):
    """Every schedule of the feature gather (static deal / ticketed tail, nvbx_set_gather_tuning) writes the
    same bytes: strict-mode features stay bit-exact against the oracle, blended (alpha < 1) frames included."""
    from nvblox_mindmap_b200 import _capi
    lib = _capi.load()
    assert lib.nvbx_set_gather_tuning(variant, permille, ticket) == 0
    try:
        mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=alpha, strict=True)
        pair = Pair(0.02, C, mp, op)
        for i, T, K, depth, feat in orbit_frames(3, 80, 80, C, S.S_TABLE):
            pair.depth(depth, T, K)
            pair.features(feat, T, K)
            assert pair.check_features(max_ulp=0) > 0
        g, c = pair.gpu.counters(0), pair.cpu.counters()
        assert g['feature_voxels_updated'] == c['feature_voxels_updated'] > 0
    finally:
        assert lib.nvbx_set_gather_tuning(7, 0, 4) == 0
    assert lib.nvbx_set_gather_tuning(11, 0, 1) != 0 and lib.nvbx_set_gather_tuning(4, 1001, 1) != 0 and lib.nvbx_set_gather_tuning(4, -2, 1) != 0


def test_decay_until_removed():
    """test_tsdf_decay.cpp DecayUntilRemoved: repeated decay frees every block (and its feature block)."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, decay=0.5)
    pair = Pair(0.02, 16, mp, op)
    for i, T, K, depth, feat in orbit_frames(1, 64, 64, 16, S.S_TABLE):
        pair.depth(depth, T, K)
        pair.features(feat, T, K)
    n0 = pair.check_tsdf()
    assert n0 > 0
    for _ in range(16):
        pair.decay()
        pair.check_tsdf()
        pair.check_features(max_ulp=1)
    assert pair.gpu.tsdf_layer_view(0).num_blocks() == 0 == pair.cpu.num_blocks(0)
    assert pair.gpu.feature_layer_view(0).num_blocks() == 0
    # the map is usable again after everything was released (slots recycled, hash rebuilt)
    for i, T, K, depth, feat in orbit_frames(2, 64, 64, 16, S.S_TABLE, seed0=50):
        pair.depth(depth, T, K)
        pair.features(feat, T, K)
    pair.check_tsdf()
    pair.check_features(max_ulp=1)
    pair.check_mesh()


def test_viewpoint_cache_static_camera():
    """Q6: a static camera re-uses the previous block list even though the depth image changed."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 16, mp, op)
    K = S.intrinsics(96, 96)
    T = S.look_at((-0.2, 0.0, 0.6), (0.4, 0.0, 0.0))
    d1 = S.render_depth(K, 96, 96, T, **S.S_TABLE)
    d2 = S.render_depth(K, 96, 96, T, **S.S_SPHERE_SMALL)
    pair.depth(d1, T, K)
    l1 = pair.last_block_list(0)
    pair.depth(d2, T, K)
    l2 = pair.last_block_list(0)
    assert np.array_equal(l1[0], l2[0]) and np.array_equal(l2[0], l2[1])
    pair.features(S.feature_frame(96, 96, 16, 3), T, K)
    pair.check_tsdf()
    pair.check_features(max_ulp=1)
    pair.clear()
    pair.depth(d2, T, K)    # cache survives clear(): blocks are re-allocated from the cached list
    pair.check_tsdf()
    pair.check_mesh()


def test_unbounded_workspace_and_subsampling():
    """Reference defaults: no workspace box, raycast subsampling 4, 7 m range (arena growth path)."""
    mp, op = make_params(workspace=None, max_dist=7.0, raycast_sub=4, alpha=0.8)
    pair = Pair(0.05, 16, mp, op)
    for i, T, K, depth, feat in orbit_frames(3, 96, 128, 16, S.S_TABLE, radius=1.5, height=1.2):
        pair.depth(depth, T, K)
        pair.features(feat, T, K)
    assert pair.check_tsdf() > 0
    assert pair.check_features() > 0
    assert pair.check_mesh() > 0


def test_full_size_cube_stacking_frames():
    """BASELINE configs[1] at FULL size (512x512, C=768, 2 cm, mindmap parameters): two frames of the bench
    workload, every product bit-exact against the oracle (features within 1 ulp: alpha == 1 fast path)."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 768, mp, op)
    K = S.intrinsics(512, 512)
    for i in range(2):
        T = S.orbit_pose(i)
        depth = S.render_depth(K, 512, 512, T, **S.S_TABLE)
        pair.depth(depth, T, K)
        g, c = pair.last_block_list(0)
        assert np.array_equal(g, c) and len(g) > 300
        pair.features(S.feature_frame(512, 512, 768, 1000 + i), T, K)
        g, c = pair.last_block_list(1)
        assert np.array_equal(g, c) and len(g) > 80
        gs, cs = pair.synthetic_depth()
        assert np.array_equal(gs.view(np.uint32), cs.view(np.uint32))
    assert pair.check_tsdf() > 300
    assert pair.check_features(max_ulp=1) > 80
    assert pair.check_mesh() > 1000
    gc, cc = pair.gpu.counters(0), pair.cpu.counters()
    assert gc['feature_voxels_updated'] == cc['feature_voxels_updated'] > 30000


def test_drill_in_box_two_cameras_1cm():
    """BASELINE configs[2] scaled to 160x160: head (static -> viewpoint cache hits) + wrist camera, 1 cm voxels,
    drill-in-box workspace (3 740-cell index grid), decay 0.999, mesh each step."""
    mp, op = make_params(workspace=S.WS_DRILL_IN_BOX, decay=0.999)
    pair = Pair(0.01, 64, mp, op)
    K = S.intrinsics(160, 160)
    T_head = S.look_at((-0.2, 0.0, 0.6), (0.4, 0.0, 0.05))
    for i in range(3):
        if i:
            pair.decay()
        for cam, T in (('head', T_head), ('wrist', S.orbit_pose(i, 64, 0.4, 0.45))):
            depth = S.render_depth(K, 160, 160, T, **S.S_TABLE)
            pair.depth(depth, T, K)
            g, c = pair.last_block_list(0)
            assert np.array_equal(g, c), cam
            pair.features(S.feature_frame(160, 160, 64, 500 + 2 * i + (cam == 'wrist')), T, K)
            g, c = pair.last_block_list(1)
            assert np.array_equal(g, c), cam
        pair.check_tsdf()
        pair.check_features(max_ulp=1)
        assert pair.check_mesh() > 0


def test_two_mappers_static_dynamic():
    """mindmap creates two maps (static / dynamic, nvblox_mapping_helpers.py:72-76): they must not interact."""
    import torch
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    from oracle import oracle as O
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    constants.set_feature_array_num_elements(16)
    gpu = Mapper(voxel_sizes_m=[0.02, 0.04], mapper_parameters=mp)
    cpu = [O.OracleMapper(0.02, 16, op), O.OracleMapper(0.04, 16, op)]
    assert gpu.num_mappers() == 2
    K = S.intrinsics(96, 96)
    for i in range(3):
        T = S.orbit_pose(i)
        depth = S.render_depth(K, 96, 96, T, **S.S_TABLE)
        feat = S.feature_frame(96, 96, 16, 700 + i)
        mid = i % 2
        gpu.add_depth_frame(torch.from_numpy(depth).cuda(), torch.from_numpy(T), torch.from_numpy(K), None, mid)
        gpu.add_feature_frame(torch.from_numpy(feat).cuda(), torch.from_numpy(T), torch.from_numpy(K), None, mid)
        cpu[mid].add_depth_frame(depth, T, K)
        cpu[mid].add_feature_frame(feat, T, K)
    gpu.decay()          # mapper_id = -1: all maps
    for c in cpu:
        c.decay()
    from tests.parity_utils import gpu_blocks
    for mid in range(2):
        gi, gd = gpu_blocks(gpu.tsdf_layer_view(mid))
        ci, cd = cpu[mid].all_blocks(0)
        assert np.array_equal(gi, ci) and len(gi) > 0
        assert np.array_equal(gd.view(np.uint32), cd.view(np.uint32))
        gi, gd = gpu_blocks(gpu.feature_layer_view(mid))
        ci, cd = cpu[mid].all_blocks(1)
        assert np.array_equal(gi, ci) and len(gi) > 0
        assert np.array_equal(gd.view(np.uint16)[..., -1], cd.view(np.uint16)[..., -1])


def test_large_view_global_bitmap_and_hash_index():
    """1 cm voxels, no workspace box, 4 m range: the view AABB has > 262 144 cells, so rays mark the GLOBAL
    bitmap (k_raycast_mark<false>), compaction runs as its own kernel (kViewFromSlots), every block lives in the
    overflow hash and the arenas grow while frames arrive."""
    mp, op = make_params(workspace=None, max_dist=4.0, raycast_sub=2, alpha=0.8)
    pair = Pair(0.01, 16, mp, op)
    for i, T, K, depth, feat in orbit_frames(2, 64, 64, 16, S.S_TABLE, radius=0.8, height=0.7):
        pair.depth(depth, T, K)
        g, c = pair.last_block_list(0)
        assert np.array_equal(g, c) and len(g) > 100
        pair.features(feat, T, K)
        g, c = pair.last_block_list(1)
        assert np.array_equal(g, c)
    assert pair.check_tsdf() > 100
    assert pair.check_features() > 0
    pair.decay()
    assert pair.check_mesh() > 0


# ---- colour layer + colour mesh ------------------------------------------------------
@pytest.mark.parametrize('alpha', [1.0, 0.3])
def test_color_layer_and_mesh(alpha):
    """mindmap's integrate_frame order (depth -> colour -> features, nvblox_mapping_helpers.py:207-261) on an
    orbit with decay: colour blocks / bytes / weights, colour mesh and feature mesh all bit-exact; the feature
    frame re-uses the colour frame's synthetic depth (same pose, same TSDF state)."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=alpha)
    pair = Pair(0.02, 32, mp, op)
    for i, T, K, depth, feat in orbit_frames(6, 96, 96, 32, S.S_TABLE):
        if i:
            pair.decay()
        pair.depth(depth, T, K)
        mask = S.border_lower_half_mask(96, 96) if i % 3 == 2 else None
        pair.color(S.color_frame(96, 96, 2000 + i), T, K, mask=mask)
        g, c = pair.last_block_list(2)
        assert np.array_equal(g, c) and len(g) > 0
        pair.features(feat, T, K, mask=mask)
        gs, cs = pair.synthetic_depth()
        assert np.array_equal(gs.view(np.uint32), cs.view(np.uint32))
        assert pair.check_color() > 0
        pair.check_features(max_ulp=0 if alpha != 1.0 else 1)
        if i % 2:
            assert pair.check_color_mesh() > 0
            pair.check_mesh()
    gc, cc = pair.gpu.counters(0), pair.cpu.counters()
    for k in ('color_frames', 'color_band_blocks', 'color_voxels_updated', 'color_blocks_allocated',
              'feature_voxels_updated'):
        assert gc[k] == cc[k] and gc[k] > 0, (k, gc[k], cc[k])


def test_color_static_camera_and_features_first():
    """Static camera (every viewpoint cache hits), features BEFORE colour (the colour frame re-uses the synthetic
    depth), a second colour frame after a TSDF change (must re-trace), decay until the colour blocks are freed."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=0.5, decay=0.5)
    pair = Pair(0.02, 16, mp, op)
    K = S.intrinsics(64, 64)
    T = S.orbit_pose(3)
    depth = S.render_depth(K, 64, 64, T, **S.S_TABLE)
    for i in range(3):
        pair.depth(depth, T, K)
        pair.features(S.feature_frame(64, 64, 16, 50 + i), T, K)
        pair.color(S.color_frame(64, 64, 60 + i), T, K)
        assert pair.check_color() > 0
        pair.check_features()
        pair.check_color_mesh()
    for _ in range(16):
        pair.decay()
    assert pair.check_tsdf() == 0
    assert pair.check_color() == 0
    assert pair.check_color_mesh() == 0
    assert pair.gpu.color_layer_view(0).num_blocks() == 0


def test_color_mesh_without_color_frames():
    """A colour mesh of a map that never saw a colour frame is Gray (AppearanceGetter<ColorVoxel>::
    getDefaultAppearance, mesh_integrator_appearance.cu:50-53); the two mesh layers update independently."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 16, mp, op)
    for i, T, K, depth, feat in orbit_frames(2, 64, 64, 16, S.S_SPHERE_SMALL):
        pair.depth(depth, T, K)
        pair.features(feat, T, K)
    n = pair.check_color_mesh()
    assert n > 0
    cols = pair.gpu.get_color_mesh(0).vertex_colors()
    assert bool((cols == 127).all())
    assert pair.check_mesh() == n
    # colour frame afterwards: only the colour mesh's dirty set makes the colour mesh pick the colours up
    T, K = S.orbit_pose(1), S.intrinsics(64, 64)
    pair.color(S.color_frame(64, 64, 5), T, K)
    assert pair.check_color_mesh() == n
    assert not bool((pair.gpu.get_color_mesh(0).vertex_colors() == 127).all())


def test_color_and_feature_share_the_planes_viewpoint_cache():
    """shareViewpointCaches (mapper.cpp:56-58): the colour and the feature integrator of a Mapper use ONE planes
    cache, so a feature frame 0.5 mm away from the preceding colour frame re-uses the colour frame's block list
    (and vice versa).  Large camera steps in between make sure a stale list would be visible."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=0.5)
    pair = Pair(0.02, 16, mp, op)
    K = S.intrinsics(64, 64)
    for i in range(4):
        T = S.orbit_pose(5 * i, 64)
        depth = S.render_depth(K, 64, 64, T, **S.S_TABLE)
        pair.depth(depth, T, K)
        T2 = T.copy()
        T2[0, 3] += np.float32(0.0005)                       # within the 1 mm / 0.1 deg cache tolerance
        if i % 2:
            pair.color(S.color_frame(64, 64, 70 + i), T, K)
            pair.features(S.feature_frame(64, 64, 16, 80 + i), T2, K)
        else:
            pair.features(S.feature_frame(64, 64, 16, 80 + i), T, K)
            pair.color(S.color_frame(64, 64, 70 + i), T2, K)
        for which in (1, 2):
            g, c = pair.last_block_list(which)
            assert np.array_equal(g, c) and len(g) > 0
        assert pair.check_color() > 0
        assert pair.check_features() > 0


@pytest.mark.parametrize('bounds,weighting', [('kHeightBounds', 'kConstantWeight'),
                                              ('kBoundingBox', 'kInverseSquareDropoffWeight'),
                                              ('kUnbounded', 'kConstantDropoffWeight')])
def test_workspace_bound_types_and_weighting_modes(bounds, weighting):
    """The reference's other WorkspaceBoundsType / WeightingFunctionType values (test_workspace_bounds.cpp,
    test_weighting_function.cpp): same block sets and bit-exact TSDF / features as the oracle."""
    from tests.parity_utils import oracle_params_from
    mp, _ = make_params(workspace=((-0.3, -0.5, 0.02), (0.9, 0.5, 0.3)), weighting=weighting, max_dist=3.0, strict=True)
    vc = mp._view_calculator_params
    vc.workspace_bounds_type = bounds
    mp._projective_integrator_params.projective_integrator_max_weight = 100.0
    pair = Pair(0.04, 16, mp, oracle_params_from(mp))
    for i, T, K, depth, feat in orbit_frames(3, 96, 128, 16, S.S_TABLE, radius=0.8, height=0.7):
        pair.depth(depth, T, K)
        g, c = pair.last_block_list(0)
        assert np.array_equal(g, c) and len(g) > 0
        pair.features(feat, T, K)
    assert pair.check_tsdf() > 0
    assert pair.check_features(max_ulp=0) > 0
    assert pair.check_mesh() > 0


def test_empty_and_degenerate_inputs():
    """Edge cases the reference's tests touch (test_mesh.cpp BlankMap, test_mesh.py empty shapes, masks that hide
    everything, invalid depth): nothing is allocated, nothing crashes, shapes are empty, and the map still works
    afterwards.  Each step is mirrored on the oracle."""
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 24, mp, op)
    H, W = 72, 104                                   # non-square, not a multiple of 16
    K = S.intrinsics(W, H)
    T = S.orbit_pose(3)
    feat = S.feature_frame(H, W, 24, 11)
    # blank map: features / decay / mesh / export on nothing
    pair.features(feat, T, K)
    pair.decay()
    assert pair.check_mesh() == 0
    mesh = pair.gpu.get_feature_mesh(0)
    assert tuple(mesh.vertices().shape) == (0, 3) and tuple(mesh.vertex_features().shape) == (0, 24)
    assert tuple(mesh.triangles().shape) == (0, 3)
    assert pair.gpu.tsdf_layer_view(0).num_blocks() == 0 and pair.gpu.feature_layer_view(0).num_blocks() == 0
    # invalid depth everywhere (0 = no return; NaN) and a mask that hides a valid frame
    depth = S.render_depth(K, H, W, T, **S.S_TABLE)
    pair.depth(np.zeros((H, W), np.float32), T, K)
    pair.depth(np.full((H, W), np.nan, np.float32), T, K)
    assert pair.check_tsdf() == 0
    pair.depth(depth, T, K, mask=np.zeros((H, W), np.uint8))
    pair.check_tsdf()
    pair.features(feat, T, K, mask=np.zeros((H, W), np.uint8))
    pair.check_features(max_ulp=1)
    assert pair.gpu.counters(0)['feature_voxels_updated'] == pair.cpu.counters()['feature_voxels_updated'] == 0
    # ... and the same map keeps working.  (From the SAME pose it would not: the viewpoint cache holds the empty
    # block list of the zero-depth frame for that pose -- view_calculator.cu:256-265 -- on both sides.)
    assert pair.check_tsdf() == 0
    T = S.orbit_pose(9)
    depth = S.render_depth(K, H, W, T, **S.S_TABLE)
    pair.depth(depth, T, K)
    pair.features(feat, T, K)
    assert pair.check_tsdf() > 0 and pair.check_features(max_ulp=1) > 0 and pair.check_mesh() > 0
    pair.clear()
    assert pair.check_tsdf() == 0 and pair.check_mesh() == 0


@pytest.mark.parametrize('async_enqueue', [False, True])
@pytest.mark.parametrize('alpha', [1.0, 0.8])
def test_pipelined_frames_are_bit_identical(alpha, async_enqueue):
    """Frame pipelining (Mapper.set_pipelining: the gather of frame i on the map's own stream, the depth path of
    frame i + 1 underneath it) changes nothing in the map: TSDF, features (alpha < 1 blends with what the previous
    frames' gathers wrote), per-frame band lists, mesh, with decay / queries / block views joining in between.
    async_enqueue: the frames are queued and issued by the mapper's worker thread (nvbx_set_pipelining(m, 2))."""
    import torch
    from nvblox_torch.mapper import QueryType
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=alpha, strict=True)
    pair = Pair(0.02, 768, mp, op)
    pair.gpu.set_pipelining(True, async_enqueue=async_enqueue)
    H = W = 128
    K = S.intrinsics(W, H)
    frames = [S.feature_frame(H, W, 768, 900 + i) for i in range(3)]
    for i in range(9):
        T = S.orbit_pose(3 * i)
        pair.depth(S.render_depth(K, H, W, T, **S.S_TABLE), T, K)
        pair.features(frames[i % 3], T, K)
        g, c = pair.last_block_list(1)
        assert np.array_equal(g, c)
        if i % 4 == 3:
            pair.decay()                      # joins the gathers in flight before freeing blocks
        if i == 5:                            # a feature query in the middle of the stream of frames
            idx, _ = pair.cpu.all_blocks(1)
            q = ((idx[:64].astype(np.float32) + 0.5) * np.float32(0.16)).astype(np.float32)
            got = pair.gpu.query_layer(QueryType.FEATURE, torch.from_numpy(q).cuda(), mapper_id=0).cpu().numpy()
            assert np.array_equal(got.view(np.uint16), pair.cpu.query_features(q).view(np.uint16))
    assert pair.check_tsdf() > 0
    assert pair.check_features(max_ulp=0) > 0
    assert pair.check_mesh() > 0
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    waits, wait_ns = C.c_int64(-1), C.c_int64(-1)
    _capi.load().nvbx_pipeline_wait_stats(C.byref(waits), C.byref(wait_ns))
    assert waits.value >= 0 and wait_ns.value >= 0     # (how often the host had to wait for a ring slot: tuning aid)
    # switching it off drains the queue and the map's streams; further frames keep matching
    pair.gpu.set_pipelining(False)
    T = S.orbit_pose(40)
    pair.depth(S.render_depth(K, H, W, T, **S.S_TABLE), T, K)
    pair.features(frames[0], T, K)
    assert pair.check_features(max_ulp=0) > 0
