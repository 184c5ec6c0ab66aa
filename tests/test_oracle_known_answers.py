"""Pins the CPU oracle against the reference's OWN known-answer tests, restated here (the reference cannot
be compiled offline and ships no stored numeric goldens for this path -- SURVEY.md 8(c)).  Each test
cites the reference test it restates."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import scenes as S
from tests.parity_utils import make_params

L = O.lib()


# ---- software binary16 vs numpy ---------------------------------------------------------------------
def test_half_conversion_matches_numpy_for_every_half_and_random_floats():
    hs = np.arange(65536, dtype=np.uint16)
    want = hs.view(np.float16).astype(np.float32)
    got = np.array([L.orc_h2f(int(h)) for h in hs], np.float32)
    ok = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert ok.all()
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(50000) * 10.0 ** rng.integers(-9, 6, 50000)).astype(np.float32)
    x = np.concatenate([x, want[~np.isnan(want)], np.float32([65504, 65519.99, 65520, 1e-8, 2.9802322e-8,
                                                              2.9802326e-8, 6.097555e-5])])
    with np.errstate(over='ignore'):
        want16 = x.astype(np.float16).view(np.uint16)
    got16 = np.array([L.orc_f2h(float(v)) for v in x], np.uint16)
    assert np.array_equal(got16, want16)


def test_half_arithmetic_is_correctly_rounded():
    rng = np.random.default_rng(1)
    a = rng.standard_normal(20000).astype(np.float16)
    b = (rng.standard_normal(20000) * 3).astype(np.float16)
    au, bu = a.view(np.uint16), b.view(np.uint16)
    for name, op in (('orc_hadd', np.add), ('orc_hsub', np.subtract), ('orc_hmul', np.multiply)):
        fn = getattr(L, name)
        got = np.array([fn(int(x), int(y)) for x, y in zip(au, bu)], np.uint16)
        want = op(a.astype(np.float64), b.astype(np.float64)).astype(np.float16).view(np.uint16)
        assert np.array_equal(got, want), name


# ---- test_ray_caster.cpp -----------------------------------------------------------------------------
def _cast(a, b):
    out = np.zeros((256, 3), np.int32)
    fa, fb = np.float32(a), np.float32(b)
    n = L.orc_raycast(fa.ctypes.data_as(O.C.POINTER(O.C.c_float)), fb.ctypes.data_as(O.C.POINTER(O.C.c_float)),
                      out.ctypes.data_as(O.C.POINTER(O.C.c_int32)), 256)
    return out[:n]


def test_raycaster_straight_ahead():    # StraightAheadCast :33-64
    idx = _cast([0, 0, 0], [5, 0, 0])
    assert len(idx) == 6
    assert np.array_equal(idx, np.stack([np.arange(6), np.zeros(6), np.zeros(6)], 1))
    neg = _cast([-0.0, -0.0, -0.0], [-5, -0.0, -0.0])      # "Finally, negative." -start, -end
    assert len(neg) == 6 and np.array_equal(idx, -neg)
    scaled = _cast([0, 0, 0], [5, 0, 0])                    # scale cancels: (2*p)/2 == p exactly
    assert np.array_equal(idx, scaled)


def test_raycaster_oblique_is_reversible():    # ObliqueCast :66-95
    fwd = _cast([0.5, -1.1, 3.1], [5.1, 0.2, 2.1])
    bwd = _cast([5.1, 0.2, 2.1], [0.5, -1.1, 3.1])
    assert len(fwd) == len(bwd)
    assert np.array_equal(fwd, bwd[::-1])
    # consecutive cells are face neighbours and the walk ends in the destination cell
    assert np.all(np.abs(np.diff(fwd, axis=0)).sum(1) == 1)
    assert np.array_equal(fwd[0], [0, -2, 3]) and np.array_equal(fwd[-1], [5, 0, 2])


def test_raycaster_zero_length():    # Length0Cast :97-106
    idx = _cast([0, 0, 0], [0, 0, 0])
    assert len(idx) == 1 and np.array_equal(idx[0], [0, 0, 0])


# ---- test_interpolation_2d.cpp ---------------------------------------------------------------------------
def test_bilinear_float_reproduces_coordinates():    # LinearInterpolation (float) :31-66, eps 1e-4
    rng = np.random.default_rng(2)
    for _ in range(1000):
        u, v = rng.uniform(0.5, 63.5), rng.uniform(0.5, 63.5)
        lx, ly = int(np.floor(u - 0.5)), int(np.floor(v - 0.5))
        if lx + 1 > 63 or ly + 1 > 63:
            continue
        x, y = np.float32(u - 0.5) - lx, np.float32(v - 0.5) - ly
        col = L.orc_interp_float(x, y, lx + 0.5, lx + 0.5, lx + 1.5, lx + 1.5)
        row = L.orc_interp_float(x, y, ly + 0.5, ly + 1.5, ly + 0.5, ly + 1.5)
        assert abs(col - u) < 1e-4 and abs(row - v) < 1e-4


def test_bilinear_half_reproduces_coordinates():    # LinearInterpolationHalfArray :126-129, eps 1e-1, 64x64
    rng = np.random.default_rng(3)
    h = lambda f: int(L.orc_f2h(float(f)))
    for _ in range(1000):
        u, v = rng.uniform(0.5, 63.5), rng.uniform(0.5, 63.5)
        lx, ly = int(np.floor(u - 0.5)), int(np.floor(v - 0.5))
        if lx + 1 > 63 or ly + 1 > 63:
            continue
        x, y = np.float32(u - 0.5) - lx, np.float32(v - 0.5) - ly
        col = L.orc_h2f(L.orc_interp_half(x, y, h(lx + 0.5), h(lx + 0.5), h(lx + 1.5), h(lx + 1.5)))
        row = L.orc_h2f(L.orc_interp_half(x, y, h(ly + 0.5), h(ly + 1.5), h(ly + 0.5), h(ly + 1.5)))
        assert abs(col - u) < 1e-1 and abs(row - v) < 1e-1


# ---- test_weighting_function.cpp / test_tsdf_integrator.cpp:359-474 ------------------------------------------
def test_weighting_functions():
    trunc = 0.4
    assert L.orc_weighting(0, 2.0, 1.0, trunc) == 1.0                      # constant
    assert L.orc_weighting(2, 2.0, 2.0, trunc) == pytest.approx(0.25, abs=1e-6)    # 1/z^2
    assert L.orc_weighting(2, 2.0, 0.005, trunc) == 1.0                    # z <= 1e-2 -> 1
    assert L.orc_weighting(2, 2.0, 2.5, trunc) == 0.0                      # behind the band -> 0
    assert L.orc_weighting(1, 2.0, 2.2, trunc) == pytest.approx(0.5, abs=1e-6)     # linear drop-off
    assert L.orc_weighting(1, 2.0, 1.5, trunc) == 1.0
    assert L.orc_weighting(1, 2.0, 2.41, trunc) == 0.0
    assert L.orc_weighting(3, 2.0, 2.2, trunc) == pytest.approx(0.5 / (2.2 * 2.2), rel=1e-5)
    assert L.orc_weighting(4, 2.0, 1.0, trunc) == pytest.approx(0.1, rel=1e-6)     # |sdf| >= trunc -> 0.1/z^2
    assert L.orc_weighting(5, 2.0, 4.0, trunc) == pytest.approx(0.25, rel=1e-6)    # linear with max


# ---- test_indexing.cpp ---------------------------------------------------------------------------------------
def test_indexing_roundtrip():
    rng = np.random.default_rng(4)
    bs = np.float32(0.16)
    for _ in range(2000):
        b = rng.integers(-50, 50, 3).astype(np.int32)
        v = rng.integers(0, 8, 3).astype(np.int32)
        c = np.zeros(3, np.float32)
        L.orc_voxel_center(bs, b.ctypes.data_as(O.C.POINTER(O.C.c_int32)), v.ctypes.data_as(O.C.POINTER(O.C.c_int32)),
                           c.ctypes.data_as(O.C.POINTER(O.C.c_float)))
        b2, v2 = np.zeros(3, np.int32), np.zeros(3, np.int32)
        L.orc_block_and_voxel(bs, c.ctypes.data_as(O.C.POINTER(O.C.c_float)),
                              b2.ctypes.data_as(O.C.POINTER(O.C.c_int32)), v2.ctypes.data_as(O.C.POINTER(O.C.c_int32)))
        assert np.array_equal(b, b2) and np.array_equal(v, v2)


# ---- test_feature_integrator.cpp -----------------------------------------------------------------------------
def _sphere_fixture(C_feat, alpha):
    """FeatureIntegratorTest fixture :55-129: analytic sphere TSDF (r=2 at (0,0,5)) in AABB (-5,-5,-5)..(10,15,5),
    voxel 0.2, truncation 2 voxels, camera 64x48 f=45 c=(32,24), identity pose."""
    vs = np.float32(0.2)
    bs = float(vs * 8)
    trunc = float(np.float32(2) * vs)
    mp, p = make_params(workspace=None, max_dist=7.0, alpha=alpha, raycast_sub=4)
    p.appearance_truncation_distance_vox = 2.0
    m = O.OracleMapper(float(vs), C_feat, p)
    lo, hi = np.array([-5, -5, -5.0]), np.array([10, 15, 5.0])
    bmin, bmax = np.floor(lo / bs).astype(int), np.floor(hi / bs).astype(int)
    g = (np.arange(8, dtype=np.float32) + 0.5) * vs
    for bx in range(bmin[0], bmax[0] + 1):
        for by in range(bmin[1], bmax[1] + 1):
            for bz in range(bmin[2], bmax[2] + 1):
                o = np.array([bx, by, bz], np.float32) * np.float32(bs)
                X, Y, Z = np.meshgrid(o[0] + g, o[1] + g, o[2] + g, indexing='ij')
                inside = ((X >= lo[0]) & (X <= hi[0]) & (Y >= lo[1]) & (Y <= hi[1]) & (Z >= lo[2]) & (Z <= hi[2]))
                d = np.sqrt(X ** 2 + Y ** 2 + (Z - 5.0) ** 2) - 2.0
                d = np.maximum(np.minimum(d, trunc), -trunc)     # scene_impl.h:73-81, scene.cpp:83-94
                blk = np.zeros((8, 8, 8, 2), np.float32)
                blk[..., 0] = np.where(inside, d, 0.0)
                blk[..., 1] = np.where(inside, 1.0, 0.0)
                m.set_tsdf_block((bx, by, bz), blk)
    K = np.array([[45, 0, 32], [0, 45, 24], [0, 0, 1]], np.float32)
    return m, K, np.eye(4, dtype=np.float32)


def test_feature_single_image_is_copied_exactly():    # IntegrateSingleFeatureImage :131-158
    C_feat = 16
    m, K, T = _sphere_fixture(C_feat, alpha=0.8)
    img = np.broadcast_to(np.arange(C_feat, dtype=np.float16), (48, 64, C_feat)).copy()
    m.add_feature_frame(img, T, K)
    _, blocks = m.all_blocks(1)
    w = blocks[..., -1].astype(np.float32)
    assert (w > 0).sum() > 0
    vals = blocks[w > 0][:, :C_feat].astype(np.float32)
    assert np.array_equal(vals, np.broadcast_to(np.arange(C_feat, dtype=np.float32), vals.shape))


def test_feature_three_frames_exponential_filter():    # IntegrateThreeTimes :160-203 (kExpected3 within 5e-3)
    C_feat = 8
    m, K, T = _sphere_fixture(C_feat, alpha=0.3)
    for v in (2.5, 1.3, 0.9):
        m.add_feature_frame(np.full((48, 64, C_feat), v, np.float16), T, K)
    e2 = 0.3 * 1.3 + 0.7 * 2.5
    e3 = 0.3 * 0.9 + 0.7 * e2
    _, blocks = m.all_blocks(1)
    w = blocks[..., -1].astype(np.float32)
    vals = blocks[w > 0][:, :C_feat].astype(np.float32)
    assert len(vals) > 0
    assert np.abs(vals - e3).max() < 5e-3


@pytest.mark.parametrize('value', [0.0, 65504.0, 6.1035e-5, -65504.0])
def test_feature_constant_corner_cases(value):    # CornerCase* :205-221
    C_feat = 8
    m, K, T = _sphere_fixture(C_feat, alpha=0.8)
    m.add_feature_frame(np.full((48, 64, C_feat), value, np.float16), T, K)
    _, blocks = m.all_blocks(1)
    w = blocks[..., -1].astype(np.float32)
    vals = blocks[w > 0][:, :C_feat].astype(np.float32)
    assert len(vals) > 0
    assert np.all(np.abs(vals - np.float32(np.float16(value))) <= abs(value) / 1e6)


# ---- test_tsdf_integrator.cpp ReconstructPlane :85-166 + test_mapper_masking.py :35-80,170-198 -----------------
def _plane_mapper(mask, C_feat=8):
    Hh, Ww = 480, 640
    fx = Ww / (2 * np.tan(np.deg2rad(90) / 2))
    K = np.array([[fx, 0, Ww / 2], [0, fx, Hh / 2], [0, 0, 1]], np.float32)
    T = np.eye(4, dtype=np.float32)
    p = O.default_params()    # reference defaults, as the python test uses Mapper(voxel_sizes_m=[0.05])
    m = O.OracleMapper(0.05, C_feat, p)
    m.add_depth_frame(np.full((Hh, Ww), 2.0, np.float32), T, K, mask)
    return m, K, T


def test_plane_reconstruction_and_depth_masking():
    ones = np.ones((480, 640), np.uint8)
    half = ones.copy()
    half[240:] = 0
    counts = {}
    for name, mask in (('all', ones), ('none', np.zeros_like(ones)), ('half', half), ('nomask', None)):
        m, _, _ = _plane_mapper(mask)
        m.update_feature_mesh()
        v, _, _ = m.get_feature_mesh()
        counts[name] = len(v)
        if name == 'all':
            assert len(v) > 0
            assert np.abs(v[:, 2] - 2.0).max() < 1e-3      # zero crossing on the plane
        if name == 'half':
            assert np.all(v[:, 1] <= 0.0)
    assert counts['none'] == 0
    assert counts['nomask'] == counts['all']
    assert abs(counts['half'] / counts['all'] - 0.5) < 0.01


def test_feature_masking_proportions():
    ones = np.ones((480, 640), np.uint8)
    half = ones.copy()
    half[240:] = 0
    props = {}
    for name, mask in (('all', ones), ('none', np.zeros_like(ones)), ('half', half)):
        m, K, T = _plane_mapper(None)
        m.add_feature_frame(np.ones((480, 640, 8), np.float16), T, K, mask)
        m.update_feature_mesh()
        v, f, _ = m.get_feature_mesh()
        props[name] = np.all(f == np.float16(1.0), axis=1).sum() / len(v)
    assert props['all'] > 0.85
    assert props['none'] == 0.0
    assert abs(props['half'] - 0.5) < 0.05


# ---- test_mapper_meshing.py :15-29 -----------------------------------------------------------------------------
def test_sphere_mesh_vertices_lie_on_sphere():
    vs = np.float32(0.05)
    bs = float(vs * 8)
    p = O.default_params()
    m = O.OracleMapper(float(vs), 8, p)
    trunc = float(np.float32(4) * vs)
    nb = int(np.ceil(1.4 / bs))
    g = (np.arange(8, dtype=np.float32) + 0.5) * vs
    for bx in range(-nb, nb):
        for by in range(-nb, nb):
            for bz in range(-nb, nb):
                o = np.array([bx, by, bz], np.float32) * np.float32(bs)
                X, Y, Z = np.meshgrid(o[0] + g, o[1] + g, o[2] + g, indexing='ij')
                d = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 1.0
                blk = np.zeros((8, 8, 8, 2), np.float32)
                blk[..., 0] = np.maximum(np.minimum(d, trunc), -trunc)
                blk[..., 1] = 1.0
                m.set_tsdf_block((bx, by, bz), blk)
    m.mark_all_dirty()
    m.update_feature_mesh()
    v, _, t = m.get_feature_mesh()
    assert len(v) > 1000 and len(t) > 0
    r = np.linalg.norm(v.astype(np.float64), axis=1)
    assert np.all(r - 1.0 < 1e-4)          # are_vertices_on_sphere: distance_off_sphere < eps
    assert np.abs(r - 1.0).max() < 2e-3    # and genuinely close (linear interpolation error at 5 cm voxels)


# ---- test_tsdf_decay.cpp DecayUntilRemoved -------------------------------------------------------------------------
def test_decay_until_removed():
    mp, p = make_params(workspace=S.WS_CUBE_STACKING, decay=0.9)
    m = O.OracleMapper(0.02, 8, p)
    K = S.intrinsics(64, 64)
    T = S.orbit_pose(0)
    m.add_depth_frame(S.render_depth(K, 64, 64, T, **S.S_TABLE), T, K)
    n0 = m.num_blocks(0)
    assert n0 > 0
    _, b0 = m.all_blocks(0)
    wsum = b0[..., 1].sum()
    m.decay()
    _, b1 = m.all_blocks(0)
    assert b1[..., 1].sum() < wsum                  # test_mapper_add_frames.py test_decay :59-78
    for _ in range(200):
        m.decay()
        if m.num_blocks(0) == 0:
            break
    assert m.num_blocks(0) == 0


# ---- test_frustum.cpp ViewpointCache ----------------------------------------------------------------------------------
def test_viewpoint_cache_tolerances():
    A = np.eye(4, dtype=np.float32)
    B = A.copy()
    B[0, 3] = 0.0009
    C_ = A.copy()
    C_[0, 3] = 0.0011
    fp = lambda a: a.ctypes.data_as(O.C.POINTER(O.C.c_float))
    assert L.orc_poses_close(fp(A.reshape(-1)), fp(B.reshape(-1)), 0.001, 0.1) == 1
    assert L.orc_poses_close(fp(A.reshape(-1)), fp(C_.reshape(-1)), 0.001, 0.1) == 0
    ang = np.deg2rad(0.05)
    R = np.eye(4, dtype=np.float32)
    R[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    assert L.orc_poses_close(fp(A.reshape(-1)), fp(R.reshape(-1).copy()), 0.001, 0.1) == 1
    ang = np.deg2rad(0.2)
    R[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    assert L.orc_poses_close(fp(A.reshape(-1)), fp(R.reshape(-1).copy()), 0.001, 0.1) == 0


def test_static_camera_stops_allocating():
    mp, p = make_params(workspace=S.WS_CUBE_STACKING)
    m = O.OracleMapper(0.02, 8, p)
    K = S.intrinsics(64, 64)
    T = S.look_at((0.1, 0.0, 0.45), (0.5, 0.0, 0.0))
    near = np.full((64, 64), 0.15, np.float32)                  # rays stop 23 cm from the camera
    far = S.render_depth(K, 64, 64, T, **S.S_TABLE)
    m.add_depth_frame(near, T, K)
    l1 = m.last_block_list(0)
    m.add_depth_frame(far, T, K)                                # cache hit: previous list although depth changed
    assert np.array_equal(l1, m.last_block_list(0))
    p.cache_last_viewpoint = 0
    m2 = O.OracleMapper(0.02, 8, p)
    m2.add_depth_frame(near, T, K)
    assert np.array_equal(l1, m2.last_block_list(0))
    m2.add_depth_frame(far, T, K)
    assert len(m2.last_block_list(0)) > len(l1)


def test_fused_half_variant_spread_is_bounded():
    """nvcc contracts the reference's `mul.f16` / `add.f16` pairs into HFMA2 (cuda_fp16.hpp emits them without
    .rn; SASS in profiles/r02_ref_contraction.md), so the fused sequence is the reference's and the oracle's /
    product's default.  This measures how far the UNCONTRACTED sequence (round 1's model) is from it on N(0,1)
    features: a few half-epsilons (2^-10) in absolute terms, but only ~60 % of the halves within 1 ulp -- ulp
    distance is unbounded where cancellation leaves a near-zero result.  So "within 1 fp16 ulp of nvblox" can
    only be met by implementing the reference's own contraction, which tests/test_ref_vectors.py pins bit for
    bit against reference-compiled vectors."""
    from tests.parity_utils import ordered_half
    res = []
    try:
        for fused in (0, 1):
            L.orc_set_fused_half(fused)
            mp, p = make_params(workspace=S.WS_CUBE_STACKING, alpha=0.3)
            m = O.OracleMapper(0.02, 32, p)
            K = S.intrinsics(64, 64)
            for i in range(2):
                T = S.orbit_pose(i)
                m.add_depth_frame(S.render_depth(K, 64, 64, T, **S.S_TABLE), T, K)
                m.add_feature_frame(S.feature_frame(64, 64, 32, 10 + i), T, K)
            res.append(m.all_blocks(1)[1])
    finally:
        L.orc_set_fused_half(-1)    # back to "follow the floating-point model"
    a, b = res[0].astype(np.float32), res[1].astype(np.float32)
    assert np.array_equal(res[0][..., -1].view(np.uint16), res[1][..., -1].view(np.uint16))    # weights identical
    assert np.isfinite(a).all() and np.isfinite(b).all()
    assert np.abs(a - b).max() < 8 * 2.0 ** -10
    d = np.abs(ordered_half(res[0].view(np.uint16)).astype(np.int64) -
               ordered_half(res[1].view(np.uint16)).astype(np.int64))
    assert (d <= 1).mean() > 0.6
    assert (d > 1).any()


# ---- colour path: test_color_integrator.cpp, test_mapper_masking.py:100-160 --------------------
def _color_blocks(m):
    _, rgb, w = m.all_color_blocks()
    return rgb.reshape(-1, 3), w.reshape(-1)


def test_color_solid_image_is_copied_exactly():    # IntegrateColorToGroundTruthDistanceField :225-296 (one view)
    m, K, T = _sphere_fixture(8, alpha=0.8)
    red = np.zeros((48, 64, 3), np.uint8)
    red[..., 0] = 255
    m.add_color_frame(red, T, K)
    rgb, w = _color_blocks(m)
    assert (w > 0).sum() > 0
    assert np.all(rgb[w > 0] == [255, 0, 0])               # every observed voxel has exactly the image colour
    assert np.all(rgb[w == 0] == 127)                      # untouched voxels stay Gray (blox.cu:32-54)
    assert np.all(w[w > 0] == np.float32(0.8))             # WeightingFunction :510-580: weight == measurement weight
    # every colour block has a TSDF block (:298-301)
    tsdf_idx = {tuple(r) for r in m.block_indices(0)}
    assert all(tuple(r) in tsdf_idx for r in m.block_indices(2))


def test_color_exponential_filter_rounding():
    """weightedSum(uint8_t ...) projective_appearance_integrator.cu:277-284 with binary16 blend weights (:301-302):
    value_2 = round(200 * half(0.7) + 100 * half(0.3)), value_3 likewise from value_2."""
    m, K, T = _sphere_fixture(8, alpha=0.3)
    h1 = float(np.float32(np.float16(np.float32(0.7) / (np.float32(0.7) + np.float32(0.3)))))
    h2 = float(np.float32(np.float16(np.float32(0.3) / (np.float32(0.7) + np.float32(0.3)))))
    expect = None
    for v in (200, 100, 31):
        m.add_color_frame(np.full((48, 64, 3), v, np.uint8), T, K)
        expect = v if expect is None else int(np.floor(np.float32(np.float32(expect) * np.float32(h1) +
                                                                  np.float32(v) * np.float32(h2)) + 0.5))
    rgb, w = _color_blocks(m)
    assert (w > 0).sum() > 0
    assert np.all(rgb[w > 0] == expect)
    assert np.all(w[w > 0] == np.float32(np.float32(np.float32(0.3) + np.float32(0.3)) + np.float32(0.3)))


def test_color_occlusion():    # OcclusionTesting :442-508: the sphere behind the first one is never painted
    vs = np.float32(0.1)
    bs = float(vs * 8)
    trunc = float(np.float32(4) * vs)
    p = O.default_params()
    m = O.OracleMapper(float(vs), 8, p)
    g = (np.arange(8, dtype=np.float32) + 0.5) * vs
    c1, c2 = np.array([5.0, 0, 0]), np.array([10.0, 0, 0])
    for bx in range(2, 16):
        for by in range(-4, 4):
            for bz in range(-4, 4):
                o = np.array([bx, by, bz], np.float32) * np.float32(bs)
                X, Y, Z = np.meshgrid(o[0] + g, o[1] + g, o[2] + g, indexing='ij')
                d1 = np.sqrt((X - c1[0]) ** 2 + Y ** 2 + Z ** 2) - 2.0
                d2 = np.sqrt((X - c2[0]) ** 2 + Y ** 2 + Z ** 2) - 2.0
                d = np.clip(np.minimum(d1, d2), -trunc, trunc)
                blk = np.zeros((8, 8, 8, 2), np.float32)
                blk[..., 0], blk[..., 1] = d, 1.0
                m.set_tsdf_block((bx, by, bz), blk)
    T = np.eye(4, dtype=np.float32)       # camera z -> world +x (rotation by +90 deg about y)
    T[:3, :3] = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], np.float32)
    K = np.array([[300, 0, 320], [0, 300, 240], [0, 0, 1]], np.float32)
    red = np.zeros((480, 640, 3), np.uint8)
    red[..., 0] = 255
    m.add_color_frame(red, T, K)
    idx, rgb, w = m.all_color_blocks()
    centers = (idx[:, None, None, None, :].astype(np.float32) * 8 +
               np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing='ij'), -1) + 0.5) * vs
    r1 = np.linalg.norm(centers - c1, axis=-1)
    r2 = np.linalg.norm(centers - c2, axis=-1)
    front = (np.abs(r1 - 2.0) < 0.1) & (centers[..., 0] < 4.0)
    assert (w[front] > 0).mean() > 0.2 and np.all(rgb[front & (w > 0)] == [255, 0, 0])
    back = np.abs(r2 - 2.0) < 0.3
    assert back.sum() > 0 and np.all(w[back] == 0.0)


def test_color_masking_proportions():    # test_mapper_color_masking :132-160
    ones = np.ones((480, 640), np.uint8)
    half = ones.copy()
    half[240:] = 0
    red = np.zeros((480, 640, 3), np.uint8)
    red[..., 0] = 255
    props = {}
    for name, mask in (('all', ones), ('none', np.zeros_like(ones)), ('half', half)):
        m, K, T = _plane_mapper(None)
        m.update_color_mesh()                               # map_plane_with_mask ends with update_color_mesh
        m.add_color_frame(red, T, K, mask)
        m.update_color_mesh()
        v, c, t = m.get_color_mesh()
        assert len(v) > 0 and len(t) > 0
        props[name] = (c[:, 0] == 255).sum() / len(v)
        if name == 'none':
            assert np.all(c == 127)                         # Gray where nothing was painted
    assert props['all'] > 0.85
    assert props['none'] == 0.0
    assert abs(props['half'] - 0.5) < 0.05


def test_color_mesh_is_a_separate_layer():
    """blocks_to_update_tracker.cpp:32-60: both mesh layers have their own dirty set; updating one leaves the other
    stale, and both meshes share the geometry once both are updated."""
    m, K, T = _plane_mapper(None)
    m.update_feature_mesh()
    v_f, _, _ = m.get_feature_mesh()
    v_c, _, _ = m.get_color_mesh()
    assert len(v_f) > 0 and len(v_c) == 0
    m.update_color_mesh()
    v_c, c, _ = m.get_color_mesh()
    assert np.array_equal(v_c, v_f) and np.all(c == 127)


# ---- N4: the extractor's bilinear up-sampling (PyTorch's algorithm, restated in oracle.upsample_bilinear) --------
def _torch_cpu_chained(low_bchw, size, C_feat):
    import torch
    import torch.nn.functional as F
    up = F.interpolate(low_bchw, size=size, mode='bilinear', align_corners=False)[0].permute(1, 2, 0)
    if C_feat > up.shape[2]:
        up = torch.cat((up, torch.zeros(size[0], size[1], C_feat - up.shape[2])), dim=2)
    return up.contiguous().to(torch.float16).numpy()


@pytest.mark.parametrize('shape', [(32, 32, 24, 32, 512, 512), (37, 37, 16, 16, 518, 518), (7, 5, 8, 16, 64, 48)])
@pytest.mark.parametrize('mode', [0, 1])
def test_upsample_restatement_against_torch_cpu(shape, mode):
    """torch's CPU kernel orders its multiply-adds differently from its CUDA kernels (which the oracle follows,
    and which tests/test_gpu_upsample.py checks bit for bit on the GPU): > 99.9 % of the halves are identical and
    the rest are within fp32 rounding noise of the same value."""
    import torch
    lh, lw, lc, C_feat, H, W = shape
    torch.manual_seed(lh * 100 + lc)
    low = torch.randn(1, lc, lh, lw)
    want = _torch_cpu_chained(low, (H, W), C_feat)
    got = O.upsample_bilinear(low[0].permute(1, 2, 0).numpy(), C_feat, H, W, mode).view(np.float16)
    same = got.view(np.uint16) == want.view(np.uint16)
    assert same.mean() > 0.999
    np.testing.assert_allclose(got.astype(np.float32), want.astype(np.float32), rtol=2e-3, atol=2e-3)
    assert not got[..., lc:].any()          # zero-padded channels


def test_upsample_known_answers():
    """align_corners=False semantics (UpSample.cuh:114-130): source index (dst + 0.5) * in/out - 0.5 clamped at 0,
    right / bottom neighbour clamped at the border."""
    low = np.array([[[0.0], [1.0]], [[2.0], [3.0]]], np.float32)       # 2 x 2 x 1
    out = O.upsample_bilinear(low, 1, 4, 4, 0).view(np.float16)[..., 0].astype(np.float32)
    col = np.array([0.0, 0.25, 0.75, 1.0], np.float32)                 # lambda along one axis for 2 -> 4
    want = col[None, :] * 1.0 + col[:, None] * 2.0
    assert np.array_equal(out, want)
    const = np.full((3, 5, 8), 0.7, np.float32)
    out = O.upsample_bilinear(const, 8, 24, 40, 1).view(np.float16)
    assert np.all(out == np.float16(0.7))
    # bf16 tensors: the kernel's bf16 store precedes the cast to fp16 (two roundings)
    v = np.full((2, 2, 8), 1.0 + 2.0 ** -9, np.float32)               # representable in fp16, not in bf16
    assert np.all(O.upsample_bilinear(v, 8, 4, 4, 0).view(np.float16) == np.float16(1.0 + 2.0 ** -9))
    assert np.all(O.upsample_bilinear(v, 8, 4, 4, 2).view(np.float16) == np.float16(1.0))


def test_lowres_descriptor_follows_torchs_kernel_choice():
    """The wrapper must predict which CUDA kernel F.interpolate would have run (NHWC iff channels-last strides and
    >= 16 channels) and pass channels-last memory through zero-copy."""
    import torch
    from torch._prims_common import suggest_memory_format
    from nvblox_torch.mapper import _strides_like_channels_last, lowres_descriptor
    feats = torch.randn(1, 16 * 16, 48)                                 # RADIO: [b, hw, c]
    bchw = feats.view(1, 16, 16, -1).permute(0, 3, 1, 2)                # feature_extraction.py:328-330
    for t in (bchw, bchw.contiguous(), torch.randn(1, 8, 4, 4), torch.randn(1, 1, 4, 4),
              torch.randn(1, 8, 4, 4).contiguous(memory_format=torch.channels_last), torch.randn(1, 24, 1, 1)):
        assert _strides_like_channels_last(t) == (suggest_memory_format(t) == torch.channels_last), t.stride()
    x, c, h, w, dtype, layout, kernel = lowres_descriptor(bchw)
    assert (c, h, w, dtype, layout, kernel) == (48, 16, 16, 0, 1, 1) and x.data_ptr() == feats.data_ptr()
    x, c, h, w, dtype, layout, kernel = lowres_descriptor(bchw.contiguous().half())
    assert (dtype, layout, kernel) == (1, 0, 0)
    small = torch.randn(1, 4, 4, 8).permute(0, 3, 1, 2)                 # channels-last but < 16 channels: NCHW kernel
    assert lowres_descriptor(small)[4:] == (0, 1, 0)
    assert lowres_descriptor(torch.randn(24, 4, 4).bfloat16())[4:] == (2, 0, 0)


# ---- test_tsdf_integrator.cpp:359-474 (WeightingFunction), integration level ----------------------------------
def _cam_640x480_f300():
    return np.array([[300, 0, 320], [0, 300, 240], [0, 0, 1]], np.float32)


def _voxel_centres(idx, voxel):
    """[N, 8, 8, 8, 3] voxel centre positions of blocks `idx` (getCenterPositionFromBlockIndexAndVoxelIndex)."""
    g = (np.arange(8, dtype=np.float32) + np.float32(0.5)) * np.float32(voxel)
    o = idx.astype(np.float32)[:, None, None, None, :] * np.float32(8 * voxel)
    X, Y, Z = np.meshgrid(g, g, g, indexing='ij')
    return o + np.stack([X, Y, Z], -1)[None]


def test_tsdf_weights_constant_and_inverse_square_on_a_plane():
    K, T = _cam_640x480_f300(), np.eye(4, dtype=np.float32)
    depth = np.full((480, 640), 5.0, np.float32)       # plane z = 5 seen from the origin: depth == 5 everywhere
    for mode, name in ((0, 'constant'), (2, 'inverse_square')):
        p = O.default_params()
        p.max_weight = 100.0
        p.weighting_mode = mode
        m = O.OracleMapper(0.2, 8, p)
        m.add_depth_frame(depth, T, K)
        idx, data = m.all_blocks(0)
        assert len(idx) > 0
        w = data[..., 1]
        seen = w > 1e-4
        assert seen.sum() > 1000
        if name == 'constant':
            assert np.abs(w[seen] - 1.0).max() < 1e-4
        else:
            z = _voxel_centres(idx, 0.2)[..., 2]
            assert np.abs(w[seen] - 1.0 / (z[seen] * z[seen])).max() < 1e-4      # hand-computed 1 / depth^2


# ---- test_workspace_bounds.cpp:40-118 (CheckAllocatedBlocks) --------------------------------------------------
def test_workspace_bounds_restrict_the_allocated_blocks():
    K, T = _cam_640x480_f300(), np.eye(4, dtype=np.float32)
    depth = np.full((480, 640), 5.0, np.float32)
    lo, hi = np.float32([-3, -3, 2]), np.float32([3, 3, 4])
    counts = {}
    for kind, name in ((0, 'unbounded'), (1, 'height'), (2, 'box')):
        p = O.default_params()
        p.workspace_bounds_type = kind
        for i in range(3):
            p.workspace_min[i], p.workspace_max[i] = float(lo[i]), float(hi[i])
        m = O.OracleMapper(0.1, 8, p)
        m.add_depth_frame(depth, T, K)
        idx, _ = m.all_blocks(0)
        assert len(idx) > 0
        # every allocated block has at least one voxel (corner position, as the reference test computes it)
        # inside the bounds
        g = np.arange(8, dtype=np.float32) * np.float32(0.1)
        pos_lo = idx.astype(np.float32) * np.float32(0.8)
        pos_hi = pos_lo + g[-1]
        if name == 'height':
            assert np.all((pos_hi[:, 2] >= lo[2]) & (pos_lo[:, 2] <= hi[2]))
        elif name == 'box':
            assert np.all((pos_hi >= lo).all(1) & (pos_lo <= hi).all(1))
        counts[name] = len(idx)
    assert counts['box'] > 0
    assert counts['height'] > counts['box'] and counts['unbounded'] > counts['box']


# ---- test_mesh.cpp:101-155 (PlaneMesh) and :423-462 (WeldingTest) ---------------------------------------------
def _plane_x0_layer(m, voxel, weld_scene=False):
    """generateLayerFromScene for the tests' scenes inside the 6 x 6 x 3 m AABB: plane x = 0 with normal -x
    (distance = -x); WeldingTest adds the plane y = 0.1 (normal -y) and the sphere c = (-2, -2, 0), r = 2."""
    trunc = np.float32(4 * voxel)
    bs = 8 * voxel
    nb = int(round(3.0 / bs)) + 1
    g = (np.arange(8, dtype=np.float32) + np.float32(0.5)) * np.float32(voxel)
    n = 0
    for bx in range(-nb, nb):
        for by in range(-nb, nb):
            for bz in range(0, nb):
                o = np.float32([bx, by, bz]) * np.float32(bs)
                X, Y, Z = np.meshgrid(o[0] + g, o[1] + g, o[2] + g, indexing='ij')
                d = -X
                if weld_scene:
                    d = np.minimum(d, -(Y - np.float32(0.1)))
                    d = np.minimum(d, np.sqrt((X + 2) ** 2 + (Y + 2) ** 2 + Z ** 2) - np.float32(2.0))
                if np.abs(d).min() > trunc + bs:
                    continue
                blk = np.zeros((8, 8, 8, 2), np.float32)
                blk[..., 0] = np.clip(d, -trunc, trunc)
                blk[..., 1] = 1.0
                m.set_tsdf_block((bx, by, bz), blk)
                n += 1
    return n


def test_plane_mesh_vertices_lie_on_the_plane():
    p = O.default_params()
    p.mesh_weld_vertices = 0
    m = O.OracleMapper(0.1, 8, p)
    n_sdf = _plane_x0_layer(m, 0.1)
    m.mark_all_dirty()
    m.update_feature_mesh()
    v, _, t, vb = m.get_feature_mesh(with_block_index=True)
    n_mesh_blocks = len(np.unique(vb, axis=0))
    assert 0 < n_mesh_blocks <= n_sdf
    assert len(v) > 0 and len(v) == 3 * len(t)          # unwelded: one vertex per triangle corner
    assert np.abs(v[:, 0]).max() < 1e-4                   # EXPECT_NEAR(vertex.x(), 0.0, kFloatEpsilon)


def test_welding_removes_duplicate_vertices_in_every_block():
    counts = {}
    for weld in (0, 1):
        p = O.default_params()
        p.mesh_weld_vertices = weld
        m = O.OracleMapper(0.1, 8, p)
        _plane_x0_layer(m, 0.1, weld_scene=True)
        m.mark_all_dirty()
        m.update_feature_mesh()
        v, _, t, vb = m.get_feature_mesh(with_block_index=True)
        blocks, n_per_block = np.unique(vb, axis=0, return_counts=True)
        counts[weld] = {tuple(b): int(c) for b, c in zip(blocks, n_per_block)}
        counts[weld, 'tris'] = len(t)
    assert counts[0, 'tris'] == counts[1, 'tris'] > 0     # welding re-indexes, it does not drop triangles
    assert counts[0].keys() == counts[1].keys() and len(counts[0]) > 0
    for b, pre in counts[0].items():                       # EXPECT_LT(num_vertices_postweld, num_vertices_preweld)
        assert counts[1][b] < pre, b


# ---- test_tsdf_integrator.cpp:168-263 (SphereSceneTest) -------------------------------------------------------
def _sphere_in_box_depth(K, T, H, W, max_dist=10.0):
    """Depth image of test_utils::getSphereInBox(): sphere c = (0, 0, 2), r = 2 inside the box
    [-5, 5] x [-5, 5] x [0, 5] seen from INSIDE the box (generateDepthImageFromScene, rays cut at 10 m)."""
    u = (np.arange(W, dtype=np.float64) + 0.5 - K[0, 2]) / K[0, 0]
    v = (np.arange(H, dtype=np.float64) + 0.5 - K[1, 2]) / K[1, 1]
    uu, vv = np.meshgrid(u, v)
    d = np.stack([uu, vv, np.ones_like(uu)], -1) @ T[:3, :3].astype(np.float64).T
    o = T[:3, 3].astype(np.float64)
    lo, hi = np.array([-5.0, -5.0, 0.0]), np.array([5.0, 5.0, 5.0])
    with np.errstate(divide='ignore', invalid='ignore'):
        t_wall = np.where(d > 0, (hi - o) / d, np.where(d < 0, (lo - o) / d, np.inf)).min(-1)
    oc = o - np.array([0.0, 0.0, 2.0])
    a, b, c = (d * d).sum(-1), 2.0 * (d * oc).sum(-1), (oc * oc).sum() - 4.0
    disc = b * b - 4 * a * c
    with np.errstate(invalid='ignore'):
        t_s = (-b - np.sqrt(disc)) / (2 * a)
    t_s[~(disc >= 0) | ~(t_s > 1e-6)] = np.inf
    t = np.minimum(t_wall, t_s)
    norm = np.sqrt((d * d).sum(-1))                         # ray length per unit of depth
    t[t * norm > max_dist] = 0.0
    return t.astype(np.float32)


def test_sphere_in_box_orbit_matches_the_analytic_tsdf():
    K = _cam_640x480_f300()
    voxel, trunc = 0.2, 0.4
    p = O.default_params()
    p.truncation_distance_vox = 2.0
    m = O.OracleMapper(voxel, 8, p)
    base = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]], np.float64)      # quaternion (w, x, y, z) = (.5, .5, .5, .5)
    n_poses = 80
    for i in range(n_poses):
        th = 2 * np.pi / n_poses * i
        a = np.pi + th
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = (Rz @ base).astype(np.float32)
        T[:3, 3] = np.float32([4 * np.cos(th), 4 * np.sin(th), 2.0])
        m.add_depth_frame(_sphere_in_box_depth(K, T, 480, 640), T, K)
    idx, data = m.all_blocks(0)
    c = _voxel_centres(idx, voxel).astype(np.float64)
    inside = ((c >= [-5, -5, 0]) & (c <= [5, 5, 5])).all(-1)           # the ground-truth layer covers the scene AABB
    gt = np.minimum.reduce([np.sqrt(c[..., 0] ** 2 + c[..., 1] ** 2 + (c[..., 2] - 2) ** 2) - 2.0,
                            c[..., 2], 5 - c[..., 2], c[..., 0] + 5, 5 - c[..., 0], c[..., 1] + 5, 5 - c[..., 1]])
    gt = np.clip(gt, -trunc, trunc)
    sel = (data[..., 1] >= 1.0) & inside                               # kMinWeight = 1.0
    assert sel.sum() > 20000
    big = np.abs(data[..., 0][sel] - gt[sel]) > trunc                  # kDistanceErrorTolerance = truncation
    assert 100.0 * big.mean() < 0.4                                    # kAcceptablePercentageOverThreshold
    # tighter than the reference asks: the fused distances are unbiased estimates of the analytic field
    assert np.abs(data[..., 0][sel] - gt[sel]).mean() < 0.25 * voxel
