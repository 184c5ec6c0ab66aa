"""SURVEY 8(f) N4 on the GPU: the extractor's bilinear up-sampling fused into the feature gather.

Three layers of evidence, all bit-exact (tolerance 0 fp16 ulp):
  1. nvbx_upsample_features == torch's own CUDA `F.interpolate(..., 'bilinear', align_corners=False)` followed
     by mindmap's rearrange / zero-pad / `.to(float16)` (feature_extraction.py:188-210,
     nvblox_mapping_helpers.py:256), for both torch kernels (NCHW / NHWC) and the three dtypes -- this is the
     reference's real upstream step executed on this box, not a restatement;
  2. the same frame == the CPU oracle's restatement (oracle.upsample_bilinear);
  3. a map integrated with add_feature_frame_lowres == a map integrated with the chained path
     (torch up-sampling -> add_feature_frame), and == the oracle fed the oracle's up-sampled frame.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import scenes as S
from tests.parity_utils import Pair, gpu_blocks, make_params

pytestmark = pytest.mark.gpu


def _mapper(C_feat, **kw):
    from nvblox_torch.constants import constants
    from nvblox_torch.mapper import Mapper
    constants.set_feature_array_num_elements(C_feat)
    mp, _ = make_params(workspace=S.WS_CUBE_STACKING, **kw)
    return Mapper(voxel_sizes_m=0.02, mapper_parameters=mp)


def _chained(low_bchw, size, C_feat):
    """mindmap's own tail of FeatureExtractor.compute() + the cast in nvblox_mapping_helpers.py:256."""
    import torch
    import torch.nn.functional as F
    up = F.interpolate(low_bchw, size=size, mode='bilinear', align_corners=False)
    hwc = up[0].permute(1, 2, 0)
    pad = C_feat - hwc.shape[2]
    if pad:
        hwc = torch.cat((hwc, torch.zeros(size[0], size[1], pad, device=hwc.device)), dim=2)
    return hwc.contiguous().to(dtype=torch.float16)


def _lowres(lc, lh, lw, dtype, channels_last, seed=0):
    import torch
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    x = torch.randn((1, lc, lh, lw), generator=g, device='cuda', dtype=torch.float32).to(dtype)
    if channels_last:   # what RADIO / DINOv2 hand over: a permuted view of [1, h, w, c] (feature_extraction.py:328)
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return x


@pytest.mark.parametrize('dtype_name', ['float32', 'float16', 'bfloat16'])
@pytest.mark.parametrize('channels_last', [False, True])
@pytest.mark.parametrize('shape', [(32, 32, 768, 768, 512, 512), (37, 37, 24, 32, 518, 518), (16, 16, 8, 16, 256, 256),
                                   (7, 5, 40, 40, 64, 48)])
def test_upsampled_frame_equals_torch_cuda(dtype_name, channels_last, shape):
    import torch
    lh, lw, lc, C_feat, H, W = shape
    dtype = getattr(torch, dtype_name)
    m = _mapper(C_feat)
    low = _lowres(lc, lh, lw, dtype, channels_last, seed=lc + lh)
    want = _chained(low, (H, W), C_feat)
    got = m.upsample_features(low, (H, W))
    torch.cuda.synchronize()
    diff = (got.view(torch.int16) != want.view(torch.int16))
    n_bad = int(diff.sum())
    assert n_bad == 0, f'{n_bad} of {diff.numel()} halves differ from torch ({dtype_name}, channels_last={channels_last})'


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_upsampled_frame_equals_oracle(mode):
    import torch
    lh, lw, lc, C_feat, H, W = 16, 12, 24, 32, 128, 96
    m = _mapper(C_feat)
    dtype = torch.bfloat16 if mode == 2 else torch.float32
    low = _lowres(lc, lh, lw, dtype, channels_last=(mode == 1), seed=5)
    got = m.upsample_features(low, (H, W)).cpu().numpy().view(np.uint16)
    want = O.upsample_bilinear(low[0].permute(1, 2, 0).float().cpu().numpy(), C_feat, H, W, mode)
    assert np.array_equal(got, want)


@pytest.mark.parametrize('alpha,strict,channels_last', [(1.0, False, True), (1.0, True, False), (0.3, False, True)])
def test_fused_integration_equals_chained_and_oracle(alpha, strict, channels_last):
    """Orbit sequence: map A gets torch's up-sampled frames through add_feature_frame, map B the low-res maps
    through add_feature_frame_lowres, the oracle gets the oracle's up-sampled frames.  All three agree bit for bit."""
    import torch
    C_feat, lc, H, W, lh, lw = 64, 48, 128, 128, 8, 8
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=alpha, strict=strict)
    pair = Pair(0.02, C_feat, mp, op)          # pair.gpu is map B
    from nvblox_torch.mapper import Mapper
    chained = Mapper(voxel_sizes_m=0.02, mapper_parameters=mp)
    K = S.intrinsics(W, H)
    K_t = torch.from_numpy(K)
    mask = np.ones((H, W), np.uint8)
    mask[:, : W // 8] = 0
    for i in range(6):
        T = S.orbit_pose(i)
        T_t = torch.from_numpy(T)
        depth = S.render_depth(K, H, W, T, **S.S_TABLE)
        d_t = torch.from_numpy(depth).cuda()
        low = _lowres(lc, lh, lw, torch.float32, channels_last, seed=100 + i)
        fm = mask if i % 2 else None
        fm_t = None if fm is None else torch.from_numpy(fm).cuda()
        chained.add_depth_frame(d_t, T_t, K_t)
        chained.add_feature_frame(_chained(low, (H, W), C_feat), T_t, K_t, fm_t)
        pair.gpu.add_depth_frame(d_t, T_t, K_t)
        pair.gpu.add_feature_frame_lowres(low, (H, W), T_t, K_t, fm_t)
        pair.cpu.add_depth_frame(depth, T, K)
        frame = O.upsample_bilinear(low[0].permute(1, 2, 0).cpu().numpy(), C_feat, H, W, 1 if channels_last else 0)
        pair.cpu.add_feature_frame(frame.view(np.float16), T, K, fm)
    ai, ad = gpu_blocks(chained.feature_layer_view(0))
    bi, bd = gpu_blocks(pair.gpu.feature_layer_view(0))
    assert len(ai) > 0 and np.array_equal(ai, bi)
    assert np.array_equal(ad.view(np.uint16), bd.view(np.uint16)), 'fused and chained feature maps differ'
    assert chained.counters(0)['feature_voxels_updated'] == pair.gpu.counters(0)['feature_voxels_updated'] > 0
    assert pair.check_features(max_ulp=0 if strict or alpha != 1.0 else 1) > 0
    assert pair.check_mesh() > 0


def test_host_lowres_entry_and_errors():
    import torch
    from nvblox_mindmap_b200._capi import NvbxError
    C_feat, lc, H, W, lh, lw = 32, 32, 64, 64, 4, 4
    a, b = _mapper(C_feat), _mapper(C_feat)
    K = S.intrinsics(W, H)
    K_t, T = torch.from_numpy(K), S.orbit_pose(0)
    T_t = torch.from_numpy(T)
    depth = torch.from_numpy(S.render_depth(K, H, W, T, **S.S_TABLE))
    low = _lowres(lc, lh, lw, torch.float16, False, seed=3)
    a.add_depth_frame(depth.cuda(), T_t, K_t)
    a.add_feature_frame_lowres(low, (H, W), T_t, K_t)
    b.integrate_frame_from_host_lowres(depth.pin_memory(), low.cpu().pin_memory(), T_t, K_t)
    ai, ad = gpu_blocks(a.feature_layer_view(0))
    bi, bd = gpu_blocks(b.feature_layer_view(0))
    assert len(ai) > 0 and np.array_equal(ai, bi) and np.array_equal(ad.view(np.uint16), bd.view(np.uint16))
    with pytest.raises(NvbxError):   # more channels than the map stores
        a.upsample_features(_lowres(C_feat + 8, lh, lw, torch.float32, False), (H, W))
    with pytest.raises(NvbxError):   # same size: torch copies, callers use add_feature_frame
        a.upsample_features(_lowres(lc, H, W, torch.float32, False), (H, W))
