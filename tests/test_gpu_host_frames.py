"""Host-resident frames through nvbx_integrate_frame_host (the end-to-end entry bench.py times).

A pinned feature frame is fetched sparsely over PCIe (k_pixel_mark / k_pixel_fetch: only the pixels the frame's
voxels sample); a pageable frame, or host_fetch_mode 'dense', is copied whole.  Both must build exactly the map the
CPU oracle builds from the same frames, and the number of pixels the sparse path moves must equal the oracle's
count of DISTINCT pixels read (SURVEY.md 8(d) `U_px`).
"""
import numpy as np
import pytest

from tests import scenes as S
from tests.parity_utils import Pair, make_params, orbit_frames

pytestmark = pytest.mark.gpu


def _host_frames(pair, frames, mode, pinned=True, feature_mask=None, depth_mask=None):
    import torch
    pair.gpu.set_host_fetch_mode(mode)
    fetched, distinct = 0, 0
    for _, T, K, depth, feat in frames:
        hd, hf = torch.from_numpy(depth), torch.from_numpy(feat)
        hfm = None if feature_mask is None else torch.from_numpy(feature_mask)
        hdm = None if depth_mask is None else torch.from_numpy(depth_mask)
        if pinned:
            hd, hf = hd.pin_memory(), hf.pin_memory()
            hfm = None if hfm is None else hfm.pin_memory()
            hdm = None if hdm is None else hdm.pin_memory()
        before = pair.gpu.counters(0)['host_pixels_fetched']
        pair.gpu.integrate_frame_from_host(hd, hf, torch.from_numpy(T), torch.from_numpy(K), depth_mask=hdm,
                                           feature_mask=hfm)
        fetched += pair.gpu.counters(0)['host_pixels_fetched'] - before      # also syncs: hf may now be released
        pair.cpu.add_depth_frame(depth, T, K, depth_mask)
        pair.cpu.add_feature_frame(feat, T, K, feature_mask)
        distinct += pair.cpu.counters()['last_distinct_pixels']
        pair.decay()
    return fetched, distinct


@pytest.mark.parametrize('C_feat,size', [(768, 160), (1024, 96), (64, 128), (40, 96)])
def test_sparse_host_fetch_matches_oracle(C_feat, size):
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, strict=True)
    pair = Pair(0.02, C_feat, mp, op)
    frames = list(orbit_frames(4, size, size, C_feat, S.S_TABLE))
    fetched, distinct = _host_frames(pair, frames, 'sparse')
    assert pair.check_tsdf() > 0
    assert pair.check_features(max_ulp=0) > 0
    assert pair.check_mesh() > 0
    g, c = pair.gpu.counters(0), pair.cpu.counters()
    assert g['feature_voxels_updated'] == c['feature_voxels_updated'] > 0
    assert fetched == distinct > 0, f'pixels over PCIe {fetched} != distinct pixels read (oracle) {distinct}'
    assert fetched < 4 * size * size        # fewer than the dense copies would have moved


def test_sparse_fetch_with_masks_and_blend():
    size, C_feat = 128, 768
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=0.3)
    pair = Pair(0.02, C_feat, mp, op)
    fmask = np.ones((size, size), np.uint8)
    fmask[size // 2:, :] = 0
    fmask[:, :6] = 0
    dmask = np.ones((size, size), np.uint8)
    dmask[:, size - 9:] = 0
    frames = list(orbit_frames(5, size, size, C_feat, S.S_TABLE))
    fetched, distinct = _host_frames(pair, frames, 'sparse', feature_mask=fmask, depth_mask=dmask)
    pair.check_tsdf()
    assert pair.check_features(max_ulp=1) > 0
    assert fetched == distinct > 0


@pytest.mark.parametrize('mode,pinned', [('dense', True), ('sparse', False)])
def test_dense_and_pageable_frames(mode, pinned):
    """`dense` mode and pageable (unregistered) memory take the whole-frame copy: same map, nothing fetched sparsely."""
    size, C_feat = 128, 768
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, strict=True)
    pair = Pair(0.02, C_feat, mp, op)
    frames = list(orbit_frames(3, size, size, C_feat, S.S_TABLE))
    fetched, _ = _host_frames(pair, frames, mode, pinned=pinned)
    pair.check_tsdf()
    assert pair.check_features(max_ulp=0) > 0
    assert fetched == 0


def test_sparse_then_device_frames_interleave():
    """The pixel bitmap is left clean by every sparse frame and the device-frame entry is unaffected by it."""
    import torch
    size, C_feat = 128, 768
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, strict=True)
    pair = Pair(0.02, C_feat, mp, op)
    frames = list(orbit_frames(6, size, size, C_feat, S.S_TABLE))
    fetched, distinct = _host_frames(pair, frames[:2], 'sparse')
    for _, T, K, depth, feat in frames[2:4]:
        pair.depth(depth, T, K)
        pair.features(feat, T, K)
    f2, d2 = _host_frames(pair, frames[4:], 'sparse')
    pair.check_tsdf()
    assert pair.check_features(max_ulp=0) > 0
    assert fetched == distinct and f2 == d2
    torch.cuda.synchronize()


def test_host_frame_rejects_wrong_channels():
    import torch
    from nvblox_mindmap_b200._capi import NvbxError
    mp, op = make_params(workspace=S.WS_CUBE_STACKING)
    pair = Pair(0.02, 768, mp, op)
    K = S.intrinsics(64, 64)
    T = S.orbit_pose(0)
    # the Python surface rejects it (input contract, AssertionError like the device-frame integrators) ...
    with pytest.raises(AssertionError):
        pair.gpu.integrate_frame_from_host(torch.zeros(64, 64), torch.zeros(64, 64, 512, dtype=torch.float16),
                                           torch.from_numpy(T), torch.from_numpy(K))
    # ... and so does the C ABI itself when called directly
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    d, f = torch.zeros(64, 64), torch.zeros(64, 64, 512, dtype=torch.float16)
    Tt = torch.from_numpy(T).contiguous()
    with pytest.raises(NvbxError):
        _capi.check(_capi.load().nvbx_integrate_frame_host(
            pair.gpu._handle, 0, d.data_ptr(), f.data_ptr(), 64, 64, 512, None, None, Tt.data_ptr(),
            float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), pair.gpu._stream()))
    # non-contiguous / wrong-dtype / wrong-size host buffers never reach the C side
    with pytest.raises(AssertionError):
        pair.gpu.integrate_frame_from_host(torch.zeros(64, 128)[:, ::2], torch.zeros(64, 64, 768, dtype=torch.float16),
                                           torch.from_numpy(T), torch.from_numpy(K))
    with pytest.raises(AssertionError):
        pair.gpu.integrate_frame_from_host(torch.zeros(64, 64), torch.zeros(64, 64, 768, dtype=torch.float16),
                                           torch.from_numpy(T), torch.from_numpy(K),
                                           depth_mask=torch.ones(32, 32, dtype=torch.uint8))
    with pytest.raises(AssertionError):
        pair.gpu.integrate_frame_from_host(torch.zeros(64, 64), torch.zeros(64, 64, 768, dtype=torch.float16),
                                           torch.from_numpy(T).double(), torch.from_numpy(K))
