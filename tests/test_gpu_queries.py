"""SURVEY 8(a) row a13 on the GPU: point queries (queryTSDFKernel / queryFeatureKernel,
NT/cpp/src/sdf_query.cu:206-270; tests NT/nvblox_torch/tests/test_query.py, test_layer.py::test_query_feature_layer)
and the one-pass get_all_blocks (py_layer.cpp:177-198), against the CPU oracle (orc_query_*)."""
import numpy as np
import pytest

from tests import scenes as S
from tests.parity_utils import Pair, make_params

pytestmark = pytest.mark.gpu


def _pair(voxel=0.02, C=32, n_frames=3, alpha=1.0):
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, alpha=alpha, strict=True)
    pair = Pair(voxel, C, mp, op)
    K = S.intrinsics(96, 96)
    for i in range(n_frames):
        T = S.orbit_pose(5 * i)
        pair.depth(S.render_depth(K, 96, 96, T, **S.S_TABLE), T, K)
        pair.features(S.feature_frame(96, 96, C, 300 + i), T, K)
    return pair


def _queries(pair, n=6000, seed=0):
    """Points inside allocated blocks (voxel centres, random positions, exact voxel / block boundaries and their float
    neighbours) and far outside the map."""
    rng = np.random.default_rng(seed)
    idx, _ = pair.cpu.all_blocks(0)
    bs = np.float32(pair.cpu.voxel_size * 8)
    pick = idx[rng.integers(0, len(idx), n)]
    inside = (pick.astype(np.float32) + rng.random((n, 3), dtype=np.float32)) * bs
    k = n // 4
    vox = rng.integers(0, 9, (k, 3)).astype(np.float32)
    edge = ((pick[:k].astype(np.float32) * np.float32(8) + vox) * np.float32(pair.cpu.voxel_size)).astype(np.float32)
    inside[:k] = edge
    inside[k:k + k // 2] = np.nextafter(edge[:k // 2], np.float32(10))
    inside[k + k // 2:2 * k] = np.nextafter(edge[k // 2:], np.float32(-10))
    outside = (rng.random((n // 4, 3), dtype=np.float32) * np.float32(50.0) + np.float32(20.0)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([inside, outside]).astype(np.float32))


def test_query_layer_matches_oracle():
    import torch
    from nvblox_torch.mapper import QueryType
    pair = _pair()
    q = _queries(pair)
    qd = torch.from_numpy(q).cuda()
    # TSDF: [N, 2] (distance, weight), misses keep the pre-fill 0
    got = pair.gpu.query_layer(QueryType.TSDF, qd, mapper_id=0).cpu().numpy()
    want = pair.cpu.query_tsdf(q)
    assert got.shape == want.shape == (len(q), 2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    hits = want[:, 1] != 0
    assert hits.sum() > 1000 and (~hits).sum() > 1000
    # FEATURE: [N, C + 1] fp16
    gotf = pair.gpu.query_layer(QueryType.FEATURE, qd, mapper_id=0).cpu().numpy()
    wantf = pair.cpu.query_features(q)
    assert gotf.shape == wantf.shape == (len(q), pair.C + 1)
    assert np.array_equal(gotf.view(np.uint16), wantf.view(np.uint16))
    assert (wantf[:, -1] != 0).sum() > 200


def test_query_layer_output_prefill_is_kept_for_misses():
    """`output=`: rows of positions outside every allocated block keep the caller's values (the kernel writes only
    on success, sdf_query.cu:229-238,264-269)."""
    import torch
    from nvblox_torch.mapper import QueryType
    pair = _pair(n_frames=1)
    q = _queries(pair, n=400, seed=1)
    qd = torch.from_numpy(q).cuda()
    out = torch.full((len(q), 2), 7.5, dtype=torch.float32, device='cuda')
    res = pair.gpu.query_layer(QueryType.TSDF, qd, output=out, mapper_id=0)
    assert res.data_ptr() == out.data_ptr()
    want = pair.cpu.query_tsdf(q)
    idx, _ = pair.cpu.all_blocks(0)
    blocks = set(map(tuple, idx.tolist()))
    bs = np.float32(pair.cpu.voxel_size * 8)
    in_block = np.asarray([tuple(np.floor(p / bs).astype(int)) in blocks for p in q])
    got = res.cpu().numpy()
    assert np.array_equal(got[in_block].view(np.uint32), want[in_block].view(np.uint32))
    assert (got[~in_block] == 7.5).all() and (~in_block).sum() > 50
    outf = torch.full((len(q), pair.C + 1), -3.0, dtype=torch.float16, device='cuda')
    gotf = pair.gpu.query_layer(QueryType.FEATURE, qd, output=outf, mapper_id=0).cpu().numpy()
    fidx, _ = pair.cpu.all_blocks(1)
    fblocks = set(map(tuple, fidx.tolist()))
    in_f = np.asarray([tuple(np.floor(p / bs).astype(int)) in fblocks for p in q])
    wantf = pair.cpu.query_features(q)
    assert np.array_equal(gotf[in_f].view(np.uint16), wantf[in_f].view(np.uint16))
    assert (gotf[~in_f] == -3.0).all()


def test_query_layer_error_paths():
    """NT/nvblox_torch/mapper.py:356-389: bad mapper ids assert; a query the native side rejects (CPU tensor, wrong
    dtype) raises ValueError; unsupported layers raise NotImplementedError."""
    import torch
    from nvblox_torch.mapper import QueryType
    pair = _pair(n_frames=1)
    q = torch.zeros((4, 3), dtype=torch.float32, device='cuda')
    with pytest.raises(AssertionError):
        pair.gpu.query_layer(QueryType.TSDF, q, mapper_id=5)
    with pytest.raises(ValueError):
        pair.gpu.query_layer(QueryType.TSDF, q.cpu(), mapper_id=0)
    with pytest.raises(ValueError):
        pair.gpu.query_layer(QueryType.FEATURE, q.double(), mapper_id=0)
    with pytest.raises(AssertionError):
        pair.gpu.query_layer(QueryType.FEATURE, q, mapper_id=-1)
    with pytest.raises(NotImplementedError):
        pair.gpu.query_layer(QueryType.ESDF, q, mapper_id=0)
    assert pair.gpu.query_layer(QueryType.TSDF, q[:0], mapper_id=0).shape == (0, 2)


def test_get_all_blocks_single_pass():
    """get_all_blocks returns every block of the layer (one launch, one sync) and the views are the same memory
    get_block_at_index hands out."""
    from nvblox_mindmap_b200 import _capi
    pair = _pair(n_frames=2)
    L = _capi.load()
    pair.gpu.pipeline_join()   # (NVBX_PIPELINING=2: frames still queued would be issued -- and counted -- inside the call)
    for view, layer_id in ((pair.gpu.tsdf_layer_view(0), 0), (pair.gpu.feature_layer_view(0), 1)):
        before = int(L.nvbx_kernel_launch_count())
        blocks, indices = view.get_all_blocks()
        launches = int(L.nvbx_kernel_launch_count()) - before
        # count pass + (TSDF views only: clear of the free-space flags) + collect pass, independent of the block count
        assert launches <= 3, f'get_all_blocks launched {launches} kernels for {len(blocks)} blocks'
        ci, cd = pair.cpu.all_blocks(layer_id)
        got = np.stack([i.numpy() for i in indices])
        order = np.lexsort((got[:, 2], got[:, 1], got[:, 0]))
        assert np.array_equal(got[order], ci)
        for k in order[:8]:
            one = view.get_block_at_index(indices[k])
            assert one.data_ptr() == blocks[k].data_ptr()
            want = cd[np.flatnonzero((ci == got[k]).all(1))[0]]
            g = blocks[k].contiguous().cpu().numpy()
            assert np.array_equal(g.view(np.uint8), want.view(np.uint8))
