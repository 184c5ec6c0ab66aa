"""BASELINE configs[3] (batched datagen): several independent maps on one GPU through MapBatch ->
nvbx_integrate_frames_batch (one stream per map, launches issued by a pool of host threads).  Every map must end up
bit-identical to the oracle fed with that map's frames, i.e. to the one-call-per-frame path."""
import numpy as np
import pytest

from tests import scenes as S
from tests.parity_utils import gpu_blocks, make_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('host_threads', [1, 0])
def test_map_batch_equals_oracle_per_map(host_threads):
    import torch
    from nvblox_mindmap_b200.replicas import MapBatch
    from nvblox_torch.constants import constants
    from oracle import oracle as O
    C_FEAT, HW, n_maps = 64, 96, 5
    constants.set_feature_array_num_elements(C_FEAT)
    mp, op = make_params(workspace=S.WS_CUBE_STACKING, strict=True)
    batch = MapBatch(n_maps, 0.02, mp, host_threads=host_threads)
    cpu = [O.OracleMapper(0.02, C_FEAT, op) for _ in range(n_maps)]
    K = S.intrinsics(HW, HW)
    K_t = torch.from_numpy(K)
    for step in range(3):
        depths, feats, poses = [], [], []
        for k in range(n_maps):                      # per-map seed / phase, as in datagen
            T = S.orbit_pose(step + 5 * k)
            d = S.render_depth(K, HW, HW, T, **(S.S_TABLE if k % 2 else S.S_SPHERE_SMALL))
            f = S.feature_frame(HW, HW, C_FEAT, 3000 + 10 * k + step)
            cpu[k].add_depth_frame(d, T, K)
            if not (k == 2 and step == 1):           # one map skips a feature frame (depth-only job)
                cpu[k].add_feature_frame(f, T, K)
                feats.append(torch.from_numpy(f).cuda())
            else:
                feats.append(None)
            depths.append(torch.from_numpy(d).cuda())
            poses.append(torch.from_numpy(T))
        batch.wait_for_current_stream()
        batch.integrate_frames(depths, feats, poses, K_t)
    batch.join_current_stream()
    torch.cuda.synchronize()
    for k in range(n_maps):
        gi, gd = gpu_blocks(batch.mappers[k].tsdf_layer_view(0))
        ci, cd = cpu[k].all_blocks(0)
        assert np.array_equal(gi, ci) and len(gi) > 0, k
        assert np.array_equal(gd.view(np.uint32), cd.view(np.uint32)), k
        gi, gd = gpu_blocks(batch.mappers[k].feature_layer_view(0))
        ci, cd = cpu[k].all_blocks(1)
        assert np.array_equal(gi, ci) and len(gi) > 0, k
        assert np.array_equal(gd.view(np.uint16), cd.view(np.uint16)), k
        assert batch.mappers[k].counters(0)['feature_voxels_updated'] == cpu[k].counters()['feature_voxels_updated']


def test_map_batch_reports_the_failing_job():
    import torch
    from nvblox_mindmap_b200 import _capi
    from nvblox_mindmap_b200.replicas import MapBatch
    from nvblox_torch.constants import constants
    constants.set_feature_array_num_elements(32)
    mp, _ = make_params(workspace=S.WS_CUBE_STACKING)
    batch = MapBatch(3, 0.02, mp)
    K = S.intrinsics(64, 64)
    T = S.orbit_pose(0)
    d = torch.from_numpy(S.render_depth(K, 64, 64, T, **S.S_TABLE)).cuda()
    good = torch.from_numpy(S.feature_frame(64, 64, 32, 1)).cuda()
    bad = torch.from_numpy(S.feature_frame(64, 64, 16, 2)).cuda()       # wrong channel count
    with pytest.raises(_capi.NvbxError, match='job 1'):
        batch.integrate_frames([d, d, d], [good, bad, good], [torch.from_numpy(T)] * 3, torch.from_numpy(K))
    torch.cuda.synchronize()
    statuses = [batch._jobs[k].status for k in range(3)]
    assert statuses[0] == 0 and statuses[1] < 0 and statuses[2] == 0
    for k in (0, 2):                                 # the other maps were integrated
        assert batch.mappers[k].counters(0)['feature_voxels_updated'] > 0
