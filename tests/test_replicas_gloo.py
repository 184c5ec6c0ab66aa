"""World-size-2 CPU test of the N>1 path: maps are sharded round-robin over ranks with NO data-path
collective; the only exchanges are the timing reduction (MAX over ranks) and the optional cloud gather.
Each rank integrates its own maps with the CPU oracle standing in for the GPU (test infrastructure)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_maps, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from nvblox_mindmap_b200.replicas import gather_clouds, maps_of_rank
    from oracle import oracle as O
    from tests import scenes as S
    from tests.parity_utils import make_params
    O.set_threads(1)
    K = S.intrinsics(48, 48)
    mine = maps_of_rank(n_maps, world, rank)
    verts, feats, frames = [], [], 0
    for map_id in mine:
        _, p = make_params(workspace=S.WS_CUBE_STACKING)
        m = O.OracleMapper(0.04, 8, p)
        for i in range(2):
            T = S.orbit_pose(i + map_id)           # per-map sequence
            m.add_depth_frame(S.render_depth(K, 48, 48, T, **S.S_TABLE), T, K)
            m.add_feature_frame(S.feature_frame(48, 48, 8, 1000 * map_id + i), T, K)   # per-map seed
            frames += 1
        m.update_feature_mesh()
        v, f, _ = m.get_feature_mesh()
        verts.append(v)
        feats.append(f)
    v = torch.from_numpy(np.concatenate(verts))
    f = torch.from_numpy(np.concatenate(feats).astype(np.float32))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)         # pretend timings: max over ranks is used
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([frames], dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    all_v, all_f = gather_clouds(v, f)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'r.npz'), tmax=t.item(), frames=n.item(), sizes=[len(x) for x in all_v],
                 feat_sizes=[len(x) for x in all_f], own=len(v))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_replicas(tmp_path):
    world, n_maps = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), n_maps, str(tmp_path)), nprocs=world, join=True)
    r = np.load(os.path.join(str(tmp_path), 'r.npz'))
    assert r['tmax'] == 2.0                   # slowest rank's time
    assert r['frames'] == n_maps * 2          # every map integrated exactly once
    assert len(r['sizes']) == world and all(s > 0 for s in r['sizes'])
    assert list(r['sizes']) == list(r['feat_sizes']) and r['sizes'][0] == r['own']
