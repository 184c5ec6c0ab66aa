"""Multi-GPU partitioning of the hot path: independent map replicas, no collective.

A single map does not shard (every stage is keyed by one spatial hash and reads neighbours); across maps
(environments / episodes in datagen and batched open-loop eval) the work is embarrassingly parallel.
Map `m` lives on rank `m mod world_size`; the only cross-rank traffic is the timing barrier and, when a
caller wants every point cloud in one place, a variable-length gather of the exported clouds.
"""
from typing import List, Sequence


def maps_of_rank(n_maps: int, world_size: int, rank: int) -> List[int]:
    """Map ids owned by `rank` (round-robin: map m -> rank m mod world_size)."""
    assert world_size >= 1 and 0 <= rank < world_size
    return list(range(rank, n_maps, world_size))


def owner_of_map(map_id: int, world_size: int) -> int:
    return map_id % world_size


def aggregate_throughput(units_per_rank: Sequence[float], seconds_per_rank: Sequence[float]) -> float:
    """Whole-job throughput: all units processed by all ranks over the SLOWEST rank's time."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)


def gather_clouds(vertices, features, group=None):
    """Optional: gather every rank's exported (vertices [N_i,3], features [N_i,C]) on all ranks with one
    variable-length all-gather (works with gloo on CPU tensors and nccl on CUDA tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([vertices.shape[0]], dtype=torch.int64, device=vertices.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes) if sizes else 0
    pad_v = torch.zeros((nmax, 3), dtype=vertices.dtype, device=vertices.device)
    pad_f = torch.zeros((nmax, features.shape[1]), dtype=features.dtype, device=features.device)
    pad_v[:vertices.shape[0]] = vertices
    pad_f[:features.shape[0]] = features
    out_v = [torch.zeros_like(pad_v) for _ in range(world)]
    out_f = [torch.zeros_like(pad_f) for _ in range(world)]
    dist.all_gather(out_v, pad_v, group=group)
    dist.all_gather(out_f, pad_f, group=group)
    return [v[:s] for v, s in zip(out_v, sizes)], [f[:s] for f, s in zip(out_f, sizes)]
