"""Voxel-centre grids (reference: nvblox_torch/indexing.py) -- pure torch."""
from typing import List

import torch

NUM_VOXELS_PER_SIDE = 8


def get_voxel_index_grid(device: torch.device = 'cuda') -> torch.Tensor:
    """[8,8,8,3] int32 grid of voxel indices inside a block."""
    r = torch.arange(NUM_VOXELS_PER_SIDE, device=device, dtype=torch.int32)
    return torch.stack(torch.meshgrid(r, r, r, indexing='ij'), dim=-1)


def get_local_voxel_center_grid(voxel_size: float, device: torch.device = 'cuda') -> torch.Tensor:
    """[8,8,8,3] float32 voxel centres relative to the block origin."""
    return (get_voxel_index_grid(device=device).to(torch.float32) + 0.5) * voxel_size


def get_voxel_center_grids(block_indices: List[torch.Tensor],
                           voxel_size: float,
                           device: torch.device = 'cuda') -> List[torch.Tensor]:
    """One [8,8,8,3] float32 grid of world-frame voxel centres per block index."""
    block_size = NUM_VOXELS_PER_SIDE * voxel_size
    local = get_local_voxel_center_grid(voxel_size, device=device)
    return [idx.to(torch.float32).to(device) * block_size + local for idx in block_indices]
