"""Layer views (reference: nvblox_torch/layer.py, cpp/src/py_layer.cpp:24-198).

A layer view shares the mapper's device-resident map.  Block tensors are zero-copy views:
TSDF [8,8,8,2] float32 (distance, weight), colour [8,8,8,3] uint8, feature [8,8,8,C+1] float16 (C features, weight) -- the
feature view is STRIDED (voxel rows are padded to C+8 halves so they stay 16-byte aligned); call
.contiguous() if a dense copy is needed.  Views are invalidated by any call that mutates the map.
"""
import abc
import ctypes as C
from typing import Callable, List, Optional, Tuple, Union

import numpy as np
import torch

from nvblox_mindmap_b200 import _capi
from nvblox_mindmap_b200.torch_interop import current_stream_ptr, device_view
from nvblox_torch import indexing
from nvblox_torch.constants import constants

_TSDF, _FEATURE, _COLOR = 0, 1, 2


class Layer(abc.ABC):
    """Base class of the voxel-block layer views."""

    block_dim_in_voxels = 8
    _layer_id = _TSDF

    def __init__(self, voxel_size_m: float, torch_class_name: str = '', c_layer=None):
        if c_layer is None:
            # A free-standing layer = a private single-map mapper (reference: new native layer).
            from nvblox_torch.mapper import Mapper
            c_layer = (Mapper(voxel_sizes_m=float(voxel_size_m)), 0)
        self._mapper, self._map_id = c_layer
        self._c_layer = c_layer

    # -- plumbing ---------------------------------------------------------------------------------
    def _h(self):
        return self._mapper._handle

    def _stream(self):
        return current_stream_ptr(self._mapper._device)

    @staticmethod
    def _xyz(index) -> Tuple[int, int, int]:
        v = index.tolist() if hasattr(index, 'tolist') else list(index)
        return int(v[0]), int(v[1]), int(v[2])

    @staticmethod
    @abc.abstractmethod
    def num_elements_per_voxel() -> int:
        """Number of elements per voxel in the block tensors."""

    # -- reference API ------------------------------------------------------------------------------
    def voxel_size(self) -> float:
        return float(_capi.load().nvbx_voxel_size(self._h(), self._map_id))

    def num_blocks(self) -> int:
        return int(_capi.check(_capi.load().nvbx_num_blocks(self._h(), self._map_id, self._layer_id, self._stream())))

    def num_allocated_blocks(self) -> int:
        return int(_capi.check(_capi.load().nvbx_num_allocated_blocks(self._h(), self._map_id, self._layer_id,
                                                                      self._stream())))

    def num_allocated_bytes(self) -> int:
        return int(_capi.check(_capi.load().nvbx_num_allocated_bytes(self._h(), self._map_id, self._layer_id,
                                                                     self._stream())))

    def clear(self) -> None:
        """Clear the layer.  Layers share one block table here, so this clears the whole map."""
        self._mapper.clear(self._map_id)

    def allocate_block_at_index(self, index: torch.Tensor) -> None:
        x, y, z = self._xyz(index)
        _capi.check(_capi.load().nvbx_allocate_block(self._h(), self._map_id, self._layer_id, x, y, z, self._stream()))

    def _block_view(self, x: int, y: int, z: int) -> Optional[torch.Tensor]:
        ptr, stride = C.c_void_p(), C.c_int64()
        rc = _capi.load().nvbx_get_block_ptr(self._h(), self._map_id, self._layer_id, x, y, z, C.byref(ptr),
                                             C.byref(stride), self._stream())
        if rc == -4:    # NVBX_ERR_NOT_FOUND
            return None
        _capi.check(rc)
        n, s = self.num_elements_per_voxel(), int(stride.value)
        dtype = {_TSDF: torch.float32, _FEATURE: torch.float16, _COLOR: torch.uint8}[self._layer_id]
        return device_view(ptr.value, (8, 8, 8, n), dtype, self._mapper._device,
                           strides_elems=(64 * s, 8 * s, s, 1), owner=self._mapper)

    def is_block_allocated(self, index: torch.Tensor) -> bool:
        return self._block_view(*self._xyz(index)) is not None

    def get_block_at_index(self, index: torch.Tensor) -> Optional[torch.Tensor]:
        """Zero-copy [8,8,8,E] view of the block, or None if it is not allocated."""
        return self._block_view(*self._xyz(index))

    def get_all_block_indices(self) -> torch.Tensor:
        """[N,3] int32 CPU tensor of allocated block indices."""
        L = _capi.load()
        n = int(_capi.check(L.nvbx_get_block_indices(self._h(), self._map_id, self._layer_id, None, 0, self._stream())))
        out = np.zeros((max(n, 1), 3), np.int32)
        n2 = int(_capi.check(L.nvbx_get_block_indices(self._h(), self._map_id, self._layer_id,
                                                      out.ctypes.data_as(C.c_void_p), n, self._stream())))
        return torch.from_numpy(out[:min(n, n2)].copy())

    def get_all_blocks(self) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """(block tensors, block indices): zero-copy device views + int32[3] CPU index tensors.

        One kernel launch and one stream synchronisation for the whole layer (reference: one pass over the
        layer, py_layer.cpp:177-198), not one per block."""
        L = _capi.load()
        stride = C.c_int64()
        n = int(_capi.check(L.nvbx_get_all_blocks(self._h(), self._map_id, self._layer_id, None, None, 0,
                                                  C.byref(stride), self._stream())))
        if n == 0:
            return [], []
        idx = np.zeros((n, 3), np.int32)
        ptrs = np.zeros(n, np.uint64)
        n2 = int(_capi.check(L.nvbx_get_all_blocks(self._h(), self._map_id, self._layer_id,
                                                   idx.ctypes.data_as(C.c_void_p), ptrs.ctypes.data_as(C.c_void_p), n,
                                                   C.byref(stride), self._stream())))
        n = min(n, n2)
        ne, s = self.num_elements_per_voxel(), int(stride.value)
        dtype = {_TSDF: torch.float32, _FEATURE: torch.float16, _COLOR: torch.uint8}[self._layer_id]
        idx_t = torch.from_numpy(idx[:n].copy())
        blocks = [device_view(int(ptrs[k]), (8, 8, 8, ne), dtype, self._mapper._device,
                              strides_elems=(64 * s, 8 * s, s, 1), owner=self._mapper) for k in range(n)]
        return blocks, [idx_t[k] for k in range(n)]

    def get_block_limits(self) -> Tuple[torch.Tensor, torch.Tensor]:
        idx = self.get_all_block_indices()
        return torch.min(idx, dim=0)[0], torch.max(idx, dim=0)[0]

    def get_voxels_matching_condition(self, get_voxel_mask: Callable) -> Tuple[torch.Tensor, torch.Tensor]:
        """Values [N,E] and centres [N,3] of the voxels for which `get_voxel_mask(block)` is true."""
        blocks, indices = self.get_all_blocks()
        dev = f'cuda:{self._mapper._device}'
        centers = indexing.get_voxel_center_grids(indices, self.voxel_size(), device=dev)
        pts = [torch.zeros((0, 3), device=dev)]
        vals = [torch.zeros((0, self.num_elements_per_voxel()), device=dev)]
        for blk, ctr in zip(blocks, centers):
            mask = get_voxel_mask(blk)
            assert mask.shape == torch.Size([8, 8, 8]), 'Your condition should generate a 8x8x8 mask.'
            pts.append(ctr[mask, :])
            vals.append(blk[mask].to(vals[0].dtype))
        return torch.vstack(vals), torch.vstack(pts)


class TsdfLayer(Layer):
    """TSDF layer view: voxels are (distance, weight) float32."""
    _layer_id = _TSDF

    def __init__(self, voxel_size_m: float, c_layer=None):
        super().__init__(voxel_size_m, 'TsdfLayer', c_layer)

    @staticmethod
    def num_elements_per_voxel() -> int:
        return 2

    def get_tsdf_mask_negative_distance(self, tsdf_block: torch.Tensor) -> torch.Tensor:
        return torch.logical_and(tsdf_block[..., 0] < 0.0, tsdf_block[..., 1] > 0.01)

    def get_tsdfs_below_zero(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.get_voxels_matching_condition(self.get_tsdf_mask_negative_distance)


class FeatureLayer(Layer):
    """Feature layer view: voxels are C fp16 features followed by an fp16 weight."""
    _layer_id = _FEATURE

    def __init__(self, voxel_size_m: float, c_layer=None):
        super().__init__(voxel_size_m, 'FeatureLayer', c_layer)

    @staticmethod
    def num_elements_per_voxel() -> int:
        return constants.feature_array_num_elements() + 1


class ColorLayer(Layer):
    """Colour layer view: [8,8,8,3] uint8 RGB, striding over the per-voxel weight exactly like the reference's
    tensorFromBlock(ColorBlock*) (py_layer.cpp:49-70: 8-byte ColorVoxel = r, g, b, pad, float weight)."""
    _layer_id = _COLOR

    def __init__(self, voxel_size_m: float, c_layer=None):
        super().__init__(voxel_size_m, 'ColorLayer', c_layer)

    @staticmethod
    def num_elements_per_voxel() -> int:
        return 3


class OccupancyLayer(Layer):
    """Not on this path."""

    @staticmethod
    def num_elements_per_voxel() -> int:
        return 1


class EsdfLayer(Layer):
    """Not on this path."""

    @staticmethod
    def num_elements_per_voxel() -> int:
        return 4


def convert_layer_to_dense_tensor(layer: Union[TsdfLayer, FeatureLayer],
                                  unobserved_value: float = 0.0,
                                  aabb_min_m: Optional[torch.Tensor] = None,
                                  aabb_max_m: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Dense (X, Y, Z, E) float32 grid of a sparse layer's values + (X, Y, Z, 3) voxel centres.

    Same contract as the reference's pure-torch helper (nvblox_torch/layer.py:254-354): the AABB is
    inclusive block-wise, unobserved voxels take `unobserved_value`, weights are dropped.
    """
    if aabb_min_m is None or aabb_max_m is None:
        bmin, bmax = layer.get_block_limits()
    else:
        bmin = torch.floor(aabb_min_m / layer.block_dim_in_voxels / layer.voxel_size()).to(torch.int)
        bmax = torch.ceil(aabb_max_m / layer.block_dim_in_voxels / layer.voxel_size()).to(torch.int)
    bmin, bmax = bmin.cpu(), bmax.cpu()
    nblk = bmax - bmin + 1
    if isinstance(layer, TsdfLayer):
        depth = 1
    elif isinstance(layer, FeatureLayer):
        depth = layer.num_elements_per_voxel() - 1
    else:
        raise TypeError(f'Unsupported layer type to convert to dense tensor: {type(layer)}')
    n = layer.block_dim_in_voxels
    dev = f'cuda:{layer._mapper._device}'
    out = torch.full((nblk * n).tolist() + [depth], fill_value=unobserved_value, dtype=torch.float32, device=dev)
    for row in layer.get_all_block_indices():
        rel = row - bmin
        if bool((rel < 0).any()) or bool((rel >= nblk).any()):
            continue
        blk = layer.get_block_at_index(row)
        if blk is None:
            continue
        x, y, z = (rel * n).tolist()
        out[x:x + n, y:y + n, z:z + n] = blk[..., :depth].to(torch.float32)
    lo, hi = bmin * n, (bmax + 1) * n
    grids = torch.meshgrid(*[torch.arange(int(lo[i]), int(hi[i]), device=dev) for i in range(3)], indexing='ij')
    centers = (torch.stack(grids, dim=-1) + 0.5) * layer.voxel_size()
    assert out.shape[:-1] == centers.shape[:-1]
    return out, centers
