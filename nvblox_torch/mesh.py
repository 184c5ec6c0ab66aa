"""Mesh holders (reference: nvblox_torch/mesh.py, cpp/src/py_mesh.cpp:19-90).

A FeatureMesh is a snapshot descriptor of the mapper's serialised mesh arena: `vertices()` [N,3] f32,
`vertex_features()` [N,C] f16 and `triangles()` [M,3] i32 are zero-copy DEVICE views (the reference
returns views over pinned host memory labelled CUDA; py_mesh.cpp:19-28,67-90).  Views are valid until
the next update_feature_mesh / clear of that mapper.
"""
from typing import Any, Optional

import torch

from nvblox_torch.constants import constants


class Mesh:
    """Serialised mesh: vertices, per-vertex appearance, triangles."""

    def __init__(self, c_mesh: Optional[Any] = None) -> None:
        self._c_mesh = c_mesh if c_mesh is not None else self._create_empty_mesh()

    def _create_empty_mesh(self) -> Any:
        dev = 'cuda' if torch.cuda.is_available() else 'cpu'
        return {
            'vertices': torch.empty((0, 3), dtype=torch.float32, device=dev),
            'appearances': torch.empty((0, self._appearance_width()), dtype=self._appearance_dtype(), device=dev),
            'triangles': torch.empty((0, 3), dtype=torch.int32, device=dev),
        }

    def _appearance_width(self) -> int:
        return 3

    def _appearance_dtype(self) -> torch.dtype:
        return torch.uint8

    def vertices(self) -> torch.Tensor:
        """Vertices (N, 3) float32."""
        return self._c_mesh['vertices']

    def triangles(self) -> torch.Tensor:
        """Index triplets (M, 3) int32 into vertices()."""
        return self._c_mesh['triangles']

    def vertex_appearances(self) -> torch.Tensor:
        """Per-vertex appearance (N, F)."""
        return self._c_mesh['appearances']

    def __str__(self) -> str:
        return (f'Mesh(vertices={self.vertices().shape}, triangles={self.triangles().shape}, '
                f'vertex_appearances={self.vertex_appearances().shape})')


class FeatureMesh(Mesh):
    """Mesh whose vertices carry a C-channel fp16 feature (closest voxel of the feature layer)."""

    def _appearance_width(self) -> int:
        return constants.feature_array_num_elements()

    def _appearance_dtype(self) -> torch.dtype:
        return torch.float16

    def vertex_features(self) -> torch.Tensor:
        """Vertex features (N, C) float16."""
        return self.vertex_appearances()


class ColorMesh(Mesh):
    """Mesh whose vertices carry the uint8 RGB of the closest colour voxel (Gray where no colour block exists)."""

    def vertex_colors(self) -> torch.Tensor:
        """Vertex colours (N, 3) uint8."""
        return self.vertex_appearances()
