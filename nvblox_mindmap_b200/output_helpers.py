"""Drop-in for mindmap's export post-processing, fused on the device.

Mirrors `mindmap/mapping/helpers/nvblox_output_helpers.py:22-89` (get_vertices_and_features) and
`mindmap/data_loading/vertex_sampling.py:17-176` (VertexSamplingMethod, sample_to_n_vertices): same names,
argument meaning, return shapes / dtypes and error behaviour, so `isaaclab_nvblox_mapper.py:207-250` can import
these instead.  The reference runs ~8 torch passes over the [N, C] cloud; here the cloud is filtered by
`Mapper.export_points` (flag / scan / scatter kernels) and sampled + zero-padded (+ optionally cast to float32) by
`Mapper.gather_points` (one kernel).  Random index lists are drawn on the host with the SAME torch calls the
reference makes (`torch.randperm(n)[:k]`, `torch.randint(0, n, (k,))`), so with the same torch seed the sampled
cloud is identical to the reference's.
"""
from enum import Enum
from typing import Optional, Tuple

import numpy as np
import torch


class VertexSamplingMethod(Enum):
    """vertex_sampling.py:17-26."""
    RANDOM_WITHOUT_REPLACEMENT = 'random_without_replacement'
    RANDOM_WITH_REPLACEMENT = 'random_with_replacement'
    LOWEST = 'lowest'
    NONE = 'none'


def _method_value(method) -> str:
    return method.value if isinstance(method, Enum) else str(method)


def _sample_exported(mapper, mapper_id: int, vertices: torch.Tensor, features: torch.Tensor, desired: int, method,
                     seed: Optional[int], features_dtype: Optional[torch.dtype]):
    """sample_to_n_vertices over the mapper's last export (vertices / features are its zero-copy views)."""
    n, channels = features.shape[0], features.shape[1]
    m = _method_value(method)
    if m == 'none' or n == desired:
        valid = torch.ones(n, device=vertices.device, dtype=torch.bool)
        if features_dtype is not None and features_dtype != features.dtype:
            vertices, features = mapper.gather_points(mapper_id, None, n, channels, features_dtype, n_rows=n)
        return vertices, features, valid
    if n > desired:
        valid = torch.ones(desired, device=vertices.device, dtype=torch.bool)
        if m == 'random_without_replacement':
            if seed is not None:
                torch.manual_seed(seed)
            idx = torch.randperm(n)[:desired]
        elif m == 'random_with_replacement':
            if seed is not None:
                torch.manual_seed(seed)
            idx = torch.randint(0, n, (desired,))
        elif m == 'lowest':
            # select_n_lowest_z_vertices (:111-126): np.argsort(-z)[:k] -- the same numpy call on the z column
            idx = torch.from_numpy(np.argsort(-vertices[:, 2].cpu().numpy())[:desired].astype(np.int64))
        else:
            raise ValueError(f'Vertex sampling method {method} is not yet implemented.')
        out_v, out_f = mapper.gather_points(mapper_id, idx, desired, channels, features_dtype or torch.float16)
        return out_v, out_f, valid
    # pad_with_zeros (:84-108): torch.cat of the fp16 features with float32 zeros promotes to float32
    out_v, out_f = mapper.gather_points(mapper_id, None, desired, channels, torch.float32, n_rows=n)
    valid = torch.ones(desired, device=vertices.device, dtype=torch.bool)
    valid[n:] = False
    return out_v, out_f, valid


def _owned(tensors, zero_copy: bool):
    """The reference returns tensors that own their memory; ours come out of arenas the next export / mesh update
    recycles.  Unless the caller opts into the zero-copy views, hand out copies."""
    return tensors if zero_copy else tuple(t.clone() for t in tensors)


def sample_to_n_vertices(vertices: torch.Tensor, features: torch.Tensor, desired_num_vertices: int, method,
                         seed: Optional[int] = None, mapper=None, mapper_id: int = 0, zero_copy: bool = False
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """vertex_sampling.py:29-81.  fp16 CUDA rows are staged through the mapper's export arena (identity filter; skipped
    when they already ARE the mapper's last export) and sampled by one gather kernel; anything else takes the
    reference's torch path and keeps its dtype.  Returns tensors that own their memory unless `zero_copy=True` (then
    they are views of the mapper's arenas, valid until its next export / gather / mesh update)."""
    assert vertices.dim() == 2
    assert features.dim() == 2
    assert vertices.shape[0] == features.shape[0]
    if features.dtype != torch.float16 or not vertices.is_cuda:
        # The device path moves fp16 rows; any other input (the reference keeps whatever dtype it is given) goes through
        # the same torch calls the reference makes, so nothing is down-cast behind the caller's back.
        return _sample_with_torch(vertices, features, desired_num_vertices, method, seed)
    if mapper is None:
        raise ValueError('sample_to_n_vertices needs the mapper whose export arena holds the points')
    last = getattr(mapper, '_last_export', {}).get(mapper_id)
    if last is not None and last == (vertices.data_ptr(), features.data_ptr(), vertices.shape[0], features.shape[1]):
        ev, ef = vertices, features         # already the arena's own views (export_points output): nothing to stage
    else:
        big = 3.0e38
        ev, ef = mapper.export_points(mapper_id, (-big,) * 3, (big,) * 3, 0, False, vertices=vertices, features=features)
    return _owned(_sample_exported(mapper, mapper_id, ev, ef, desired_num_vertices, method, seed, None), zero_copy)


def _sample_with_torch(vertices, features, desired, method, seed):
    """vertex_sampling.py:29-176 with plain torch indexing (inputs that are not fp16 CUDA rows)."""
    n = features.shape[0]
    m = _method_value(method)
    if m == 'none' or n == desired:
        return vertices, features, torch.ones(n, device=vertices.device, dtype=torch.bool)
    if n > desired:
        if m == 'random_without_replacement':
            if seed is not None:
                torch.manual_seed(seed)
            idx = torch.randperm(n)[:desired]
        elif m == 'random_with_replacement':
            if seed is not None:
                torch.manual_seed(seed)
            idx = torch.randint(0, n, (desired,))
        elif m == 'lowest':
            idx = torch.from_numpy(np.argsort(-vertices[:, 2].cpu().numpy())[:desired].astype(np.int64))
        else:
            raise ValueError(f'Vertex sampling method {method} is not yet implemented.')
        idx = idx.to(vertices.device)
        return vertices[idx, :], features[idx, :], torch.ones(desired, device=vertices.device, dtype=torch.bool)
    pad = desired - n
    features = torch.cat([features, torch.zeros((pad, features.shape[1]), device=features.device)], dim=0)
    vertices = torch.cat([vertices, torch.zeros((pad, vertices.shape[1]), device=vertices.device)], dim=0)
    valid = torch.ones(desired, device=vertices.device, dtype=torch.bool)
    valid[n:] = False
    return vertices, features, valid


def get_vertices_and_features(mapper, mapper_id: int, nvblox_mapping_config, remove_zero_features: bool,
                              num_excess_features: int, sample_vertices: bool,
                              number_of_vertices_to_sample: Optional[int] = None, vertex_sampling_method=None,
                              seed: Optional[int] = None, features_dtype: Optional[torch.dtype] = None,
                              zero_copy: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """nvblox_output_helpers.py:22-89.  Returns (vertices, features, valid_mask) with the reference's shapes: a
    leading batch dimension of 1 on all three when sampling, on the mask only otherwise.

    `features_dtype=torch.float32` (ours) additionally folds the cast of isaaclab_nvblox_mapper.py:243-246 into
    the gather pass.  The returned tensors OWN their memory, as the reference's do (the caller keeps them across
    later steps); `zero_copy=True` (ours) returns views of the mapper's export arenas instead, valid until the
    mapper's next export / gather / mesh update."""
    mapper.update_feature_mesh(mapper_id)
    mesh = mapper.get_feature_mesh(mapper_id)
    assert mesh.vertices().shape[0] == mesh.vertex_features().shape[0]
    assert mesh.vertices().shape[0] != 0, 'No vertices found in the mesh.'
    lo = [float(v) for v in nvblox_mapping_config.aabb_min_m]
    hi = [float(v) for v in nvblox_mapping_config.aabb_max_m]
    vertices, features = mapper.export_points(mapper_id, lo, hi, num_excess_features, remove_zero_features)
    if not sample_vertices:
        if features_dtype is not None and features_dtype != features.dtype:
            vertices, features = mapper.gather_points(mapper_id, None, vertices.shape[0], features.shape[1],
                                                      features_dtype, n_rows=vertices.shape[0])
        valid_mask = torch.ones(vertices.shape[0], dtype=torch.bool, device=vertices.device).unsqueeze(0)
        return _owned((vertices, features, valid_mask), zero_copy)
    vertices, features, valid_mask = _sample_exported(mapper, mapper_id, vertices, features,
                                                      number_of_vertices_to_sample, vertex_sampling_method, seed,
                                                      features_dtype)
    if valid_mask.ndim == 1:
        vertices = vertices.unsqueeze(0)
        features = features.unsqueeze(0)
        valid_mask = valid_mask.unsqueeze(0)
    return _owned((vertices, features, valid_mask), zero_copy)
