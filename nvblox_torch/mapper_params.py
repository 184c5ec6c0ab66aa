"""Parameter classes (reference: nvblox_torch/mapper_params.py + py_nvblox.cu:75-287).

Same class names, attribute names and MapperParams get_/set_ methods as the reference; the values are
plain Python attributes that `MapperParams.to_nvbx()` packs into the `nvbx_params` POD crossing the C ABI.
Defaults are the reference's (nvblox/integrators/projective_integrator_params.h:24-73,
view_calculator_params.h:23-59, tsdf_decay_integrator_params.h:22-38, mesh_integrator_params.h:21-26,
block_memory_pool_params).  Parameter groups that do not reach the hot path (ESDF, occupancy decay) are
accepted and ignored so that existing configuration code keeps working.
"""
import copy
from typing import Any, Dict

from nvblox_mindmap_b200.params import WEIGHTING_MODES, WORKSPACE_BOUNDS_TYPES, NvbxParams


class NvbloxParameterClass:
    """Base: attributes listed in `_defaults`; `get_<name>()` / `set_<name>(v)` work for each of them."""

    _defaults: Dict[str, Any] = {}

    def __init__(self, c_params: Any = None) -> None:
        values = dict(self._defaults)
        if isinstance(c_params, NvbloxParameterClass):
            values.update(c_params.__dict__)
        for k, v in values.items():
            object.__setattr__(self, k, v)

    @property
    def _c_params(self) -> 'NvbloxParameterClass':
        return self

    def __setattr__(self, name: str, value: Any) -> None:
        if name not in self._defaults:
            raise AttributeError(f'{type(self).__name__} has no parameter {name!r}')
        object.__setattr__(self, name, value)

    def __getattr__(self, name: str) -> Any:
        if name.startswith('get_') and name[4:] in self._defaults:
            return lambda: getattr(self, name[4:])
        if name.startswith('set_') and name[4:] in self._defaults:
            return lambda v: setattr(self, name[4:], v)
        raise AttributeError(name)

    def _method_names(self):
        return [f'get_{k}' for k in self._defaults] + [f'set_{k}' for k in self._defaults]

    def __repr__(self) -> str:
        return f'{type(self).__name__}({ {k: getattr(self, k) for k in self._defaults} })'


class ProjectiveIntegratorParams(NvbloxParameterClass):
    """Parameters governing the projective integrators."""
    _defaults = {
        'projective_integrator_max_integration_distance_m': 7.0,
        'lidar_projective_integrator_max_integration_distance_m': 10.0,
        'projective_integrator_truncation_distance_vox': 4.0,
        'projective_integrator_weighting_mode': 'kInverseSquareWeight',
        'projective_integrator_max_weight': 5.0,
        'projective_tsdf_integrator_invalid_depth_decay_factor': -1.0,
        'projective_appearance_integrator_measurement_weight': 0.8,
    }


class MeshIntegratorParams(NvbloxParameterClass):
    """Parameters governing the mesh integrator."""
    _defaults = {'mesh_integrator_min_weight': 1e-4, 'mesh_integrator_weld_vertices': True}


class DecayIntegratorBaseParams(NvbloxParameterClass):
    """Base parameters for the decay integrators."""
    _defaults = {'decay_integrator_deallocate_decayed_blocks': True}


class TsdfDecayIntegratorParams(NvbloxParameterClass):
    """Parameters governing the TSDF decay integrator."""
    _defaults = {
        'tsdf_decay_factor': 0.95,
        'tsdf_decayed_weight_threshold': 1e-3,
        'tsdf_set_free_distance_on_decayed': False,
        'tsdf_decayed_free_distance_vox': 4.0,
    }


class OccupancyDecayIntegratorParams(NvbloxParameterClass):
    """Accepted for compatibility; occupancy mapping is not on this path."""
    _defaults = {
        'free_region_decay_probability': 0.55,
        'occupied_region_decay_probability': 0.4,
        'occupancy_decay_to_free': False,
    }


class EsdfIntegratorParams(NvbloxParameterClass):
    """Accepted for compatibility; there is no ESDF layer on this path."""
    _defaults = {
        'esdf_integrator_max_distance_m': 2.0,
        'esdf_integrator_min_weight': 1e-4,
        'esdf_integrator_max_site_distance_vox': 1.0,
        'esdf_slice_min_height': 0.0,
        'esdf_slice_max_height': 1.0,
        'esdf_slice_height': 1.0,
        'slice_height_above_plane_m': 0.0,
        'slice_height_thickness_m': 0.0,
    }


class ViewCalculatorParams(NvbloxParameterClass):
    """Parameters governing the view calculator."""
    _defaults = {
        'raycast_subsampling_factor': 4,
        'workspace_bounds_type': 'kUnbounded',
        'workspace_bounds_min_height_m': 0.0,
        'workspace_bounds_max_height_m': 1.0,
        'workspace_bounds_min_corner_x_m': 0.0,
        'workspace_bounds_max_corner_x_m': 0.0,
        'workspace_bounds_min_corner_y_m': 2.0,
        'workspace_bounds_max_corner_y_m': 2.0,
    }


class BlockMemoryPoolParams(NvbloxParameterClass):
    """Parameters governing memory allocation (only seeds our slab arenas)."""
    _defaults = {'num_preallocated_blocks': 2048, 'expansion_factor': 2.0}


_GROUPS = {
    'projective_integrator_params': ProjectiveIntegratorParams,
    'mesh_integrator_params': MeshIntegratorParams,
    'decay_integrator_base_params': DecayIntegratorBaseParams,
    'tsdf_decay_integrator_params': TsdfDecayIntegratorParams,
    'occupancy_decay_integrator_params': OccupancyDecayIntegratorParams,
    'esdf_integrator_params': EsdfIntegratorParams,
    'view_calculator_params': ViewCalculatorParams,
    'block_memory_pool_params': BlockMemoryPoolParams,
}


class MapperParams:
    """Aggregate of the parameter groups (reference MapperParams, mapper_params.h:52-67)."""

    def __init__(self, c_params: Any = None) -> None:
        src = c_params if isinstance(c_params, MapperParams) else None
        for name, cls in _GROUPS.items():
            setattr(self, '_' + name, cls(getattr(src, '_' + name)) if src is not None else cls())
        # (ours) see nvbx_params.strict_blend / appearance_truncation_distance_vox
        self.strict_blend = bool(getattr(src, 'strict_blend', False))
        self.appearance_truncation_distance_vox = float(getattr(src, 'appearance_truncation_distance_vox', 4.0))
        self.cache_last_viewpoint = bool(getattr(src, 'cache_last_viewpoint', True))

    @property
    def _c_params(self) -> 'MapperParams':
        return self

    def to_nvbx(self) -> NvbxParams:
        """Pack into the POD of include/nvbx_c_api.h."""
        from nvblox_mindmap_b200 import _capi
        p = _capi.default_params()
        pi, vc, td = self._projective_integrator_params, self._view_calculator_params, self._tsdf_decay_integrator_params
        p.max_integration_distance_m = pi.projective_integrator_max_integration_distance_m
        p.truncation_distance_vox = pi.projective_integrator_truncation_distance_vox
        # py_mapper_params.cpp:15-35: every unknown string falls into kInverseSquareTsdfDistancePenalty
        p.weighting_mode = WEIGHTING_MODES.get(pi.projective_integrator_weighting_mode, 4) \
            if pi.projective_integrator_weighting_mode != 'kLinearWithMax' else 4
        p.max_weight = pi.projective_integrator_max_weight
        p.invalid_depth_decay_factor = pi.projective_tsdf_integrator_invalid_depth_decay_factor
        p.appearance_measurement_weight = pi.projective_appearance_integrator_measurement_weight
        p.appearance_truncation_distance_vox = self.appearance_truncation_distance_vox
        p.tsdf_decay_factor = td.tsdf_decay_factor
        p.tsdf_decayed_weight_threshold = td.tsdf_decayed_weight_threshold
        p.tsdf_set_free_distance_on_decayed = int(bool(td.tsdf_set_free_distance_on_decayed))
        p.tsdf_decayed_free_distance_vox = td.tsdf_decayed_free_distance_vox
        p.deallocate_decayed_blocks = int(bool(
            self._decay_integrator_base_params.decay_integrator_deallocate_decayed_blocks))
        p.raycast_subsampling_factor = int(vc.raycast_subsampling_factor)
        if vc.workspace_bounds_type not in WORKSPACE_BOUNDS_TYPES:
            raise ValueError(f'Unrecognized workspace bound type: {vc.workspace_bounds_type}')
        p.workspace_bounds_type = WORKSPACE_BOUNDS_TYPES[vc.workspace_bounds_type]
        p.workspace_min[0] = vc.workspace_bounds_min_corner_x_m
        p.workspace_min[1] = vc.workspace_bounds_min_corner_y_m
        p.workspace_min[2] = vc.workspace_bounds_min_height_m
        p.workspace_max[0] = vc.workspace_bounds_max_corner_x_m
        p.workspace_max[1] = vc.workspace_bounds_max_corner_y_m
        p.workspace_max[2] = vc.workspace_bounds_max_height_m
        p.cache_last_viewpoint = int(self.cache_last_viewpoint)
        p.mesh_min_weight = self._mesh_integrator_params.mesh_integrator_min_weight
        p.mesh_weld_vertices = int(bool(self._mesh_integrator_params.mesh_integrator_weld_vertices))
        p.num_preallocated_blocks = int(self._block_memory_pool_params.num_preallocated_blocks)
        p.expansion_factor = float(self._block_memory_pool_params.expansion_factor)
        p.strict_blend = int(self.strict_blend)
        return p

    def copy(self) -> 'MapperParams':
        return copy.deepcopy(self)


def _add_group_accessors() -> None:
    for name, cls in _GROUPS.items():

        def getter(self, _n=name, _c=cls):
            return _c(getattr(self, '_' + _n))

        def setter(self, params, _n=name, _c=cls):
            assert isinstance(params, _c), f'expected {_c.__name__}'
            setattr(self, '_' + _n, _c(params))

        setattr(MapperParams, 'get_' + name, getter)
        setattr(MapperParams, 'set_' + name, setter)


_add_group_accessors()
