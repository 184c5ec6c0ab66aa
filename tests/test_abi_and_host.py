"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol,
the POD layout agrees between Python and C, and the host-side mirror of the reference interface behaves
like the reference's (assertions, parameter plumbing).  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from nvblox_mindmap_b200 import build
    return C.CDLL(build.build())


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'nvbx_c_api.h')).read()
    declared = set(re.findall(r'\b(nvbx_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    lib = _lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from nvblox_mindmap_b200._capi import API_SYMBOLS
    assert set(API_SYMBOLS) == declared


def test_default_params_match_reference_defaults_and_oracle():
    from nvblox_mindmap_b200.params import NvbxParams
    from oracle import oracle as O
    lib = _lib()
    p = NvbxParams()
    lib.nvbx_default_params.argtypes = [C.POINTER(NvbxParams)]
    lib.nvbx_default_params(C.byref(p))
    q = O.default_params()
    assert bytes(p) == bytes(q)    # same POD, byte for byte
    assert p.max_integration_distance_m == 7.0 and p.truncation_distance_vox == 4.0
    assert p.weighting_mode == 2 and p.max_weight == 5.0 and p.invalid_depth_decay_factor == -1.0
    assert abs(p.appearance_measurement_weight - 0.8) < 1e-7 and abs(p.tsdf_decay_factor - 0.95) < 1e-7
    assert p.raycast_subsampling_factor == 4 and p.sphere_tracing_subsampling == 4
    assert p.sphere_tracing_max_ray_length_m == 7.0 and p.sphere_tracing_max_steps == 100
    assert p.mesh_weld_vertices == 1 and abs(p.mesh_min_weight - 1e-4) < 1e-10
    assert list(p.workspace_min) == [0.0, 2.0, 0.0] and list(p.workspace_max) == [0.0, 2.0, 1.0]


def test_error_reporting_without_gpu():
    lib = _lib()
    lib.nvbx_last_error.restype = C.c_char_p
    h = C.c_void_p()
    vs = (C.c_float * 1)(0.02)
    rc = lib.nvbx_create(1, vs, None, 12, 0, C.byref(h))    # 12 is not a multiple of 8
    assert rc == -1 and b'multiple of 8' in lib.nvbx_last_error()
    if not torch.cuda.is_available():
        rc = lib.nvbx_create(1, vs, None, 768, 0, C.byref(h))
        assert rc < 0 and b'no CPU fallback' in lib.nvbx_last_error()
        assert not h.value


def test_mapper_raises_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from nvblox_torch.mapper import Mapper
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Mapper(voxel_sizes_m=0.02)


def test_reference_module_surface():
    """Every import mindmap/mapping does (nvblox_mapping_helpers.py:12-21, nvblox_output_helpers.py:13-14,
    feature_extraction.py:18, visualizer.py:20-21, paper/utils/utils.py:16-18) resolves."""
    from nvblox_torch.constants import constants
    from nvblox_torch.indexing import get_voxel_center_grids
    from nvblox_torch.layer import FeatureLayer, Layer, TsdfLayer, convert_layer_to_dense_tensor
    from nvblox_torch.mapper import Mapper, QueryType
    from nvblox_torch.mapper_params import (BlockMemoryPoolParams, MapperParams, ProjectiveIntegratorParams,
                                            TsdfDecayIntegratorParams, ViewCalculatorParams)
    from nvblox_torch.mesh import FeatureMesh
    from nvblox_torch.projective_integrator_types import ProjectiveIntegratorType
    from nvblox_torch.timer import Timer, get_last_time, get_mean_time, print_timers, timer_status_string
    for name in ('add_depth_frame', 'add_color_frame', 'add_feature_frame', 'update_feature_mesh', 'get_feature_mesh',
                 'decay', 'clear', 'save_map', 'num_mappers', 'update_color_mesh', 'get_color_mesh', 'tsdf_layer_view',
                 'feature_layer_view', 'query_layer', 'load_from_file', 'params'):
        assert callable(getattr(Mapper, name))
    assert constants.feature_array_num_elements() % 8 == 0
    assert ProjectiveIntegratorType.TSDF.value == 'tsdf' and QueryType.FEATURE.value == 'feature'
    assert TsdfLayer.num_elements_per_voxel() == 2
    assert FeatureLayer.num_elements_per_voxel() == constants.feature_array_num_elements() + 1
    assert FeatureMesh().vertices().shape == (0, 3)     # test_mesh.py:21-33 (empty mesh shapes)
    assert FeatureMesh().triangles().shape == (0, 3)
    assert FeatureMesh().vertex_features().shape == (0, constants.feature_array_num_elements())
    with Timer('t'):
        pass
    assert get_last_time('t') >= 0 and get_mean_time('t') >= 0 and 't' in timer_status_string()
    grids = get_voxel_center_grids([torch.tensor([1, 0, -1], dtype=torch.int32)], 0.02, device='cpu')
    assert grids[0].shape == (8, 8, 8, 3)
    assert torch.allclose(grids[0][0, 0, 0], torch.tensor([0.17, 0.01, -0.15]), atol=1e-6)


def test_mindmap_parameter_plumbing():
    """get_nvblox_mapper (nvblox_mapping_helpers.py:30-76) written against our classes -> the C POD."""
    from nvblox_mindmap_b200.params import NvbxParams
    from nvblox_torch.mapper_params import (BlockMemoryPoolParams, MapperParams, ProjectiveIntegratorParams,
                                            TsdfDecayIntegratorParams, ViewCalculatorParams)
    pi = ProjectiveIntegratorParams()
    pi.projective_integrator_max_integration_distance_m = 5.0
    pi.projective_appearance_integrator_measurement_weight = 1.0
    td = TsdfDecayIntegratorParams()
    td.tsdf_decay_factor = 0.98
    vc = ViewCalculatorParams()
    vc.raycast_subsampling_factor = 1
    vc.workspace_bounds_type = 'kBoundingBox'
    vc.workspace_bounds_min_corner_x_m = -0.25
    vc.workspace_bounds_min_corner_y_m = -0.65
    vc.workspace_bounds_min_height_m = -0.07
    vc.workspace_bounds_max_corner_x_m = 1.0
    vc.workspace_bounds_max_corner_y_m = 0.62
    vc.workspace_bounds_max_height_m = 0.56
    bp = BlockMemoryPoolParams()
    bp.expansion_factor = 1.0
    bp.num_preallocated_blocks = 0
    mp = MapperParams()
    mp.set_projective_integrator_params(pi)
    mp.set_tsdf_decay_integrator_params(td)
    mp.set_view_calculator_params(vc)
    mp.set_block_memory_pool_params(bp)
    p = mp.to_nvbx()
    assert isinstance(p, NvbxParams)
    assert p.max_integration_distance_m == 5.0 and p.appearance_measurement_weight == 1.0
    assert abs(p.tsdf_decay_factor - 0.98) < 1e-7 and p.raycast_subsampling_factor == 1
    assert p.workspace_bounds_type == 2
    assert np.allclose(list(p.workspace_min), [-0.25, -0.65, -0.07]) and np.allclose(list(p.workspace_max),
                                                                                    [1.0, 0.62, 0.56])
    assert mp.get_view_calculator_params().get_raycast_subsampling_factor() == 1
    with pytest.raises(AttributeError):
        pi.not_a_parameter = 1
    # py_mapper_params.cpp:15-35: unknown weighting strings fall into the distance-penalty mode
    pi.projective_integrator_weighting_mode = 'kLinearWithMax'
    mp.set_projective_integrator_params(pi)
    assert mp.to_nvbx().weighting_mode == 4


def test_check_integrator_inputs_contract():
    """test_mapper_add_frames.py:141-206: wrong dtype / dim / device raise AssertionError."""
    from nvblox_torch.mapper import check_integrator_inputs
    T, K = torch.eye(4), torch.eye(3)
    cpu_img = torch.zeros(4, 4)
    with pytest.raises(AssertionError):    # image must be on the GPU
        check_integrator_inputs(cpu_img, T, K, 'Depth', 2, torch.float32)
    with pytest.raises(AssertionError):    # wrong number of dims (checked before the device)
        check_integrator_inputs(torch.zeros(4, 4, 1), T, K, 'Depth', 2, torch.float32)
    if torch.cuda.is_available():
        img = torch.zeros(4, 4, device='cuda')
        check_integrator_inputs(img, T, K, 'Depth', 2, torch.float32)
        with pytest.raises(AssertionError):
            check_integrator_inputs(img.double(), T, K, 'Depth', 2, torch.float32)
        with pytest.raises(AssertionError):
            check_integrator_inputs(img, T.cuda(), K, 'Depth', 2, torch.float32)
        with pytest.raises(AssertionError):
            check_integrator_inputs(torch.zeros(4, 4, 5, device='cuda', dtype=torch.float16), T, K, 'Feature', 3,
                                    torch.float16, 8)


def test_mc_table_structure():
    """Every marching-cubes row uses exactly the edges whose end points have different signs."""
    import subprocess
    import sys
    assert subprocess.call([sys.executable, os.path.join(ROOT, 'tools', 'gen_mc_tables.py'), '--check'],
                           stdout=subprocess.DEVNULL) == 0


def test_replica_partitioning():
    from nvblox_mindmap_b200.replicas import aggregate_throughput, maps_of_rank, owner_of_map
    for world in (1, 2, 4, 8):
        owned = [maps_of_rank(64, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(64))
        assert all(len(o) == 64 // world for o in owned)
        assert all(owner_of_map(m, world) == r for r, o in enumerate(owned) for m in o)
    assert aggregate_throughput([10, 10], [1.0, 2.0]) == 10.0


def test_frames_batch_argument_errors_without_gpu():
    """nvbx_integrate_frames_batch validates its job list before touching the device."""
    import ctypes as C
    from nvblox_mindmap_b200 import _capi
    from nvblox_mindmap_b200.params import NvbxFrameJob
    lib = _capi.load()
    assert C.sizeof(NvbxFrameJob) == 152
    assert lib.nvbx_integrate_frames_batch(None, 0, 0) == 0
    assert lib.nvbx_integrate_frames_batch(None, 2, 0) < 0
    jobs = (NvbxFrameJob * 2)()
    assert lib.nvbx_integrate_frames_batch(jobs, 2, 4) < 0
    assert b'null mapper handle' in lib.nvbx_last_error()
