"""ctypes mirror of `nvbx_params` / `nvbx_counters` (include/nvbx_c_api.h)."""
import ctypes as C


class NvbxParams(C.Structure):
    """POD carrying every reference parameter that reaches the hot path (see nvbx_c_api.h)."""

    _fields_ = [
        ('max_integration_distance_m', C.c_float),
        ('truncation_distance_vox', C.c_float),
        ('weighting_mode', C.c_int32),
        ('max_weight', C.c_float),
        ('invalid_depth_decay_factor', C.c_float),
        ('appearance_measurement_weight', C.c_float),
        ('appearance_truncation_distance_vox', C.c_float),
        ('sphere_tracing_subsampling', C.c_int32),
        ('sphere_tracing_max_ray_length_m', C.c_float),
        ('sphere_tracing_max_steps', C.c_int32),
        ('sphere_tracing_surface_epsilon_vox', C.c_float),
        ('tsdf_decay_factor', C.c_float),
        ('tsdf_decayed_weight_threshold', C.c_float),
        ('tsdf_set_free_distance_on_decayed', C.c_int32),
        ('tsdf_decayed_free_distance_vox', C.c_float),
        ('deallocate_decayed_blocks', C.c_int32),
        ('raycast_subsampling_factor', C.c_int32),
        ('workspace_bounds_type', C.c_int32),
        ('workspace_min', C.c_float * 3),
        ('workspace_max', C.c_float * 3),
        ('cache_last_viewpoint', C.c_int32),
        ('mesh_min_weight', C.c_float),
        ('mesh_weld_vertices', C.c_int32),
        ('mesh_cutoff_distance_vox', C.c_float),
        ('num_preallocated_blocks', C.c_int32),
        ('expansion_factor', C.c_float),
        ('strict_blend', C.c_int32),
    ]

    def copy(self) -> 'NvbxParams':
        out = NvbxParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(NvbxParams))
        return out


class NvbxCounters(C.Structure):
    _fields_ = [
        ('depth_frames', C.c_int64),
        ('feature_frames', C.c_int64),
        ('tsdf_blocks_in_view', C.c_int64),
        ('tsdf_voxels_updated', C.c_int64),
        ('tsdf_blocks_allocated', C.c_int64),
        ('feature_candidate_blocks', C.c_int64),
        ('feature_band_blocks', C.c_int64),
        ('feature_voxels_updated', C.c_int64),
        ('feature_blocks_allocated', C.c_int64),
        ('blocks_deallocated', C.c_int64),
        ('mesh_blocks_remeshed', C.c_int64),
        ('mesh_vertices', C.c_int64),
        ('color_frames', C.c_int64),
        ('color_band_blocks', C.c_int64),
        ('color_voxels_updated', C.c_int64),
        ('color_blocks_allocated', C.c_int64),
        ('host_pixels_fetched', C.c_int64),
        ('reserved', C.c_int64 * 4),
    ]

    def as_dict(self) -> dict:
        d = {n: int(getattr(self, n)) for n, _ in self._fields_ if n != 'reserved'}
        if any(self.reserved):                     # libnvbx_prof.so only (NVBX_PROFILE=1)
            d['profile'] = [int(v) for v in self.reserved]
        return d


WEIGHTING_MODES = {
    'kConstantWeight': 0,
    'kConstantDropoffWeight': 1,
    'kInverseSquareWeight': 2,
    'kInverseSquareDropoffWeight': 3,
    'kInverseSquareTsdfDistancePenalty': 4,
    'kLinearWithMax': 5,
}
WORKSPACE_BOUNDS_TYPES = {'kUnbounded': 0, 'kHeightBounds': 1, 'kBoundingBox': 2}


class NvbxFrameJob(C.Structure):
    """`nvbx_frame_job` of include/nvbx_c_api.h (one map's frame inside nvbx_integrate_frames_batch)."""
    _fields_ = [
        ('mapper', C.c_void_p),
        ('map_id', C.c_int32),
        ('height', C.c_int32),
        ('width', C.c_int32),
        ('channels', C.c_int32),
        ('depth', C.c_void_p),
        ('depth_mask', C.c_void_p),
        ('features', C.c_void_p),
        ('feature_mask', C.c_void_p),
        ('T_L_C', C.c_float * 16),
        ('fx', C.c_float),
        ('fy', C.c_float),
        ('cx', C.c_float),
        ('cy', C.c_float),
        ('stream', C.c_void_p),
        ('status', C.c_int32),
        ('reserved', C.c_int32),
    ]
