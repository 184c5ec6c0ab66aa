/*
 * nvbx_c_api.h -- C ABI of libnvbx.so: the B200-native reconstruction hot path behind the
 * nvblox_torch `Mapper` surface used by mindmap/mapping.
 *
 * Every entry point replaces one method of the reference's torch custom class
 * `pynvblox::Mapper` (TORCH_LIBRARY(pynvblox) in
 * submodules/nvblox/nvblox_torch/cpp/src/py_nvblox.cu:55-344) or of its layer / mesh holders.
 * The reference-side file:line each function stands in for is cited next to it.
 *
 * Conventions
 *   - plain C types only: device pointers are passed as `const void*` / `void*`, poses as 16
 *     row-major floats (T_L_C, camera -> layer/world), intrinsics as fx, fy, cx, cy.
 *   - `stream` is a cudaStream_t cast to void* (0 = legacy default stream).  All kernels are enqueued
 *     on it; calls do NOT synchronise unless documented (the reference synchronises after every kernel:
 *     projective_integrator_impl.cuh:341,443).
 *   - every function returns NVBX_OK (0) or a negative error code; nvbx_last_error() returns the
 *     message of the calling thread's last failure.  Nothing aborts the process.
 *   - a handle is not thread-safe; use one handle per host thread / per GPU.
 */
#ifndef NVBX_C_API_H_
#define NVBX_C_API_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVBX_OK 0
#define NVBX_ERR_INVALID_ARGUMENT (-1)
#define NVBX_ERR_CUDA (-2)
#define NVBX_ERR_OUT_OF_MEMORY (-3)
#define NVBX_ERR_NOT_FOUND (-4)
#define NVBX_ERR_UNSUPPORTED (-5)

/* Layer selectors. */
#define NVBX_LAYER_TSDF 0
#define NVBX_LAYER_FEATURE 1
#define NVBX_LAYER_COLOR 2

/* WeightingFunctionType, nvblox/integrators/weighting_function.h:11-18 */
#define NVBX_WEIGHT_CONSTANT 0
#define NVBX_WEIGHT_CONSTANT_DROPOFF 1
#define NVBX_WEIGHT_INVERSE_SQUARE 2
#define NVBX_WEIGHT_INVERSE_SQUARE_DROPOFF 3
#define NVBX_WEIGHT_INVERSE_SQUARE_TSDF_DISTANCE_PENALTY 4
#define NVBX_WEIGHT_LINEAR_WITH_MAX 5

/* WorkspaceBoundsType, nvblox/geometry/workspace_bounds.h:24 */
#define NVBX_WORKSPACE_UNBOUNDED 0
#define NVBX_WORKSPACE_HEIGHT_BOUNDS 1
#define NVBX_WORKSPACE_BOUNDING_BOX 2

/*
 * One POD carrying every reference parameter that reaches the hot path.  Defaults (set by
 * nvbx_default_params) are the reference's: projective_integrator_params.h:24-73,
 * view_calculator_params.h:23-59, tsdf_decay_integrator_params.h:22-38,
 * decay_integrator_base_params.h:22-24, mesh_integrator_params.h:21-26, mesh_integrator.h:126-133,
 * sphere_tracer.h:216-218, projective_appearance_integrator.cu:61 (ray length frozen at 7 m).
 */
typedef struct nvbx_params {
  /* ProjectiveIntegratorParams */
  float max_integration_distance_m;          /* 7.0  */
  float truncation_distance_vox;             /* 4.0  (TSDF integrator) */
  int32_t weighting_mode;                    /* NVBX_WEIGHT_INVERSE_SQUARE */
  float max_weight;                          /* 5.0  */
  float invalid_depth_decay_factor;          /* -1.0 (disabled) */
  float appearance_measurement_weight;       /* 0.8  (mindmap: 1.0) */
  /* The appearance integrators keep their own truncation distance; Mapper::setMapperParams
   * (mapper.cpp:126-141) never forwards projective_integrator_truncation_distance_vox to them, so it
   * stays at the default unless set directly on the integrator (as the reference's gtests do). */
  float appearance_truncation_distance_vox;  /* 4.0  */
  int32_t sphere_tracing_subsampling;        /* 4    */
  float sphere_tracing_max_ray_length_m;     /* 7.0  */
  int32_t sphere_tracing_max_steps;          /* 100  */
  float sphere_tracing_surface_epsilon_vox;  /* 0.1  */
  /* TsdfDecayIntegratorParams + DecayIntegratorBaseParams */
  float tsdf_decay_factor;                   /* 0.95 */
  float tsdf_decayed_weight_threshold;       /* 1e-3 */
  int32_t tsdf_set_free_distance_on_decayed; /* 0    */
  float tsdf_decayed_free_distance_vox;      /* 4.0  */
  int32_t deallocate_decayed_blocks;         /* 1    */
  /* ViewCalculatorParams */
  int32_t raycast_subsampling_factor;        /* 4 (mindmap: 1) */
  int32_t workspace_bounds_type;             /* NVBX_WORKSPACE_UNBOUNDED */
  float workspace_min[3];                    /* x 0, y 2, z(height) 0 */
  float workspace_max[3];                    /* x 0, y 2, z(height) 1 */
  int32_t cache_last_viewpoint;              /* 1    */
  /* MeshIntegratorParams */
  float mesh_min_weight;                     /* 1e-4 */
  int32_t mesh_weld_vertices;                /* 1    */
  float mesh_cutoff_distance_vox;            /* 5.0  */
  /* BlockMemoryPoolParams: accepted for drop-in compatibility.  Our pools are slab arenas; these two
   * only seed the initial slab size. */
  int32_t num_preallocated_blocks;           /* 2048 */
  float expansion_factor;                    /* 2.0  */
  /* Ours.  0 (default): when appearance_measurement_weight == 1 the old feature vector is not
   * re-read (x_new = x_meas; the reference computes 0*x_old + 1*x_meas in fp16, which differs only
   * in the sign of zero and in NaN poisoning from non-finite old values).  1: always read + blend. */
  int32_t strict_blend;
} nvbx_params;

typedef struct nvbx_mapper nvbx_mapper;

/* Per-map counters that define the algorithmic bytes of SURVEY.md 8(d); accumulated on the device,
 * read back (with a stream sync) by nvbx_get_counters.  Reset by nvbx_reset_counters. */
typedef struct nvbx_counters {
  int64_t depth_frames;        /* nvbx_integrate_depth calls                                   */
  int64_t feature_frames;      /* nvbx_integrate_features calls                                */
  int64_t tsdf_blocks_in_view; /* sum over depth frames of blocks handed to the TSDF update    */
  int64_t tsdf_voxels_updated; /* N_tsdf: voxels whose TSDF value/weight was rewritten         */
  int64_t tsdf_blocks_allocated;
  int64_t feature_candidate_blocks; /* N_cand: TSDF blocks read by the truncation-band test    */
  int64_t feature_band_blocks;      /* blocks handed to the feature update                     */
  int64_t feature_voxels_updated;   /* N_upd                                                   */
  int64_t feature_blocks_allocated;
  int64_t blocks_deallocated;
  int64_t mesh_blocks_remeshed;
  int64_t mesh_vertices;            /* N_v of the last feature-mesh update                     */
  int64_t color_frames;             /* nvbx_integrate_color calls                              */
  int64_t color_band_blocks;        /* blocks handed to the colour update                      */
  int64_t color_voxels_updated;
  int64_t color_blocks_allocated;
  int64_t host_pixels_fetched;      /* nvbx_integrate_frame_host: feature pixels that crossed PCIe */
  int64_t reserved[4];              /* NVBX_PROFILE_COUNTERS builds only                       */
} nvbx_counters;

/* ---- lifecycle ---------------------------------------------------------------------------------- */

/* Fill `p` with the reference defaults listed above. */
void nvbx_default_params(nvbx_params* p);

/* pynvblox::Mapper::Mapper(voxel_sizes, integrator_types, params), py_mapper.cu:33-70.  Only TSDF maps
 * exist on this path.  `feature_channels` is the reference's compile-time
 * NVBLOX_FEATURE_ARRAY_NUM_ELEMENTS (core/feature_array.h:22-28), a run-time value here; it must be a
 * multiple of 8 (128-bit vectors).  `device` is the CUDA ordinal (the reference hard-codes 0). */
int nvbx_create(int n_maps, const float* voxel_sizes_m, const nvbx_params* params, int feature_channels,
                int device, nvbx_mapper** out);
void nvbx_destroy(nvbx_mapper* m);
int nvbx_num_maps(const nvbx_mapper* m);       /* Mapper::getNumMappers, py_mapper.cu:72 */
int nvbx_feature_channels(const nvbx_mapper* m);
int nvbx_get_params(const nvbx_mapper* m, nvbx_params* out); /* Mapper::getMapperParams, py_mapper.cu:80 */
const char* nvbx_last_error(void);

/* ---- frame integration -------------------------------------------------------------------------- */

/* Mapper::integrateDepth, py_mapper.cu:85-113 -> nvblox::Mapper::integrateDepth mapper.cpp:358-407.
 * depth: device float[H*W] row-major; mask: device uint8[H*W] or NULL (non-zero = active).
 * Stream contract: the depth image must be complete in stream order, as for any kernel argument.  The ray-casting
 * kernel is a programmatic dependent launch that reads the image (and nothing else in memory) BEFORE its
 * griddepcontrol.wait, to overlap the previous frame's feature kernel; this is invisible as long as the image's
 * producer is an ordinary stream operation (memcpy, any kernel that does not itself execute
 * griddepcontrol.launch_dependents before its last write to the image) -- true of torch, cuBLAS/cuDNN and this
 * library's own outputs.  NVBX_PDL=0 disables programmatic launches altogether. */
int nvbx_integrate_depth(nvbx_mapper* m, int map_id, const void* depth, int height, int width,
                         const void* mask, const float* T_L_C, float fx, float fy, float cx, float cy,
                         void* stream);

/* Mapper::integrateFeatures, py_mapper.cu:146-174 -> mapper.cpp:452-464 ->
 * ProjectiveAppearanceIntegrator<FeatureLayer>::integrateFrame projective_appearance_integrator.cu:72-169.
 * features: device fp16 [H*W*C] (HWC, contiguous, 16-byte aligned); channels must equal
 * nvbx_feature_channels(); height/width must be multiples of sphere_tracing_subsampling
 * (sphere_tracer.cu:426-427 CHECKs the same). */
int nvbx_integrate_features(nvbx_mapper* m, int map_id, const void* features, int height, int width,
                            int channels, const void* mask, const float* T_L_C, float fx, float fy,
                            float cx, float cy, void* stream);

/* Mapper::integrateColor, py_mapper.cu:115-144 -> mapper.cpp:436-449 ->
 * ProjectiveAppearanceIntegrator<ColorLayer>::integrateFrame projective_appearance_integrator.cu:72-169
 * (SURVEY 8(f) N1).  rgb: device uint8 [H*W*3] (HWC); height/width must be multiples of
 * sphere_tracing_subsampling.  The colour layer stores the reference's ColorVoxel record (r, g, b, pad,
 * float weight: 8 bytes, voxels.h:77-83); a colour frame that follows a feature / colour frame with the same
 * pose, intrinsics and TSDF state re-uses that frame's synthetic depth image instead of sphere tracing again
 * (the result is identical: the image is a pure function of those three). */
int nvbx_integrate_color(nvbx_mapper* m, int map_id, const void* rgb, int height, int width,
                         const void* mask, const float* T_L_C, float fx, float fy, float cx, float cy,
                         void* stream);

/* Same as integrate_depth + integrate_features but from HOST buffers: the end-to-end entry bench.py times.
 * (The reference accepts CUDA tensors only -- mapper.py:458-490 -- so its caller pays `.cuda()` on the whole
 * H*W*C frame first.)  Depth and masks are copied H2D on `stream`.  The feature frame:
 *   - pinned / registered host memory (cudaHostAlloc, cudaHostRegister, torch `.pin_memory()`), 16-byte aligned,
 *     mode NVBX_HOST_FETCH_SPARSE (default): the GPU reads it through its device mapping and fetches ONLY the
 *     pixels this frame's voxels sample (each distinct pixel crosses PCIe once, 2C bytes); the map result is
 *     identical to the dense copy.  The buffer must stay unchanged until `stream` has passed this call.
 *   - pageable memory, or mode NVBX_HOST_FETCH_DENSE: one cudaMemcpyAsync of all H*W*C halves.
 * nvbx_counters.host_pixels_fetched counts the pixels the sparse path moved. */
#define NVBX_HOST_FETCH_SPARSE 0
#define NVBX_HOST_FETCH_DENSE 1
int nvbx_set_host_fetch_mode(nvbx_mapper* m, int mode);
int nvbx_integrate_frame_host(nvbx_mapper* m, int map_id, const float* depth_host, const void* features_host,
                              int height, int width, int channels, const uint8_t* depth_mask_host,
                              const uint8_t* feature_mask_host, const float* T_L_C, float fx, float fy,
                              float cx, float cy, void* stream);

/* ---- batched datagen (BASELINE configs[3]): many independent maps per GPU ------------------------
 * mindmap's datagen steps N environments and integrates one frame per environment per step
 * (mindmap/run_isaaclab_datagen.py:194-235, one IsaacLabNvbloxMapper per env).  The frame of a single map is five
 * latency-bound launches around one memory-bound one, so one map leaves most of a B200 idle; maps are independent,
 * so their frames can overlap.  nvbx_integrate_frames_batch integrates one frame (depth, then features when
 * `features` is not NULL: exactly nvbx_integrate_depth + nvbx_integrate_features) for each job, each on the job's own
 * stream, with the jobs of DIFFERENT mapper handles issued concurrently by an internal pool of `n_threads` host
 * threads (jobs of one handle keep their order on one thread).  Returns when every job has been ENQUEUED (not
 * executed); job[i].status holds each job's status, the return value is the first failure (its message in
 * nvbx_last_error()).  A handle must not appear in two concurrent batch calls. */
typedef struct nvbx_frame_job {
  nvbx_mapper* mapper;
  int32_t map_id;
  int32_t height, width, channels;
  const void* depth;        /* device float[H*W] */
  const void* depth_mask;   /* device uint8[H*W] or NULL */
  const void* features;     /* device fp16[H*W*C] or NULL (depth only) */
  const void* feature_mask; /* device uint8[H*W] or NULL */
  float T_L_C[16];
  float fx, fy, cx, cy;
  void* stream;
  int32_t status; /* out */
  int32_t reserved;
} nvbx_frame_job;
int nvbx_integrate_frames_batch(nvbx_frame_job* jobs, int n_jobs, int n_threads);

/* ---- the feature extractor's up-sampling fused into the integration ----------------
 * Replaces, for one frame, mindmap's
 *   FeatureExtractor.compute(): scale_image(features_bchw, (H, W)) -> rearrange -> zero-pad -> .to(float16)
 *   (mindmap/image_processing/feature_extraction.py:110-129,170-211, nvblox_mapping_helpers.py:255-261)
 * followed by Mapper::integrateFeatures.  `lowres` is the backbone's patch-feature map (e.g. RADIO 32x32x768):
 *   dtype  NVBX_LOWRES_F32 / _F16 / _BF16 (the dtype torch.nn.functional.interpolate would have run in);
 *   layout NVBX_LOWRES_CHW (a contiguous [1,c,h,w] tensor) or NVBX_LOWRES_HWC (a channels-last one, which is
 *          what RADIO / DINOv2's permuted view is: feature_extraction.py:328-330);
 *   kernel NVBX_UPSAMPLE_TORCH_NCHW or NVBX_UPSAMPLE_TORCH_NHWC: which of PyTorch's two CUDA kernels the
 *          chained path would have selected (they contract multiply-adds differently for fp32; see
 *          csrc/nvbx_upsample.cuh).  torch picks NHWC iff the input is channels-last and has >= 16 channels.
 * low_c <= nvbx_feature_channels(); missing channels are zero.  The stored features are bit-identical to the
 * chained path's; the [H, W, C] frame is never materialised.  mask (device uint8[H*W] or NULL) and the camera
 * refer to the up-sampled H x W frame. */
#define NVBX_LOWRES_F32 0
#define NVBX_LOWRES_F16 1
#define NVBX_LOWRES_BF16 2
#define NVBX_LOWRES_CHW 0
#define NVBX_LOWRES_HWC 1
#define NVBX_UPSAMPLE_TORCH_NCHW 0
#define NVBX_UPSAMPLE_TORCH_NHWC 1
int nvbx_integrate_features_lowres(nvbx_mapper* m, int map_id, const void* lowres, int low_h, int low_w,
                                   int low_c, int dtype, int layout, int kernel, int height, int width,
                                   const void* mask, const float* T_L_C, float fx, float fy, float cx,
                                   float cy, void* stream);

/* The [H, W, C] fp16 frame the chained path would have produced, written to `out` (device, 16-byte aligned):
 * parity checks and visualisation only. */
int nvbx_upsample_features(nvbx_mapper* m, int map_id, const void* lowres, int low_h, int low_w, int low_c,
                           int dtype, int layout, int kernel, int height, int width, void* out, void* stream);

/* integrate_depth + integrate_features_lowres from HOST buffers: depth (H*W floats) and the low-res feature
 * map (low_h*low_w*low_c elements) are the only per-frame uploads -- 2.5 MB instead of 385 MB at 512^2 x 768. */
int nvbx_integrate_frame_host_lowres(nvbx_mapper* m, int map_id, const float* depth_host,
                                     const void* lowres_host, int low_h, int low_w, int low_c, int dtype,
                                     int layout, int kernel, int height, int width,
                                     const uint8_t* depth_mask_host, const uint8_t* feature_mask_host,
                                     const float* T_L_C, float fx, float fy, float cx, float cy, void* stream);

/* Mapper::decayTsdf, py_mapper.cu:264-273 -> mapper.cpp:466-495.  map_id -1 = all maps. */
int nvbx_decay(nvbx_mapper* m, int map_id, void* stream);

/* Mapper::clear, py_mapper.cu:286-306. */
int nvbx_clear(nvbx_mapper* m, int map_id, void* stream);

/* Mark every TSDF block "to update" for both mesh layers (update*Mesh(UpdateFullLayer::kYes), what
 * Mapper::loadMap does after swapping in a loaded layer cake, mapper.cpp:884-900).  Used by load_from_file after
 * the block payloads were written through the layer views. */
int nvbx_mark_all_dirty(nvbx_mapper* m, int map_id, void* stream);

/* ---- frame pipelining (ours) ------------------------------------------------------------------------------
 * nvbx_set_pipelining(m, 1): the geometry kernel and the memory-bound gather of feature frame i run on two streams
 * owned by the map, so that the latency-bound depth path of frame i + 1 (raycast, TSDF update, sphere tracing + band
 * selection), which the caller enqueues next on ITS stream, runs underneath them instead of behind them.  Results are
 * unchanged (bit for bit).
 * Contract while it is on: the FEATURE frame (and its mask) passed to nvbx_integrate_features must stay valid and
 * unmodified until four more nvbx_integrate_features calls on that map have returned (the per-frame lists live in a
 * ring of four; the host waits for a slot's last gather before re-using it), or until any call that reads or frees
 * feature data (decay, clear, mesh update, feature block views, feature queries, nvbx_pipeline_join) -- every such
 * call first orders its stream behind the gathers in flight.  Depth frames, depth masks and colour frames are
 * consumed on the caller's stream as before.  Off by default (the reference's callers free or overwrite a frame right
 * after the call).
 * nvbx_set_pipelining(m, 2): the same, plus ASYNCHRONOUS ENQUEUE -- nvbx_integrate_depth / nvbx_integrate_features
 * validate their arguments, queue them (up to eight calls) and return; a worker thread owned by the mapper issues the
 * CUDA work in call order on the stream each call named.  Every other entry point taking this handle first waits for
 * the queue to be issued, and returns the error of a queued frame if there was one (frames queued behind a failed one
 * are dropped).  Additional contract: EVERY input buffer of a queued call stays valid and unmodified until such a
 * joining call, and work the caller itself enqueues on the stream is no longer ordered behind a queued frame. */
int nvbx_set_pipelining(nvbx_mapper* m, int on);
/* Order `stream` behind every gather in flight of map_id (-1: all maps). */
int nvbx_pipeline_join(nvbx_mapper* m, int map_id, void* stream);

/* ---- surface extraction ------------------------------------------------------------------------- */

/* Mapper::updateFeatureMesh, py_mapper.cu:196-204 -> Mapper::updateMeshTemplate mapper.cpp:580-614
 * (MeshIntegrator::integrateBlocksGPU mesh_integrator.cu:64-103 + updateAppearance
 * mesh_integrator_appearance.cu:290-340).  Synchronises `stream` once (the vertex total sizes the
 * output arena). */
int nvbx_update_feature_mesh(nvbx_mapper* m, int map_id, void* stream);

/* Mapper::getFeatureMesh, py_mapper.cu:223-239 + PyMesh::vertices/vertex_appearances/triangles
 * py_mesh.cpp:19-90.  Returns DEVICE pointers owned by the handle, valid until the next
 * update_feature_mesh / clear / decay of that map: vertices float[n_vertices*3], features
 * fp16[n_vertices*C], triangles int32[n_triangles*3] (global vertex ids).  No kernel, no copy. */
int nvbx_get_feature_mesh(nvbx_mapper* m, int map_id, const void** vertices, const void** features,
                          const void** triangles, int64_t* n_vertices, int64_t* n_triangles);

/* Mapper::updateColorMesh, py_mapper.cu:186-194 (the colour mesh is its own mesh layer with its own
 * "to update" set, blocks_to_update_tracker.cpp:32-60) and Mapper::getColorMesh, py_mapper.cu:206-221 +
 * PyMesh<Color>::vertex_appearances py_mesh.cpp:54-90.  colors: uint8[n_vertices*3]. */
int nvbx_update_color_mesh(nvbx_mapper* m, int map_id, void* stream);
int nvbx_get_color_mesh(nvbx_mapper* m, int map_id, const void** vertices, const void** colors,
                        const void** triangles, int64_t* n_vertices, int64_t* n_triangles);

/* ---- fused export post-processing -------------------------------------------------- */

/* mindmap's get_vertices_and_features, mapping/helpers/nvblox_output_helpers.py:57-74, in one flag pass + one
 * scatter pass on the device: keep the vertices strictly inside (aabb_min, aabb_max), drop the trailing
 * `num_excess_features` channels, and (remove_zero_features != 0) drop the rows whose remaining channels are all
 * == 0.  Order is preserved (boolean-mask indexing).  vertices (float[n*3]) / features (fp16[n*channels]) are
 * device pointers; vertices == NULL selects the current feature mesh of `map_id` (n and channels are then
 * ignored).  The result lives in a handle-owned arena (valid until the next nvbx_export_points on this map):
 * out_vertices float[count*3], out_features fp16[count*(channels - num_excess_features)].  Returns the count
 * (one stream synchronisation) or a negative error. */
int64_t nvbx_export_points(nvbx_mapper* m, int map_id, const void* vertices, const void* features, int64_t n,
                           int channels, const float* aabb_min, const float* aabb_max, int num_excess_features,
                           int remove_zero_features, const void** out_vertices, const void** out_features,
                           void* stream);

/* mindmap's sample_to_n_vertices, data_loading/vertex_sampling.py:29-108, applied to the last
 * nvbx_export_points result of `map_id` in one pass: output row i = exported row indices[i] for i < n_indices
 * (indices: HOST int64, e.g. torch.randperm(count)[:n] drawn by the caller exactly as the reference draws it;
 * NULL = identity), zero rows for n_indices <= i < n_out (pad_with_zeros :84-108).  out_vertices: device
 * float[n_out*3]; out_features: device fp16 or (features_f32 != 0) float32 [n_out*C_keep] -- the float32 cast of
 * isaaclab_nvblox_mapper.py:243-246 rides the same pass. */
int nvbx_gather_points(nvbx_mapper* m, int map_id, const int64_t* indices, int64_t n_indices, int64_t n_out,
                       void* out_vertices, void* out_features, int features_f32, void* stream);

/* ---- layer views (PyVoxelBlockLayer, py_layer.cpp:24-47,99-198) ---------------------------------- */

int64_t nvbx_num_blocks(nvbx_mapper* m, int map_id, int layer, void* stream);           /* numBlocks            */
int64_t nvbx_num_allocated_blocks(nvbx_mapper* m, int map_id, int layer, void* stream); /* numAllocatedBlocks   */
int64_t nvbx_num_allocated_bytes(nvbx_mapper* m, int map_id, int layer, void* stream);  /* numAllocatedBytes    */
float nvbx_voxel_size(const nvbx_mapper* m, int map_id);
/* get_all_block_indices: writes up to `capacity` int32 triples to HOST memory `out_xyz`, returns the count. */
int64_t nvbx_get_block_indices(nvbx_mapper* m, int map_id, int layer, int32_t* out_xyz, int64_t capacity,
                               void* stream);
/* get_all_blocks (PyVoxelBlockLayer::getAllBlocks, py_layer.cpp:177-198): index triples AND device payload pointers
 * of every block of `layer` in ONE kernel launch and one stream synchronisation; HOST outputs (either may be
 * null), up to `capacity` entries; returns the block count.  Pointer layout and *voxel_stride_elems as for
 * nvbx_get_block_ptr below.  Pointers stay valid until the map is next mutated. */
int64_t nvbx_get_all_blocks(nvbx_mapper* m, int map_id, int layer, int32_t* out_xyz, void** out_ptrs, int64_t capacity,
                            int64_t* voxel_stride_elems, void* stream);
/* get_block_at_index: device pointer of the block's voxel array and its voxel stride in ELEMENTS.
 * TSDF: float [8][8][8][2], stride 2.  Feature: fp16 [8][8][8][stride], stride = C + 8 (the first C are
 * the feature, element C is the weight, the rest is padding that keeps rows 16-byte aligned), to be
 * viewed as [8,8,8,C+1].  Colour: uint8 [8][8][8][8] = (r, g, b, pad, float weight), stride 8, to be viewed
 * as the [8,8,8,3] uint8 tensor of py_layer.cpp:49-70.  NVBX_ERR_NOT_FOUND if the block is not allocated. */
int nvbx_get_block_ptr(nvbx_mapper* m, int map_id, int layer, int x, int y, int z, void** ptr,
                       int64_t* voxel_stride_elems, void* stream);
/* allocate_block_at_index (zero-initialised). */
int nvbx_allocate_block(nvbx_mapper* m, int map_id, int layer, int x, int y, int z, void* stream);
/* Layer::clear of one layer is not exposed by mindmap; nvbx_clear drops everything. */

/* ---- point queries (queryTsdf / queryFeatures, py_mapper.cu:596-698; sdf_query.cu:206-270) -------- */

/* xyz: device float[n*3]; out: device float[n*2] = (distance, weight); rows of unallocated positions
 * are left untouched (the caller pre-fills zeros, mapper.py:352-356). */
int nvbx_query_tsdf(nvbx_mapper* m, int map_id, const void* xyz, int64_t n, void* out, void* stream);
/* out: device fp16[n*(C+1)] = (f_0..f_{C-1}, weight). */
int nvbx_query_features(nvbx_mapper* m, int map_id, const void* xyz, int64_t n, void* out, void* stream);

/* ---- accounting --------------------------------------------------------------------------------- */

int nvbx_get_counters(nvbx_mapper* m, int map_id, nvbx_counters* out, void* stream);
int nvbx_reset_counters(nvbx_mapper* m, int map_id, void* stream);
/* Live per-kernel timing for the roofline line of bench.py: while enabled, the library brackets every
 * launch of the dominant kernel (which = 0: k_feature_integrate, 1: k_tsdf_update) with CUDA events on
 * the launching stream.  nvbx_get_kernel_timing synchronises, returns the summed milliseconds and the
 * number of launches since timing was enabled, and resets the accumulators. */
int nvbx_set_kernel_timing(nvbx_mapper* m, int enabled);
/* enabled = 2: bracket EVERY kernel launch of the library with an event pair (tuning aid; perturbs the
 * pipeline).  nvbx_kernel_timing_report synchronises and writes a JSON object
 * {"kernel name": [total_ms, launches], ...} into `json` (returns its length or a negative error) and
 * resets the records. */
int64_t nvbx_kernel_timing_report(nvbx_mapper* m, char* json, int64_t capacity);
int nvbx_get_kernel_timing(nvbx_mapper* m, int which, double* total_ms, int64_t* launches);
/* Process-wide schedule of the feature gather kernel (tuning aid; results are identical for every setting):
 * variant 0..3 = static round-robin deal <units in flight, CTAs/SM>, 4..9 = k_feature_gather_dyn with
 * `dyn_permille`/1000 of the units handed out by atomic ticket, `ticket_units` per grab; dyn_permille = -1: each
 * CTA owns a contiguous range of the work-item list (L1 reuse between neighbouring voxels). */
int nvbx_set_gather_tuning(int variant, int dyn_permille, int ticket_units);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t nvbx_kernel_launch_count(void);
/* Frame pipelining back-pressure (tuning aid): how often, and for how many nanoseconds in total, the HOST had to wait
 * for a ring slot (the gather of four feature frames ago) since load.  Near zero: the host, not the device, sets the
 * frame rate.  Either pointer may be null. */
void nvbx_pipeline_wait_stats(int64_t* waits, int64_t* wait_ns);
/* Debug / parity hooks: copy the last frame's intermediate products to HOST memory.
 *   which = 0: block indices handed to the last TSDF update   (int32 triples)
 *   which = 1: block indices handed to the last feature update (int32 triples)
 *   which = 2: block indices handed to the last colour update  (int32 triples)
 * returns the count (or a negative error). */
int64_t nvbx_debug_last_block_list(nvbx_mapper* m, int map_id, int which, int32_t* out_xyz, int64_t capacity,
                                   void* stream);
/* last synthetic depth image (device float[rows*cols], sphere_tracer.cu:191-236) */
int nvbx_debug_last_synthetic_depth(nvbx_mapper* m, int map_id, const void** ptr, int* rows, int* cols);
/* Profile build only (NVBX_PROFILE=1 -> libnvbx_prof.so; NVBX_ERR_UNSUPPORTED otherwise): the pipeline timeline.
 * out[64][8][2] uint64 = (%globaltimer ns when the first CTA passed griddepcontrol.wait, ns when the last CTA
 * ended) of kernel k = 0 raycast, 1 tsdf, 2 trace+band, 3 geometry, 4 gather, 5 raycast's pre-wait march, for
 * frame (n mod 64).  Synchronises the device; reset != 0 re-arms the stamps.  tools/pipeline_timeline.py. */
int nvbx_debug_profile_stamps(nvbx_mapper* m, uint64_t* out, int reset);

const char* nvbx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NVBX_C_API_H_ */
