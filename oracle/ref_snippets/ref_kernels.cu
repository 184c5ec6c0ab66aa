// ref_kernels.cu -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
//
// Compiles the REFERENCE's own device code -- its headers included verbatim from /root/reference, its
// kernels / functors that live inside .cu files extracted by line range at build time into
// oracle/_ref/gen/*.inc (see Makefile; no reference source is committed) -- with nvblox's own nvcc flags
// (-O2, default -fmad=true, --expt-relaxed-constexpr; NB/cmake/nvblox_targets.cmake:128-152), against the
// Eigen / glog / stdgpu stand-ins under shim/.  The result, oracle/_ref/libref_kernels_c<C>.so, exposes the
// kernels through a C ABI so that tests/golden/make_ref_vectors.py can run them on a B200 on seeded
// inputs and record what the reference's arithmetic -- as nvcc contracts it -- produces.
//
// What is reference code here: every __global__ kernel and every __device__ function that computes
// something (projection, weighting, TSDF fuse, sphere-trace marching, bilinear interpolation, fp16 blend,
// marching-cubes vertex interpolation).  What is ours: the host-side glue below (block storage, the
// dense lookup table that stands in for the stdgpu hash, launches).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <thrust/pair.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <optional>
#include <vector>

#include "nvblox/core/indexing.h"
#include "nvblox/core/types.h"
#include "nvblox/gpu_hash/internal/cuda/gpu_indexing.cuh"   // -> shim gpu_hash_interface.cuh
#include "nvblox/integrators/weighting_function.h"
#include "nvblox/interpolation/interpolation_2d.h"
#include "nvblox/map/blox.h"
#include "nvblox/map/voxels.h"
#include "nvblox/mesh/internal/marching_cubes.h"
#include "nvblox/mesh/mesh_block.h"
#include "nvblox/rays/ray_caster.h"
#include "nvblox/sensors/camera.h"
#include "nvblox/sensors/image.h"
// device-only halves of the reference's headers
#include "nvblox/integrators/internal/cuda/impl/projective_integrators_common_impl.cuh"
#include "nvblox/mesh/internal/impl/cuda/marching_cubes_impl.cuh"

namespace nvblox {
// NB/include/nvblox/integrators/projective_integrator_params.h -- only the two constants the functors'
// default member initialisers name (the harness overwrites every member before use).
struct ShimParamDesc {
  float default_value;
};
struct ShimWeightDesc {
  WeightingFunctionType default_value;
};
// NB/include/nvblox/map/common_names.h:27 (that header drags in layer.h -> the stdgpu-backed GPU layer view)
using TsdfBlock = VoxelBlock<TsdfVoxel>;
constexpr ShimParamDesc kProjectiveIntegratorMaxWeightParamDesc{5.0f};
constexpr ShimParamDesc kProjectiveAppearanceIntegratorMeasurementWeightParamDesc{1.0f};
constexpr ShimWeightDesc kProjectiveIntegratorWeightingModeParamDesc{WeightingFunctionType::kInverseSquareWeight};

#include "gen/tsdf_functor.inc"          // NB/src/integrators/projective_tsdf_integrator.cu:25-99
#include "gen/tsdf_kernel.inc"           // NB/include/.../projective_integrator_impl.cuh:57-103
#include "gen/appearance_kernel.inc"     // NB/include/.../projective_integrator_impl.cuh:155-214
#include "gen/appearance_accessors.inc"  // NB/src/integrators/projective_appearance_integrator.cu:28-47
#include "gen/appearance_functor.inc"    // NB/src/integrators/projective_appearance_integrator.cu:267-353
#include "gen/raycast_kernel.inc"        // NB/src/integrators/view_calculator.cu:150-177,197-248
#include "gen/sphere_tracer.inc"         // NB/src/rays/sphere_tracer.cu:26-131,191-236
#include "gen/mesh_kernels.inc"          // NB/src/mesh/mesh_integrator.cu:327-487
#include "gen/mesh_appearance.inc"       // NB/src/mesh/mesh_integrator_appearance.cu:29-67,85-146
}  // namespace nvblox

using namespace nvblox;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "ref_kernels: %s -> %s\n", #x, cudaGetErrorString(e_));          \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

namespace {
Transform transform_from_row_major(const float* m) {
  Transform T;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.linear()(i, j) = m[i * 4 + j];
    T.translation()[i] = m[i * 4 + 3];
  }
  return T;
}
template <typename T>
struct DevBuf {
  T* p = nullptr;
  explicit DevBuf(size_t n) { cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  ~DevBuf() { cudaFree(p); }
};
template <typename T>
int upload(DevBuf<T>& b, const std::vector<T>& v) {
  if (!v.empty()) CK(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
}  // namespace

extern "C" {

int ref_feature_channels() { return (int)FeatureArray::size(); }
int ref_sizeof_feature_voxel() { return (int)sizeof(FeatureVoxel); }
int ref_sizeof_tsdf_block() { return (int)sizeof(VoxelBlock<TsdfVoxel>); }
int ref_sizeof_feature_block() { return (int)sizeof(VoxelBlock<FeatureVoxel>); }

// K1: combinedBlockIndicesInImageKernel<Camera>, launched as ViewCalculator::getBlocksByRaycastingPixelsAsync
// does (NB/src/integrators/view_calculator.cu:356-390: 16x16 threads over ceil((rows+1)/s) x ceil((cols+1)/s) rays).
// aabb_updated_dev: bool[prod(aabb_size)], zeroed by the caller.
int ref_raycast_blocks(const float* T_L_C_rm, float fu, float fv, float cu, float cv, int width, int height,
                       const float* depth_dev, float block_size, float max_integration_distance_m,
                       float max_behind_m, int subsampling, const int* aabb_min, const int* aabb_size,
                       uint8_t* aabb_updated_dev) {
  static_assert(sizeof(bool) == 1, "bool grid");
  const Transform T_L_C = transform_from_row_major(T_L_C_rm);
  const Camera camera(fu, fv, cu, cv, width, height);
  const int rays_rows = (int)std::ceil(static_cast<float>(height + 1) / static_cast<float>(subsampling));
  const int rays_cols = (int)std::ceil(static_cast<float>(width + 1) / static_cast<float>(subsampling));
  const dim3 threads(16, 16);
  const dim3 blocks((int)std::ceil(rays_cols / 16.0f), (int)std::ceil(rays_rows / 16.0f));
  combinedBlockIndicesInImageKernel<Camera><<<blocks, threads>>>(
      T_L_C, camera, depth_dev, height, width, block_size, max_integration_distance_m, max_behind_m, subsampling,
      Index3D(aabb_min[0], aabb_min[1], aabb_min[2]), Index3D(aabb_size[0], aabb_size[1], aabb_size[2]),
      reinterpret_cast<bool*>(aabb_updated_dev));
  CK(cudaDeviceSynchronize());
  return 0;
}

// K2: integrateBlocksKernel<TsdfVoxel, UpdateTsdfVoxelFunctor> over n blocks.  tsdf_storage_dev is an
// array of VoxelBlock<TsdfVoxel> (4096 B each); slots[i] selects the storage of block_idx[i].
int ref_integrate_tsdf(const float* T_C_L_rm, float fu, float fv, float cu, float cv, int width, int height,
                       const float* depth_dev, const uint8_t* mask_dev, float block_size,
                       float max_integration_distance_m, float truncation_distance_m, float max_weight,
                       float invalid_depth_decay_factor, int weighting_mode, const int* block_idx, const int* slots,
                       int n, void* tsdf_storage_dev) {
  if (n <= 0) return 0;
  const Transform T_C_L = transform_from_row_major(T_C_L_rm);
  const Camera camera(fu, fv, cu, cv, width, height);
  ImageView<const float> depth_view(height, width, depth_dev);
  std::optional<ImageView<const uint8_t>> mask;
  if (mask_dev) mask = ImageView<const uint8_t>(height, width, mask_dev);
  const MaskedDepthImageConstView image(depth_view, mask);

  UpdateTsdfVoxelFunctor op_host;
  op_host.truncation_distance_m_ = truncation_distance_m;
  op_host.max_weight_ = max_weight;
  op_host.invalid_depth_decay_factor_ = invalid_depth_decay_factor;
  op_host.weighting_function_ = WeightingFunction(static_cast<WeightingFunctionType>(weighting_mode));
  DevBuf<UpdateTsdfVoxelFunctor> op(1);
  CK(cudaMemcpy(op.p, &op_host, sizeof(op_host), cudaMemcpyHostToDevice));

  std::vector<Index3D> idx(n);
  std::vector<VoxelBlock<TsdfVoxel>*> ptrs(n);
  auto* base = static_cast<VoxelBlock<TsdfVoxel>*>(tsdf_storage_dev);
  for (int i = 0; i < n; ++i) {
    idx[i] = Index3D(block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]);
    ptrs[i] = base + slots[i];
  }
  DevBuf<Index3D> d_idx(n);
  DevBuf<VoxelBlock<TsdfVoxel>*> d_ptrs(n);
  if (upload(d_idx, idx) || upload(d_ptrs, ptrs)) return -1;
  integrateBlocksKernel<TsdfVoxel, UpdateTsdfVoxelFunctor><<<n, dim3(8, 8, 8)>>>(
      d_idx.p, camera, image, T_C_L, block_size, max_integration_distance_m, op.p, d_ptrs.p);
  CK(cudaDeviceSynchronize());
  return 0;
}

// K4: sphereTracingKernel(camera, T_L_C, hash, image, ...) as SphereTracer::renderImageOnGPU launches it
// (NB/src/rays/sphere_tracer.cu:421-480: 8x8 threads, cols/8+1 x rows/8+1 blocks over the subsampled image).  The map is given as
// all_block_idx[n_all] with storage slot i for entry i.
int ref_sphere_trace(const float* T_L_C_rm, float fu, float fv, float cu, float cv, int width, int height,
                     const int* all_block_idx, int n_all, void* tsdf_storage_dev, float truncation_distance_m,
                     float block_size, int maximum_steps, float maximum_ray_length_m, float surface_distance_epsilon_m,
                     int subsampling, float* out_depth_dev) {
  const Transform T_L_C = transform_from_row_major(T_L_C_rm);
  const Camera camera(fu, fv, cu, cv, width, height);
  Index3DDeviceHashMapType<TsdfBlock> hash;
  if (hash.build(all_block_idx, n_all, static_cast<TsdfBlock*>(tsdf_storage_dev))) return -1;
  const int rows = height / subsampling, cols = width / subsampling;
  const dim3 threads(8, 8, 1);
  const dim3 blocks(cols / 8 + 1, rows / 8 + 1, 1);
  sphereTracingKernel<<<blocks, threads>>>(camera, T_L_C, hash, out_depth_dev, truncation_distance_m, block_size,
                                           maximum_steps, maximum_ray_length_m, surface_distance_epsilon_m,
                                           subsampling);
  CK(cudaDeviceSynchronize());
  hash.destroy();
  return 0;
}

// K5: integrateBlocksKernel<UpdateAppearanceVoxelFunctor<FeatureVoxel>, FeatureVoxel>.
int ref_integrate_features(const float* T_C_L_rm, float fu, float fv, float cu, float cv, int width, int height,
                           const void* feature_dev, const uint8_t* mask_dev, const float* synth_depth_dev,
                           int depth_subsample, float block_size, float max_integration_distance_m,
                           float truncation_distance_m, float max_weight, float measurement_weight,
                           const int* block_idx, const int* slots, int n, void* feature_storage_dev) {
  if (n <= 0) return 0;
  const Transform T_C_L = transform_from_row_major(T_C_L_rm);
  const Camera camera(fu, fv, cu, cv, width, height);
  ImageView<const FeatureArray> feat_view(height, width, static_cast<const FeatureArray*>(feature_dev));
  std::optional<ImageView<const uint8_t>> mask;
  if (mask_dev) mask = ImageView<const uint8_t>(height, width, mask_dev);
  const MaskedFeatureImageConstView feature_frame(feat_view, mask);
  const DepthImageConstView depth_frame(height / depth_subsample, width / depth_subsample, synth_depth_dev);

  using Functor = UpdateAppearanceVoxelFunctor<FeatureVoxel>;
  Functor op_host;
  op_host.truncation_distance_m_ = truncation_distance_m;
  op_host.max_weight_ = max_weight;
  op_host.measurement_weight_ = measurement_weight;
  DevBuf<Functor> op(1);
  CK(cudaMemcpy(op.p, &op_host, sizeof(op_host), cudaMemcpyHostToDevice));

  std::vector<Index3D> idx(n);
  std::vector<VoxelBlock<FeatureVoxel>*> ptrs(n);
  auto* base = static_cast<VoxelBlock<FeatureVoxel>*>(feature_storage_dev);
  for (int i = 0; i < n; ++i) {
    idx[i] = Index3D(block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]);
    ptrs[i] = base + slots[i];
  }
  DevBuf<Index3D> d_idx(n);
  DevBuf<VoxelBlock<FeatureVoxel>*> d_ptrs(n);
  if (upload(d_idx, idx) || upload(d_ptrs, ptrs)) return -1;
  integrateBlocksKernel<Functor, FeatureVoxel><<<n, dim3(8, 8, 8)>>>(
      d_idx.p, camera, feature_frame, depth_frame, T_C_L, block_size, max_integration_distance_m, depth_subsample,
      op.p, d_ptrs.p);
  CK(cudaDeviceSynchronize());
  return 0;
}

// K5 for the colour layer: integrateBlocksKernel<UpdateAppearanceVoxelFunctor<ColorVoxel>, ColorVoxel>.
// color_storage_dev: array of VoxelBlock<ColorVoxel> ({uint8 r,g,b(+pad); float weight} per voxel).
int ref_sizeof_color_voxel() { return (int)sizeof(ColorVoxel); }
int ref_sizeof_color() { return (int)sizeof(Color); }
int ref_integrate_color(const float* T_C_L_rm, float fu, float fv, float cu, float cv, int width, int height,
                        const void* color_dev, const uint8_t* mask_dev, const float* synth_depth_dev,
                        int depth_subsample, float block_size, float max_integration_distance_m,
                        float truncation_distance_m, float max_weight, float measurement_weight,
                        const int* block_idx, const int* slots, int n, void* color_storage_dev) {
  if (n <= 0) return 0;
  const Transform T_C_L = transform_from_row_major(T_C_L_rm);
  const Camera camera(fu, fv, cu, cv, width, height);
  ImageView<const Color> color_view(height, width, static_cast<const Color*>(color_dev));
  std::optional<ImageView<const uint8_t>> mask;
  if (mask_dev) mask = ImageView<const uint8_t>(height, width, mask_dev);
  const MaskedColorImageConstView color_frame(color_view, mask);
  const DepthImageConstView depth_frame(height / depth_subsample, width / depth_subsample, synth_depth_dev);
  using Functor = UpdateAppearanceVoxelFunctor<ColorVoxel>;
  Functor op_host;
  op_host.truncation_distance_m_ = truncation_distance_m;
  op_host.max_weight_ = max_weight;
  op_host.measurement_weight_ = measurement_weight;
  DevBuf<Functor> op(1);
  CK(cudaMemcpy(op.p, &op_host, sizeof(op_host), cudaMemcpyHostToDevice));
  std::vector<Index3D> idx(n);
  std::vector<VoxelBlock<ColorVoxel>*> ptrs(n);
  auto* base = static_cast<VoxelBlock<ColorVoxel>*>(color_storage_dev);
  for (int i = 0; i < n; ++i) {
    idx[i] = Index3D(block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]);
    ptrs[i] = base + slots[i];
  }
  DevBuf<Index3D> d_idx(n);
  DevBuf<VoxelBlock<ColorVoxel>*> d_ptrs(n);
  if (upload(d_idx, idx) || upload(d_ptrs, ptrs)) return -1;
  integrateBlocksKernel<Functor, ColorVoxel><<<n, dim3(8, 8, 8)>>>(
      d_idx.p, camera, color_frame, depth_frame, T_C_L, block_size, max_integration_distance_m, depth_subsample,
      op.p, d_ptrs.p);
  CK(cudaDeviceSynchronize());
  return 0;
}

// K11: updateAppearanceBlockByClosestVoxel<FeatureVoxel> as MeshIntegrator::updateAppearanceGPU launches it
// (NB/src/mesh/mesh_integrator_appearance.cu:255-277: one 256-thread block per mesh block,
// voxel_size = block_size / 8 computed on the host).  vertices_dev: float[n][max_v][3]; counts: vertices per
// block; slots: feature storage slot per block; out_dev: FeatureArray[n][max_v].
int ref_paint_features(const int* block_idx, const int* slots, const int* counts, int n, int max_v,
                       const void* feature_storage_dev, float block_size, float* vertices_dev, void* out_dev) {
  if (n <= 0) return 0;
  using MB = CudaMeshBlock<FeatureArray>;
  const auto* base = static_cast<const VoxelBlock<FeatureVoxel>*>(feature_storage_dev);
  std::vector<const VoxelBlock<FeatureVoxel>*> ptrs(n);
  std::vector<Index3D> idx(n);
  std::vector<MB> mbs(n);
  for (int i = 0; i < n; ++i) {
    ptrs[i] = base + slots[i];
    idx[i] = Index3D(block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]);
    mbs[i].vertices = reinterpret_cast<Vector3f*>(vertices_dev) + (size_t)i * max_v;
    mbs[i].vertex_normals = nullptr;
    mbs[i].triangles = nullptr;
    mbs[i].vertex_appearances = static_cast<FeatureArray*>(out_dev) + (size_t)i * max_v;
    mbs[i].vertices_size = counts[i];
    mbs[i].triangles_size = counts[i];
  }
  DevBuf<const VoxelBlock<FeatureVoxel>*> d_ptrs(n);
  DevBuf<Index3D> d_idx(n);
  DevBuf<MB> d_mb(n);
  if (upload(d_ptrs, ptrs) || upload(d_idx, idx) || upload(d_mb, mbs)) return -1;
  const float voxel_size = block_size / VoxelBlock<TsdfVoxel>::kVoxelsPerSide;
  updateAppearanceBlockByClosestVoxel<FeatureVoxel><<<n, 8 * 32>>>(d_ptrs.p, d_idx.p, block_size, voxel_size, d_mb.p);
  CK(cudaDeviceSynchronize());
  return 0;
}

// K8 + K9: marching cubes over n blocks (no weld: welding only removes duplicates and has no arithmetic
// that could contract, NB/src/mesh/mesh_integrator.cu:690-822).  neighbour_slots[n][8]: storage slot of
// neighbour (dx<<2 | dy<<1 | dz), -1 when absent; entry 0 is the block itself.  Outputs, per block, up to
// max_vertices_per_block vertices (xyz) in the kernel's (atomic-arrival) order, and the vertex count.
int ref_mesh_blocks(const int* block_idx, const int* neighbour_slots, int n, const void* tsdf_storage_dev,
                    float block_size, float voxel_size, float min_weight, int max_vertices_per_block,
                    float* vertices_out_host, float* normals_out_host, int* counts_out_host) {
  if (n <= 0) return 0;
  using MB = CudaMeshBlock<Color>;
  const auto* base = static_cast<const VoxelBlock<TsdfVoxel>*>(tsdf_storage_dev);
  std::vector<const VoxelBlock<TsdfVoxel>*> ptrs(8 * (size_t)n);
  std::vector<Vector3f> pos(n);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 8; ++k)
      ptrs[8 * i + k] = neighbour_slots[8 * i + k] < 0 ? nullptr : base + neighbour_slots[8 * i + k];
    // MeshIntegrator::getTriangleCandidatesAroundBlocks, mesh_integrator.cu:275-276 (host code)
    pos[i] = getPositionFromBlockIndex(block_size, Index3D(block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]));
  }
  DevBuf<const VoxelBlock<TsdfVoxel>*> d_ptrs(ptrs.size());
  DevBuf<Vector3f> d_pos(n);
  DevBuf<marching_cubes::PerVoxelMarchingCubesResults> d_res((size_t)n * 512);
  DevBuf<int> d_sizes(n);
  CK(cudaMemset(d_res.p, 0, (size_t)n * 512 * sizeof(marching_cubes::PerVoxelMarchingCubesResults)));
  CK(cudaMemset(d_sizes.p, 0, n * sizeof(int)));
  if (upload(d_ptrs, ptrs) || upload(d_pos, pos)) return -1;
  const dim3 threads(8, 8, 8);
  meshBlocksCalculateTableIndicesKernel<<<n, threads>>>(n, d_ptrs.p, d_pos.p, voxel_size, min_weight, d_res.p,
                                                        d_sizes.p);
  CK(cudaDeviceSynchronize());
  std::vector<int> sizes(n);
  CK(cudaMemcpy(sizes.data(), d_sizes.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  DevBuf<Vector3f> d_v((size_t)n * max_vertices_per_block), d_n((size_t)n * max_vertices_per_block);
  DevBuf<int> d_t((size_t)n * max_vertices_per_block);
  std::vector<MB> mbs(n);
  for (int i = 0; i < n; ++i) {
    if (sizes[i] > max_vertices_per_block) {
      fprintf(stderr, "ref_mesh_blocks: block %d has %d vertices (> %d)\n", i, sizes[i], max_vertices_per_block);
      return -2;
    }
    mbs[i].vertices = d_v.p + (size_t)i * max_vertices_per_block;
    mbs[i].vertex_normals = d_n.p + (size_t)i * max_vertices_per_block;
    mbs[i].triangles = d_t.p + (size_t)i * max_vertices_per_block;
    mbs[i].vertex_appearances = nullptr;
    mbs[i].vertices_size = sizes[i];
    mbs[i].triangles_size = sizes[i];
  }
  DevBuf<MB> d_mb(n);
  if (upload(d_mb, mbs)) return -1;
  meshBlocksCalculateVerticesKernel<Color><<<n, threads>>>(n, d_res.p, d_sizes.p, d_mb.p);
  CK(cudaDeviceSynchronize());
  static_assert(sizeof(Vector3f) == 12, "packed Vector3f");
  CK(cudaMemcpy(vertices_out_host, d_v.p, (size_t)n * max_vertices_per_block * 12, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(normals_out_host, d_n.p, (size_t)n * max_vertices_per_block * 12, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) counts_out_host[i] = sizes[i];
  return 0;
}

// ---- function-level probes (one thread per sample) -----------------------------------------------------
}  // extern "C"

namespace {
// interpolatePixels<__half> / <float>, NB/include/.../interpolation_2d_impl.h:33-48 (verbatim include)
__global__ void k_interp_half(int n, const float* xy, const __half* f, __half* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  interpolation::internal::interpolatePixels<__half>(Vector2f(xy[2 * i], xy[2 * i + 1]), f[4 * i], f[4 * i + 1],
                                                    f[4 * i + 2], f[4 * i + 3], out + i);
}
__global__ void k_interp_float(int n, const float* xy, const float* f, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  interpolation::internal::interpolatePixels<float>(Vector2f(xy[2 * i], xy[2 * i + 1]), f[4 * i], f[4 * i + 1],
                                                   f[4 * i + 2], f[4 * i + 3], out + i);
}
// blendTwoArrays<FeatureArray>, projective_appearance_integrator.cu:285-305 (extracted)
__global__ void k_blend(int n, const FeatureArray* a, const FeatureArray* b, const float* w, FeatureArray* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  blendTwoArrays(a[i], w[2 * i], b[i], w[2 * i + 1], out + i);
}
// interpolateVertex, marching_cubes_impl.h:28-41 (verbatim include)
__global__ void k_interp_vertex(int n, const float* v1, const float* v2, const float* sdf, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Vector3f r = marching_cubes::interpolateVertex(Vector3f(v1[3 * i], v1[3 * i + 1], v1[3 * i + 2]),
                                                       Vector3f(v2[3 * i], v2[3 * i + 1], v2[3 * i + 2]),
                                                       sdf[2 * i], sdf[2 * i + 1]);
  out[3 * i] = r.x();
  out[3 * i + 1] = r.y();
  out[3 * i + 2] = r.z();
}
// getBlockAndVoxelIndexFromPositionInLayer on the device (the query kernel's / sphere tracer's lookup),
// indexing_impl.h:39-52
__global__ void k_block_voxel(int n, float block_size, const float* p, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Index3D b, v;
  getBlockAndVoxelIndexFromPositionInLayer(block_size, Vector3f(p[3 * i], p[3 * i + 1], p[3 * i + 2]), &b, &v);
  for (int k = 0; k < 3; ++k) {
    out[6 * i + k] = b[k];
    out[6 * i + 3 + k] = v[k];
  }
}
// projectThreadVoxel, projective_integrators_common_impl.cuh:21-55
__global__ void k_project(int n, const Transform T_C_L, const Camera camera, float block_size, float max_depth,
                          const int* bv, float* out, int* ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Eigen::Vector2f u(0.f, 0.f);
  float depth = 0.f;
  Vector3f p_C(0.f, 0.f, 0.f);
  ok[i] = projectThreadVoxel(Index3D(bv[6 * i], bv[6 * i + 1], bv[6 * i + 2]),
                             Index3D(bv[6 * i + 3], bv[6 * i + 4], bv[6 * i + 5]), camera, T_C_L, block_size,
                             max_depth, &u, &depth, &p_C);
  out[6 * i] = u.x();
  out[6 * i + 1] = u.y();
  out[6 * i + 2] = depth;
  out[6 * i + 3] = p_C.x();
  out[6 * i + 4] = p_C.y();
  out[6 * i + 5] = p_C.z();
}
// UpdateTsdfVoxelFunctor on explicit (measured depth, voxel depth, active, old voxel) tuples
__global__ void k_tsdf_functor(int n, UpdateTsdfVoxelFunctor* op, const float* in, const uint8_t* active,
                               float* voxels, uint8_t* updated) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  TsdfVoxel v;
  v.distance = voxels[2 * i];
  v.weight = voxels[2 * i + 1];
  updated[i] = (*op)(in[2 * i], in[2 * i + 1], active[i] != 0, &v);
  voxels[2 * i] = v.distance;
  voxels[2 * i + 1] = v.weight;
}
}  // namespace

extern "C" {
// All pointers below are DEVICE pointers.
int ref_fn_interp_half(int n, const float* xy, const void* f, void* out) {
  k_interp_half<<<(n + 127) / 128, 128>>>(n, xy, static_cast<const __half*>(f), static_cast<__half*>(out));
  CK(cudaDeviceSynchronize());
  return 0;
}
int ref_fn_interp_float(int n, const float* xy, const float* f, float* out) {
  k_interp_float<<<(n + 127) / 128, 128>>>(n, xy, f, out);
  CK(cudaDeviceSynchronize());
  return 0;
}
int ref_fn_blend(int n, const void* a, const void* b, const float* w, void* out) {
  k_blend<<<(n + 63) / 64, 64>>>(n, static_cast<const FeatureArray*>(a), static_cast<const FeatureArray*>(b), w,
                                 static_cast<FeatureArray*>(out));
  CK(cudaDeviceSynchronize());
  return 0;
}
int ref_fn_interp_vertex(int n, const float* v1, const float* v2, const float* sdf, float* out) {
  k_interp_vertex<<<(n + 127) / 128, 128>>>(n, v1, v2, sdf, out);
  CK(cudaDeviceSynchronize());
  return 0;
}
int ref_fn_block_voxel(int n, float block_size, const float* p, int* out) {
  k_block_voxel<<<(n + 127) / 128, 128>>>(n, block_size, p, out);
  CK(cudaDeviceSynchronize());
  return 0;
}
int ref_fn_project(int n, const float* T_C_L_rm, float fu, float fv, float cu, float cv, int width, int height,
                   float block_size, float max_depth, const int* bv, float* out, int* ok) {
  k_project<<<(n + 127) / 128, 128>>>(n, transform_from_row_major(T_C_L_rm), Camera(fu, fv, cu, cv, width, height),
                                      block_size, max_depth, bv, out, ok);
  CK(cudaDeviceSynchronize());
  return 0;
}
int ref_fn_tsdf_functor(int n, float truncation_distance_m, float max_weight, float invalid_depth_decay_factor,
                        int weighting_mode, const float* in, const uint8_t* active, float* voxels, uint8_t* updated) {
  UpdateTsdfVoxelFunctor op_host;
  op_host.truncation_distance_m_ = truncation_distance_m;
  op_host.max_weight_ = max_weight;
  op_host.invalid_depth_decay_factor_ = invalid_depth_decay_factor;
  op_host.weighting_function_ = WeightingFunction(static_cast<WeightingFunctionType>(weighting_mode));
  DevBuf<UpdateTsdfVoxelFunctor> op(1);
  CK(cudaMemcpy(op.p, &op_host, sizeof(op_host), cudaMemcpyHostToDevice));
  k_tsdf_functor<<<(n + 127) / 128, 128>>>(n, op.p, in, active, voxels, updated);
  CK(cudaDeviceSynchronize());
  return 0;
}
}  // extern "C"
