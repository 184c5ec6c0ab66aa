// nvbx_kernels.cuh -- the sm_100a kernels of the reconstruction hot path.
//
// Design rules:
//   * every kernel whose work size is only known on the device (view lists, band lists, slot table)
//     is PERSISTENT: grid = SMs x resident CTAs, CTAs grid-stride over a device-side list whose
//     length they read from HBM.  No kernel launch depends on a host read-back, so a frame is a
//     fixed sequence of launches with zero host synchronisation.
//   * voxel payloads are laid out so that a warp touches contiguous bytes: TSDF float2[512] with z
//     fastest, feature rows of C+8 halves (16-byte aligned) moved with 128-bit loads/stores.
//   * the path is HBM-bound byte shuffling; there is no GEMM shape in it and no tensor-core use.
//   * all fp32 geometry follows nvbx_math.cuh's operation order; fp16 uses the *_rn intrinsics so
//     ptxas cannot contract mul+add.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nvbx_map.cuh"

#define NVBX_MC_QUAL static __device__ const
#include "mc_tables.h"

namespace nvbx {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ void count_add(const MapDev& m, int id, unsigned long long v) {
  atomicAdd(&m.ctrl->counters[id], v);
}

// ================================================================================================
// a2. Blocks in view by ray casting -- one thread per (subsampled) depth pixel, 3-D DDA into a dense
// byte grid over the view AABB.  Follows combinedBlockIndicesInImageKernel
// (NB/src/integrators/view_calculator.cu:196-248) and RayCaster (ray_caster_impl.h:26-75), including
// the linear-index alias of setIndexUpdated (:171-180).
// ================================================================================================
struct ViewGrid {
  I3 mn;
  int sx, sy, sz;
  int n_cells;
};

__device__ __forceinline__ void mark_cell(uint8_t* grid, const ViewGrid& g, int x, int y, int z) {
  const int lx = x - g.mn.x, ly = y - g.mn.y, lz = z - g.mn.z;
  const size_t lin = (size_t)(int)(lx + ly * g.sx + lz * g.sx * g.sy);
  if (lin < (size_t)g.n_cells) {
    if (!grid[lin]) grid[lin] = 1;  // benign same-value race, as in the reference
  }
}

__global__ void __launch_bounds__(256) k_raycast_mark(Pose T_L_C, Cam cam, const float* __restrict__ depth, int rows,
                                                      int cols, float block_size, float max_dist, float behind,
                                                      int sub, ViewGrid g, uint8_t* grid) {
  const int ray_col = blockIdx.x * blockDim.x + threadIdx.x;
  const int ray_row = blockIdx.y * blockDim.y + threadIdx.y;
  int prow = ray_row * sub, pcol = ray_col * sub;
  if (prow >= rows + sub - 1 || pcol >= cols + sub - 1) return;
  if (prow >= rows) prow = rows - 1;
  if (pcol >= cols) pcol = cols - 1;
  float d = depth[(size_t)prow * cols + pcol];
  if (d <= 0.0f) return;
  if (max_dist > 0.0f && d > max_dist) d = max_dist;
  const V3 ray = ray_from_image_plane(cam, (float)pcol + 0.5f, (float)prow + 0.5f);
  const float len = d + behind;
  V3 p_C;
  p_C.x = len * ray.x;
  p_C.y = len * ray.y;
  p_C.z = len * ray.z;
  const V3 p_L = xform(T_L_C, p_C);
  const I3 b = block_index_from_position(block_size, p_L);
  mark_cell(grid, g, b.x, b.y, b.z);

  // RayCaster(origin / bs, p_L / bs), scale 1
  const float s[3] = {(T_L_C.t[0] / block_size) / 1.0f, (T_L_C.t[1] / block_size) / 1.0f,
                      (T_L_C.t[2] / block_size) / 1.0f};
  const float e[3] = {(p_L.x / block_size) / 1.0f, (p_L.y / block_size) / 1.0f, (p_L.z / block_size) / 1.0f};
  int cur[3], sign[3];
  float t_next[3], t_step[3];
  int steps = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    cur[i] = (int)floorf(s[i]);
    const int end = (int)floorf(e[i]);
    steps += abs(end - cur[i]);
    const float r = e[i] - s[i];
    sign[i] = (r > 0.0f) ? 1 : ((r < 0.0f) ? -1 : 0);
    const int corrected = max(sign[i], 0);
    const float shifted = s[i] - (float)cur[i];
    const float dist = (float)corrected - shifted;
    t_next[i] = dist / r;
    t_step[i] = (float)sign[i] / r;
  }
  for (int step = 0; step <= steps; ++step) {
    mark_cell(grid, g, cur[0], cur[1], cur[2]);
    int mi = 0;
    float mv = t_next[0];
    if (t_next[1] < mv) {
      mi = 1;
      mv = t_next[1];
    }
    if (t_next[2] < mv) {
      mi = 2;
    }
    // branch-free select keeps cur/t_next in registers
    if (mi == 0) {
      cur[0] += sign[0];
      t_next[0] += t_step[0];
    } else if (mi == 1) {
      cur[1] += sign[1];
      t_next[1] += t_step[1];
    } else {
      cur[2] += sign[2];
      t_next[2] += t_step[2];
    }
  }
}

// Number of marked cells (only launched while the block arena is still growing, see ensure_slots()).
__global__ void __launch_bounds__(256) k_count_marked(const uint8_t* __restrict__ grid, int n_cells, int* out) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) c += grid[i] ? 1 : 0;
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane_id() == 0 && c) atomicAdd(out, c);
}

// Zero one TSDF payload with the whole warp (4 KiB = 8 x 32 x 16 B).
__device__ __forceinline__ void warp_zero_tsdf(const MapDev& m, int slot) {
  uint4* p = reinterpret_cast<uint4*>(tsdf_block(m, slot));
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i * 32 + lane_id()] = z;
}

// Give `slot` a TSDF layer if it has none; returns true if the payload must be zeroed.
__device__ __forceinline__ bool ensure_tsdf_layer(const MapDev& m, int slot) {
  const uint8_t layers = m.blk_layers[slot];
  if (layers & kLayerTsdfBit) return false;
  m.blk_layers[slot] = layers | kLayerTsdfBit;
  atomicAdd(&m.ctrl->n_tsdf, 1);
  return true;
}

// Scan the marked grid -> block indices (appended to the viewpoint-cache entry), find-or-allocate each
// block in the map, emit the slot list for the TSDF update.  Replaces the D2H copy + CPU scan + CPU
// allocation + H2D pointer tables of view_calculator.cu:313-323 / layer_impl.h:130-162 /
// integrators_common_impl.h:60-121.
__global__ void __launch_bounds__(256) k_view_compact_alloc(MapDev m, const uint8_t* __restrict__ grid, ViewGrid g,
                                                            int3* entry_idx, int* entry_count, int* view_slots) {
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = lane_id();
  for (int base = warp * 32; base < g.n_cells; base += warps_total * 32) {
    const int i = base + lane;
    const bool marked = (i < g.n_cells) && grid[i];
    const unsigned ballot = __ballot_sync(0xffffffffu, marked);
    if (!ballot) continue;
    int pos0 = 0;
    if (lane == 0) pos0 = atomicAdd(entry_count, __popc(ballot));
    pos0 = __shfl_sync(0xffffffffu, pos0, 0);
    int slot = -1;
    bool zero = false;
    if (marked) {
      const int x = i % g.sx + g.mn.x;
      const int y = (i / g.sx) % g.sy + g.mn.y;
      const int z = i / (g.sx * g.sy) + g.mn.z;
      const int pos = pos0 + __popc(ballot & ((1u << lane) - 1u));
      entry_idx[pos] = make_int3(x, y, z);
      bool is_new;
      slot = acquire_slot(m, x, y, z, &is_new);
      if (slot >= 0) zero = ensure_tsdf_layer(m, slot);
      view_slots[pos] = slot;
    }
    unsigned zb = __ballot_sync(0xffffffffu, zero);
    if (zb && lane == 0) count_add(m, kCntTsdfBlocksAllocated, __popc(zb));
    while (zb) {
      const int src = __ffs(zb) - 1;
      zb &= zb - 1;
      warp_zero_tsdf(m, __shfl_sync(0xffffffffu, slot, src));
    }
  }
}

// Viewpoint-cache hit: the cached index list is re-used as is (view_calculator.cu:256-265) and blocks
// are (re-)allocated where required (projective_integrator_impl.cuh:288-291).
__global__ void __launch_bounds__(256) k_view_alloc_from_list(MapDev m, const int3* __restrict__ entry_idx,
                                                              const int* __restrict__ entry_count, int* view_slots) {
  const int n = *entry_count;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = lane_id();
  for (int base = warp * 32; base < n; base += warps_total * 32) {
    const int i = base + lane;
    int slot = -1;
    bool zero = false;
    if (i < n) {
      const int3 b = entry_idx[i];
      bool is_new;
      slot = acquire_slot(m, b.x, b.y, b.z, &is_new);
      if (slot >= 0) zero = ensure_tsdf_layer(m, slot);
      view_slots[i] = slot;
    }
    unsigned zb = __ballot_sync(0xffffffffu, zero);
    if (zb && lane == 0) count_add(m, kCntTsdfBlocksAllocated, __popc(zb));
    while (zb) {
      const int src = __ffs(zb) - 1;
      zb &= zb - 1;
      warp_zero_tsdf(m, __shfl_sync(0xffffffffu, slot, src));
    }
  }
}

// ================================================================================================
// a4. TSDF update.  integrateBlocksKernel<TsdfVoxel> (projective_integrator_impl.cuh:58-103) +
// UpdateTsdfVoxelFunctor (projective_tsdf_integrator.cu:25-99).  One CTA of 512 threads per block,
// thread t owns voxel t of the [x][y][z] array (z fastest -> 32 lanes read 256 contiguous bytes).
// ================================================================================================
struct DepthFrame {
  const float* depth;
  const uint8_t* mask;  // may be null
  int rows, cols;
  Cam cam;
  Pose T_C_L;
  float max_depth;
  float trunc;
  float max_weight;
  float invalid_decay;
  int weighting_mode;
};

__device__ __forceinline__ bool project_voxel(const Cam& cam, const Pose& T_C_L, float block_size, float max_depth,
                                              const int3 b, int vx, int vy, int vz, float* u, float* v, float* vd) {
  I3 bi;
  bi.x = b.x;
  bi.y = b.y;
  bi.z = b.z;
  const V3 pl = voxel_center(block_size, bi, vx, vy, vz);
  const V3 pc = xform(T_C_L, pl);
  if (!project(cam, pc, u, v)) return false;
  *vd = pc.z;
  if (max_depth > 0.0f && *vd > max_depth) return false;
  return true;
}

__global__ void __launch_bounds__(512, 2) k_tsdf_update(MapDev m, const int* __restrict__ view_slots,
                                                        const int* __restrict__ view_count, DepthFrame f) {
  const int n = *view_count;
  const int t = threadIdx.x;
  const int vx = t >> 6, vy = (t >> 3) & 7, vz = t & 7;
  unsigned updated = 0;
  for (int bi = blockIdx.x; bi < n; bi += gridDim.x) {
    const int slot = view_slots[bi];
    if (slot < 0) continue;
    if (t == 0) m.blk_dirty[slot] = 1;  // blocks_to_update_tracker_.addBlocksToUpdate (mapper.cpp:406)
    const int3 b = m.blk_index[slot];
    float u, v, vd;
    if (!project_voxel(f.cam, f.T_C_L, m.block_size, f.max_depth, b, vx, vy, vz, &u, &v, &vd)) continue;
    const int ui = (int)floorf(u), vi = (int)floorf(v);
    if (ui < 0 || vi < 0 || ui >= f.cols || vi >= f.rows) continue;
    const float meas = __ldg(f.depth + (size_t)vi * f.cols + ui);
    if (isnan(meas)) continue;
    const bool active = (f.mask == nullptr) || __ldg(f.mask + (size_t)vi * f.cols + ui);
    float2* vox = tsdf_block(m, slot) + t;
    if (meas <= 0.0f) {
      if (f.invalid_decay >= 0.0f) vox->y *= f.invalid_decay;
      continue;
    }
    const float sdf = meas - vd;
    if (sdf < -f.trunc) continue;
    if (!active && sdf < f.trunc) continue;
    const float2 cur = *vox;
    const float w_m = weighting(f.weighting_mode, meas, vd, f.trunc);
    float fused = (sdf * w_m + cur.x * cur.y) / (w_m + cur.y);
    if (fused > 0.0f)
      fused = fminf(f.trunc, fused);
    else
      fused = fmaxf(-f.trunc, fused);
    const float w_new = fminf(w_m + cur.y, f.max_weight);
    *vox = make_float2(fused, w_new);
    ++updated;
  }
  // accounting: one atomic per warp
  for (int o = 16; o; o >>= 1) updated += __shfl_xor_sync(0xffffffffu, updated, o);
  if (lane_id() == 0 && updated) count_add(m, kCntTsdfVoxelsUpdated, updated);
  if (blockIdx.x == 0 && t == 0) {
    count_add(m, kCntTsdfBlocksInView, (unsigned long long)n);
    count_add(m, kCntDepthFrames, 1);
  }
}

// ================================================================================================
// a6. Feature candidates.  getBlocksInViewPlanes (view_calculator.cu:392-470) on the device: one thread
// per block of the frustum AABB, block-centre test against the normalised viewport (+10 px margin).
// The reference enumerates x-outer / z-inner; the order of the list is irrelevant here.
// ================================================================================================
__global__ void __launch_bounds__(256) k_planes_view(ViewGrid g, float block_size, Pose T_C_L, float vmin_x,
                                                     float vmin_y, float vmax_x, float vmax_y, int3* entry_idx,
                                                     int* entry_count) {
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < g.n_cells; base += stride) {
    const int i = base + lane_id();
    bool in_view = false;
    int x = 0, y = 0, z = 0;
    if (i < g.n_cells) {
      x = i % g.sx + g.mn.x;
      y = (i / g.sx) % g.sy + g.mn.y;
      z = i / (g.sx * g.sy) + g.mn.z;
      V3 c;
      c.x = block_size * ((float)x + 0.5f);
      c.y = block_size * ((float)y + 0.5f);
      c.z = block_size * ((float)z + 0.5f);
      const V3 r = rotate(T_C_L, c);
      const float px = r.x + T_C_L.t[0], py = r.y + T_C_L.t[1], pz = r.z + T_C_L.t[2];
      if (pz > 1e-6f) {
        const float un = px / pz, vn = py / pz;
        in_view = (vmin_x <= un) && (vmin_y <= vn) && (un <= vmax_x) && (vn <= vmax_y);
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, in_view);
    if (!ballot) continue;
    int pos0 = 0;
    if (lane_id() == 0) pos0 = atomicAdd(entry_count, __popc(ballot));
    pos0 = __shfl_sync(0xffffffffu, pos0, 0);
    if (in_view) entry_idx[pos0 + __popc(ballot & ((1u << lane_id()) - 1u))] = make_int3(x, y, z);
  }
}

// reduceBlocksToThoseInTruncationBand (projective_appearance_integrator.cu:374-477) + feature block
// allocation (:120-123), one warp per candidate: hash probe, 4 KiB band scan with early exit,
// feature-slot pop.  Emits the band list and the list of freshly allocated feature slots.
__global__ void __launch_bounds__(256) k_band_select(MapDev m, const int3* __restrict__ entry_idx,
                                                     const int* __restrict__ entry_count, float trunc,
                                                     int* band_slots, int* newfeat_slots) {
  const int n = *entry_count;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = lane_id();
  unsigned cand = 0;
  for (int i = warp; i < n; i += warps_total) {
    int slot = -1;
    if (lane == 0) {
      const int3 b = entry_idx[i];
      slot = hash_find(m, b.x, b.y, b.z);
      if (slot >= 0 && !(m.blk_layers[slot] & kLayerTsdfBit)) slot = -1;
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < 0) continue;
    ++cand;
    const float4* p = reinterpret_cast<const float4*>(tsdf_block(m, slot));
    bool in_band = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 q = p[k * 32 + lane];  // two voxels: (d, w, d, w)
      const bool hit = (q.y > 0.0f && fabsf(q.x) < trunc) || (q.w > 0.0f && fabsf(q.z) < trunc);
      if (__any_sync(0xffffffffu, hit)) {
        in_band = true;
        break;
      }
    }
    if (!in_band) continue;
    if (lane == 0) {
      if (m.blk_feat[slot] < 0) {
        const int fs = pop_id(&m.ctrl->feat_free_top, &m.ctrl->feat_high, m.feat_free, m.feat_capacity,
                              &m.ctrl->overflow);
        if (fs >= 0) {
          m.blk_feat[slot] = fs;
          m.blk_layers[slot] |= kLayerFeatBit;
          atomicAdd(&m.ctrl->n_feat, 1);
          newfeat_slots[atomicAdd(&m.ctrl->newfeat_count, 1)] = fs;
          count_add(m, kCntFeatBlocksAllocated, 1);
        }
      }
      if (m.blk_feat[slot] >= 0) band_slots[atomicAdd(&m.ctrl->band_count, 1)] = slot;
    }
  }
  if (lane == 0 && cand) count_add(m, kCntFeatCandidateBlocks, cand);
}

// Zero-fill freshly allocated feature blocks (the reference memsets every new block:
// NB/include/nvblox/map/internal/impl/blox_impl.h:92-97).
__global__ void __launch_bounds__(256) k_zero_feature_blocks(MapDev m, const int* __restrict__ newfeat_slots) {
  const int n = m.ctrl->newfeat_count;
  const int vec_per_block = (kVoxelsPerBlock * m.row) / 8;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    uint4* p = reinterpret_cast<uint4*>(feat_block(m, newfeat_slots[i]));
    for (int k = threadIdx.x; k < vec_per_block; k += blockDim.x) p[k] = z;
  }
}

// ================================================================================================
// a7. Synthetic depth by sphere tracing.  sphereTracingKernel + cast (sphere_tracer.cu:31-131,191-236).
// One thread per ray; the last block hit is cached in registers so consecutive samples inside one
// block cost no hash probe.
// ================================================================================================
struct TraceParams {
  Cam cam;
  Pose T_L_C;
  float trunc;
  int max_steps;
  float max_ray_length;
  float eps;
  int sub;
  int rows, cols;  // synthetic image size
};

__global__ void __launch_bounds__(64) k_sphere_trace(MapDev m, TraceParams tp, float* __restrict__ image) {
  const int c = threadIdx.x % 8 + blockIdx.x * 8;
  const int r = threadIdx.x / 8 + blockIdx.y * 8;
  if (r >= tp.rows || c >= tp.cols) return;
  const float pu = (float)(c * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const float pv = (float)(r * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const V3 ray = ray_from_image_plane(tp.cam, pu, pv);
  const float sq = sum3(ray.x * ray.x, ray.y * ray.y, ray.z * ray.z);
  V3 dc = ray;
  if (sq > 0.0f) {
    const float nrm = sqrtf(sq);
    dc.x = ray.x / nrm;
    dc.y = ray.y / nrm;
    dc.z = ray.z / nrm;
  }
  const V3 dl = rotate(tp.T_L_C, dc);
  const float ox = tp.T_L_C.t[0], oy = tp.T_L_C.t[1], oz = tp.T_L_C.t[2];

  int first = 0;
  float t = 0.0f;
  bool ok = false;
  I3 cached_b;
  cached_b.x = cached_b.y = cached_b.z = 0x7fffffff;
  const float2* cached_ptr = nullptr;
  for (int i = 0; (i < tp.max_steps) && (t < tp.max_ray_length); ++i) {
    V3 p;
    p.x = ox + t * dl.x;
    p.y = oy + t * dl.y;
    p.z = oz + t * dl.z;
    I3 b, v;
    block_and_voxel_from_position(m.block_size, m.voxel_size_inv, p, &b, &v);
    if (b.x != cached_b.x || b.y != cached_b.y || b.z != cached_b.z) {
      cached_b = b;
      const int slot = hash_find(m, b.x, b.y, b.z);
      cached_ptr = (slot >= 0 && (m.blk_layers[slot] & kLayerTsdfBit)) ? tsdf_block(m, slot) : nullptr;
    }
    bool valid = false;
    float dist = 0.0f;
    if (cached_ptr) {
      const float2 q = cached_ptr[(v.x * 8 + v.y) * 8 + v.z];
      if (q.y > 1e-4f) {
        valid = true;
        dist = q.x;
      }
    }
    float step;
    if (!valid) {
      if (first == 0) {
        step = tp.trunc;
      } else {
        break;  // left observed space: fail
      }
    } else {
      if (first == 0) first = (dist >= 0.0f) ? 1 : -1;
      if (first == 1) {
        if (dist < tp.eps) {
          t += dist;
          ok = true;
          break;
        }
        step = dist;
      } else {
        if (dist > -tp.eps) {
          t -= dist;
          ok = true;
          break;
        }
        step = -dist;
      }
    }
    t += step;
  }
  image[(size_t)r * tp.cols + c] = ok ? t * dc.z : -1.0f;
}

// ================================================================================================
// a8. Feature integration -- THE hot kernel.
//
// Reference: integrateBlocksKernel<UpdateAppearanceVoxelFunctor<FeatureVoxel>> (projective_integrator_
// impl.cuh:156-214) runs one THREAD per voxel that serially walks all C channels through three
// 1.5 KB per-thread arrays and a 2-byte-aligned 1538-byte AoS record.  Here one CTA owns one voxel
// block:
//   phase 1 (thread = voxel): geometry only -- project, bilinear synthetic depth, band test, image
//     bounds, mask, old weight -> a compacted list of work items in shared memory;
//   phase 2 (warp = work item): the 32 lanes sweep the C channels as 16-byte vectors: four 128-bit
//     read-only loads per vector (the 4 bilinear neighbours, each C contiguous halves in the HWC
//     image), fp16 interpolation/blend in the reference's operation order on half2 lanes, one 128-bit
//     store into the voxel's 16-byte-aligned row.  Voxels adjacent in z are consecutive work items and
//     mostly share pixel rows, so re-reads hit L1/L2, not HBM.
// Algorithmic bytes per updated voxel: 2C x (distinct pixels, <= 4) read + 2(C+8) written
// (+ 2C read when the old feature must be blended).
// ================================================================================================
struct FeatFrame {
  const __half* img;
  const uint8_t* mask;  // may be null
  const float* synth;
  int rows, cols;
  int srows, scols;
  int sub;
  Cam cam;
  Pose T_C_L;
  float max_depth;
  float trunc;
  float alpha;
  float max_weight;
  unsigned short h_w1, h_w2;  // half bits of (1-alpha)/(total), alpha/(total)
  int read_old;               // 1: blend with the stored feature (alpha < 1 or strict mode)
};

__device__ __forceinline__ __half2 interp_h2(__half2 x, __half2 y, __half2 xy, __half2 f00, __half2 f01, __half2 f10,
                                             __half2 f11) {
  const __half2 dx = __hsub2(f10, f00);
  const __half2 t2 = __hadd2_rn(f00, __hmul2_rn(x, dx));
  const __half2 t5 = __hadd2_rn(t2, __hmul2_rn(y, __hsub2(f01, f00)));
  const __half2 t9 = __hmul2_rn(xy, __hsub2(__hsub2(f11, f01), dx));
  return __hadd2_rn(t5, t9);
}
__device__ __forceinline__ uint4 interp_vec(__half2 x, __half2 y, __half2 xy, uint4 a00, uint4 a01, uint4 a10,
                                            uint4 a11) {
  uint4 o;
  const __half2* p00 = reinterpret_cast<const __half2*>(&a00);
  const __half2* p01 = reinterpret_cast<const __half2*>(&a01);
  const __half2* p10 = reinterpret_cast<const __half2*>(&a10);
  const __half2* p11 = reinterpret_cast<const __half2*>(&a11);
  __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) po[i] = interp_h2(x, y, xy, p00[i], p01[i], p10[i], p11[i]);
  return o;
}
__device__ __forceinline__ uint4 blend_vec(uint4 oldv, uint4 meas, __half2 w1, __half2 w2) {
  uint4 o;
  const __half2* po = reinterpret_cast<const __half2*>(&oldv);
  const __half2* pm = reinterpret_cast<const __half2*>(&meas);
  __half2* pr = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hadd2_rn(__hmul2_rn(po[i], w1), __hmul2_rn(pm[i], w2));
  return o;
}
__device__ __forceinline__ uint4 ldg_nc(const uint4* p) { return __ldg(p); }

struct FeatItem {
  int pix;                // (ly * cols + lx): offset of the top-left neighbour in pixels
  unsigned short hx, hy;  // half bits of the interpolation offsets
  unsigned short wnew;    // half bits of the new weight
  unsigned short vox;     // voxel linear id | (first-observation flag << 15)
};

template <int VPL>  // 16-byte vectors per lane (C = 256 * VPL); 0 = any C (multiple of 8)
__global__ void __launch_bounds__(512, 2) k_feature_integrate(MapDev m, const int* __restrict__ band_slots,
                                                              FeatFrame f) {
  __shared__ FeatItem s_items[kVoxelsPerBlock];
  __shared__ int s_warp_base[17];
  const int n = m.ctrl->band_count;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int vx = t >> 6, vy = (t >> 3) & 7, vz = t & 7;
  const int C = m.C;
  const int nvec = C >> 3;
  const __half2 w1 = __half2half2(__ushort_as_half(f.h_w1));
  const __half2 w2 = __half2half2(__ushort_as_half(f.h_w2));
  unsigned long long n_updated = 0;

  for (int bi = blockIdx.x; bi < n; bi += gridDim.x) {
    const int slot = band_slots[bi];
    const int3 b = m.blk_index[slot];
    __half* blk = feat_block(m, m.blk_feat[slot]);
    if (t == 0) m.blk_dirty[slot] = 1;  // mapper.cpp:462

    // ---- phase 1: geometry, one thread per voxel -------------------------------------------------
    bool active = false;
    FeatItem it;
    it.pix = 0;
    it.hx = it.hy = it.wnew = it.vox = 0;
    do {
      float u, v, vd;
      if (!project_voxel(f.cam, f.T_C_L, m.block_size, f.max_depth, b, vx, vy, vz, &u, &v, &vd)) break;
      // occlusion test against the synthetic depth (bilinear, no validity check)
      const float ud = u / (float)f.sub, vdp = v / (float)f.sub;
      const float uc = ud - 0.5f, vc = vdp - 0.5f;
      const int lx = (int)floorf(uc), ly = (int)floorf(vc);
      if (lx < 0 || ly < 0 || (lx + 1) > (f.scols - 1) || (ly + 1) > (f.srows - 1)) break;
      const float* sp = f.synth + (size_t)ly * f.scols + lx;
      const float surface = interp_float(uc - (float)lx, vc - (float)ly, sp[0], sp[f.scols], sp[1], sp[f.scols + 1]);
      if (fabsf(surface - vd) > f.trunc) break;
      const float fu = u - 0.5f, fv = v - 0.5f;
      const int px = (int)floorf(fu), py = (int)floorf(fv);
      if (px < 0 || py < 0 || (px + 1) > (f.cols - 1) || (py + 1) > (f.rows - 1)) break;
      if (f.mask != nullptr && !__ldg(f.mask + (size_t)((int)v) * f.cols + (int)u)) break;
      const __half w_cur_h = blk[(size_t)t * m.row + C];
      const float w_cur = __half2float(w_cur_h);
      it.pix = py * f.cols + px;
      it.hx = __half_as_ushort(__float2half_rn(fu - (float)px));
      it.hy = __half_as_ushort(__float2half_rn(fv - (float)py));
      it.wnew = __half_as_ushort(__float2half_rn(fminf(f.alpha + w_cur, f.max_weight)));
      it.vox = (unsigned short)(t | ((w_cur == 0.0f) ? 0x8000 : 0));
      active = true;
    } while (false);

    // ---- compaction (ballot + 16-entry scan) ----------------------------------------------------
    const unsigned ballot = __ballot_sync(0xffffffffu, active);
    if (lane == 0) s_warp_base[warp + 1] = __popc(ballot);
    __syncthreads();
    if (t == 0) {
      int acc = 0;
      s_warp_base[0] = 0;
#pragma unroll
      for (int w = 1; w <= 16; ++w) {
        acc += s_warp_base[w];
        s_warp_base[w] = acc;
      }
    }
    __syncthreads();
    if (active) s_items[s_warp_base[warp] + __popc(ballot & ((1u << lane) - 1u))] = it;
    const int n_items = s_warp_base[16];
    __syncthreads();

    // ---- phase 2: one warp per work item, lanes sweep the channels ----------------------------------
    for (int i = warp; i < n_items; i += 16) {
      const FeatItem w = s_items[i];
      const int vox = w.vox & 0x1ff;
      const bool first = (w.vox & 0x8000) != 0;
      const __half hx = __ushort_as_half(w.hx), hy = __ushort_as_half(w.hy);
      const __half2 x2 = __half2half2(hx), y2 = __half2half2(hy), xy2 = __half2half2(__hmul_rn(hx, hy));
      const uint4* p00 = reinterpret_cast<const uint4*>(f.img + (size_t)w.pix * C);
      const uint4* p10 = p00 + nvec;
      const uint4* p01 = p00 + (size_t)f.cols * nvec;
      const uint4* p11 = p01 + nvec;
      uint4* dst = reinterpret_cast<uint4*>(blk + (size_t)vox * m.row);
      const bool blend = (!first) && f.read_old;
      if (VPL > 0) {
        uint4 a00[VPL > 0 ? VPL : 1], a01[VPL > 0 ? VPL : 1], a10[VPL > 0 ? VPL : 1], a11[VPL > 0 ? VPL : 1];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {  // all loads first: 4*VPL 128-bit requests in flight per lane
          const int c = lane + 32 * k;
          a00[k] = ldg_nc(p00 + c);
          a10[k] = ldg_nc(p10 + c);
          a01[k] = ldg_nc(p01 + c);
          a11[k] = ldg_nc(p11 + c);
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int c = lane + 32 * k;
          uint4 o = interp_vec(x2, y2, xy2, a00[k], a01[k], a10[k], a11[k]);
          if (blend) o = blend_vec(dst[c], o, w1, w2);
          dst[c] = o;
        }
      } else {
        for (int c = lane; c < nvec; c += 32) {
          uint4 o = interp_vec(x2, y2, xy2, ldg_nc(p00 + c), ldg_nc(p01 + c), ldg_nc(p10 + c), ldg_nc(p11 + c));
          if (blend) o = blend_vec(dst[c], o, w1, w2);
          dst[c] = o;
        }
      }
      if (lane == 0) dst[nvec] = make_uint4((unsigned)w.wnew, 0u, 0u, 0u);  // weight + zero padding
    }
    if (t == 0) n_updated += (unsigned long long)n_items;
    __syncthreads();  // s_items / s_warp_base are reused by the next block
  }
  if (t == 0) {
    if (n_updated) count_add(m, kCntFeatVoxelsUpdated, n_updated);
    if (blockIdx.x == 0) {
      count_add(m, kCntFeatBandBlocks, (unsigned long long)n);
      count_add(m, kCntFeatureFrames, 1);
    }
  }
}

// ================================================================================================
// a9. Decay.  decayKernel + TsdfDecayFunctor (decayer_impl.cuh:83-125, tsdf_decay_integrator.cu:58-112)
// fused with deallocateFullyDecayedBlocks (:259-274) and Mapper::clearBlocksInLayers
// (mapper.cpp:761-849): a fully decayed block releases its TSDF slot, feature slot and mesh extent on
// the device.  256 threads per block, two voxels (one float4) per thread.
// ================================================================================================
struct DecayParams {
  float factor;
  float threshold;
  int set_free;
  float free_distance;
  int deallocate;
};

__global__ void __launch_bounds__(256) k_decay(MapDev m, DecayParams dp) {
  const int n = m.ctrl->slot_high;
  const float lo = dp.threshold - 1e-6f;
  const float hi = dp.threshold + 1e-6f;
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    const uint8_t layers = m.blk_layers[slot];
    if (!(layers & kLayerTsdfBit)) continue;  // uniform per CTA
    float4* p = reinterpret_cast<float4*>(tsdf_block(m, slot)) + threadIdx.x;
    float4 q = *p;
    bool touched = false;
    if (!(q.y < lo)) {
      q.y = fmaxf(q.y * dp.factor, dp.threshold);
      if (dp.set_free && q.y < hi) q.x = dp.free_distance;
      touched = true;
    }
    if (!(q.w < lo)) {
      q.w = fmaxf(q.w * dp.factor, dp.threshold);
      if (dp.set_free && q.w < hi) q.z = dp.free_distance;
      touched = true;
    }
    if (touched) *p = q;
    const bool decayed = (q.y < hi) && (q.w < hi);
    const int all = __syncthreads_and(decayed);
    if (threadIdx.x == 0) {
      if (all && dp.deallocate) {
        // release everything hanging off this block index
        m.blk_layers[slot] = 0;
        atomicAdd(&m.ctrl->n_tsdf, -1);
        const int fs = m.blk_feat[slot];
        if (fs >= 0) {
          push_id(&m.ctrl->feat_free_top, m.feat_free, fs);
          atomicAdd(&m.ctrl->n_feat, -1);
          m.blk_feat[slot] = -1;
        }
        m.blk_mesh[slot] = make_int4(0, 0, 0, 0);
        m.blk_dirty[slot] = 0;
        push_id(&m.ctrl->slot_free_top, m.slot_free, slot);
        m.ctrl->rebuild = 1;
        count_add(m, kCntBlocksDeallocated, 1);
      } else {
        m.blk_dirty[slot] = 1;  // decayTsdf marks every TSDF block "to update" (mapper.cpp:469-471)
      }
    }
  }
}

// Hash rebuild after releases (three tiny launches, all no-ops unless ctrl->rebuild is set).
__global__ void __launch_bounds__(256) k_hash_clear(MapDev m, int force) {
  if (!force && !m.ctrl->rebuild) return;
  const unsigned n = m.hash_mask + 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m.keys[i] = kEmptyKey;
}
__global__ void __launch_bounds__(256) k_hash_reinsert(MapDev m, int force) {
  if (!force && !m.ctrl->rebuild) return;
  const int n = m.ctrl->slot_high;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (m.blk_layers[s]) {
      const int3 b = m.blk_index[s];
      hash_insert(m, b.x, b.y, b.z, s);
    }
  }
}
__global__ void k_hash_rebuild_done(MapDev m) { m.ctrl->rebuild = 0; }

// Mapper::clear (py_mapper.cu:286-306): drop every block of every layer.
__global__ void __launch_bounds__(256) k_clear_all(MapDev m) {
  const unsigned n = m.hash_mask + 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m.keys[i] = kEmptyKey;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    Ctrl* c = m.ctrl;
    c->slot_free_top = 0;
    c->slot_high = 0;
    c->feat_free_top = 0;
    c->feat_high = 0;
    c->n_tsdf = 0;
    c->n_feat = 0;
    c->rebuild = 0;
    c->mesh_total_v = 0;
    c->mesh_total_t = 0;
  }
}

// ================================================================================================
// a13. Layer views and point queries.
// ================================================================================================
__global__ void __launch_bounds__(256) k_collect_block_indices(MapDev m, uint8_t layer_bit, int3* out, int capacity) {
  const int n = m.ctrl->slot_high;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (m.blk_layers[s] & layer_bit) {
      const int pos = atomicAdd(&m.ctrl->list_count, 1);
      if (pos < capacity) out[pos] = m.blk_index[s];
    }
  }
}

// single-thread helpers for allocate_block_at_index / get_block_at_index
__global__ void k_allocate_one(MapDev m, int x, int y, int z, int layer, int* newfeat_slot_out) {
  bool is_new;
  const int slot = acquire_slot(m, x, y, z, &is_new);
  *newfeat_slot_out = -1;
  if (slot < 0) return;
  if (layer == 0) {
    if (ensure_tsdf_layer(m, slot)) {
      float2* p = tsdf_block(m, slot);
      for (int i = 0; i < kVoxelsPerBlock; ++i) p[i] = make_float2(0.f, 0.f);
    }
  } else if (m.blk_feat[slot] < 0) {
    const int fs =
        pop_id(&m.ctrl->feat_free_top, &m.ctrl->feat_high, m.feat_free, m.feat_capacity, &m.ctrl->overflow);
    if (fs >= 0) {
      m.blk_feat[slot] = fs;
      m.blk_layers[slot] |= kLayerFeatBit;
      atomicAdd(&m.ctrl->n_feat, 1);
      *newfeat_slot_out = fs;
    }
  }
}
__global__ void k_zero_one_feature_block(MapDev m, const int* fslot) {
  if (*fslot < 0) return;
  uint4* p = reinterpret_cast<uint4*>(feat_block(m, *fslot));
  const int vec_per_block = (kVoxelsPerBlock * m.row) / 8;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < vec_per_block; k += gridDim.x * blockDim.x)
    p[k] = make_uint4(0, 0, 0, 0);
}
__global__ void k_find_one(MapDev m, int x, int y, int z, int layer, unsigned long long* ptr_out) {
  const int slot = hash_find(m, x, y, z);
  *ptr_out = 0ull;
  if (slot < 0) return;
  if (layer == 0) {
    if (m.blk_layers[slot] & kLayerTsdfBit) *ptr_out = (unsigned long long)tsdf_block(m, slot);
  } else {
    if (m.blk_feat[slot] >= 0) *ptr_out = (unsigned long long)feat_block(m, m.blk_feat[slot]);
  }
}

// queryTSDFKernel (NT/cpp/src/sdf_query.cu:240-270): one thread per query, rows of misses untouched.
__global__ void __launch_bounds__(128) k_query_tsdf(MapDev m, const float* __restrict__ xyz, long long n,
                                                    float2* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 p;
  p.x = xyz[3 * i];
  p.y = xyz[3 * i + 1];
  p.z = xyz[3 * i + 2];
  I3 b, v;
  block_and_voxel_from_position(m.block_size, m.voxel_size_inv, p, &b, &v);
  const int slot = hash_find(m, b.x, b.y, b.z);
  if (slot < 0 || !(m.blk_layers[slot] & kLayerTsdfBit)) return;
  out[i] = tsdf_block(m, slot)[(v.x * 8 + v.y) * 8 + v.z];
}
// queryFeatureKernel (sdf_query.cu:206-238): one WARP per query; the output row has C+1 halves
// (2-byte aligned), so lanes write single halves, coalesced.
__global__ void __launch_bounds__(128) k_query_features(MapDev m, const float* __restrict__ xyz, long long n,
                                                        __half* __restrict__ out) {
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n) return;
  V3 p;
  p.x = xyz[3 * q];
  p.y = xyz[3 * q + 1];
  p.z = xyz[3 * q + 2];
  I3 b, v;
  block_and_voxel_from_position(m.block_size, m.voxel_size_inv, p, &b, &v);
  const int slot = hash_find(m, b.x, b.y, b.z);
  if (slot < 0 || m.blk_feat[slot] < 0) return;
  const __half* row = feat_block(m, m.blk_feat[slot]) + (size_t)((v.x * 8 + v.y) * 8 + v.z) * m.row;
  __half* o = out + q * (long long)(m.C + 1);
  for (int c = lane_id(); c <= m.C; c += 32) o[c] = row[c];
}

}  // namespace nvbx
