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

// Programmatic dependent launch (sm_90+): every kernel of the library is launched with the
// programmatic-stream-serialization attribute and starts with pdl_prologue().  launch_dependents lets the
// NEXT kernel of the stream be scheduled while this one still runs (its CTAs fill free SM slots and park
// in griddepcontrol.wait); wait blocks until the PREVIOUS kernel has completed and its writes are
// visible.  Net effect: the ~2-3 us launch latency of each of the 6 kernels of a frame is hidden behind
// its predecessor.  Nothing produced by an earlier kernel may be read before pdl_prologue().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}

__device__ __forceinline__ void count_add(const MapDev& m, int id, unsigned long long v) {
  atomicAdd(&m.ctrl->counters[id], v);
}

// ---- pipeline timeline (NVBX_PROFILE_COUNTERS builds; tools/pipeline_timeline.py) ------------------
// Every kernel of a frame stamps %globaltimer when its first CTA passes griddepcontrol.wait and when its last
// CTA ends, into g_prof[frame & 63][kernel]: what the PDL-chained pipeline really looks like in steady state,
// which neither isolated event timings nor a serialising profiler can show.  The frame number is the host's count
// of feature frames enqueued so far (MapDev::seq / RaycastFrame::seq), so kernels of different frames that run
// side by side on the pipelining streams stamp their own rows.
enum { kProfRaycast = 0, kProfTsdf, kProfTrace, kProfGeometry, kProfGather, kProfRaycastPre, kProfKernels = 8 };
#ifdef NVBX_PROFILE_COUNTERS
__device__ unsigned long long g_prof[64 * 2 * kProfKernels];
// per-CTA stamps of the raycast kernel (entry, march done, wait passed, exit) + SM id, frames mod 4 (NVBX_PROF_CTAS dump)
__device__ unsigned long long g_prof_cta[4][256][5];
__device__ __forceinline__ unsigned long long prof_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void prof_begin(int seq, int k, unsigned long long t) {
  atomicMin(&g_prof[((seq & 63) * kProfKernels + k) * 2], t);
}
__device__ __forceinline__ void prof_end(int seq, int k, unsigned long long t) {
  atomicMax(&g_prof[((seq & 63) * kProfKernels + k) * 2 + 1], t);
}
#define PROF_BEGIN(seq, k)                       \
  const int prof_seq_ = (seq);                   \
  if (threadIdx.x == 0) prof_begin(prof_seq_, (k), prof_now())
#define PROF_END(k)                              \
  do {                                           \
    __syncthreads();                             \
    if (threadIdx.x == 0) prof_end(prof_seq_, (k), prof_now()); \
  } while (0)
#else
#define PROF_BEGIN(seq, k)
#define PROF_END(k)
#endif

// ================================================================================================
// a2. Blocks in view by ray casting -- one thread per (subsampled) depth pixel, 3-D DDA over the view
// AABB.  Follows combinedBlockIndicesInImageKernel (NB/src/integrators/view_calculator.cu:196-248) and
// RayCaster (ray_caster_impl.h:26-75), including the linear-index alias of setIndexUpdated (:171-180).
//
// The reference marks a dense bool grid in global memory: 263 k rays hammer a few hundred bytes, and
// the same-address traffic serialises in L2.  Here the grid is a BITMAP; every CTA marks a private
// copy in shared memory (read-before-atomicOr, so a set bit costs one broadcast load) and ORs its
// non-zero words into the global bitmap once, at the end of its persistent tile loop.  Views whose
// AABB has more cells than the shared bitmap holds mark the global bitmap directly.
// ================================================================================================
struct ViewGrid {
  I3 mn;
  int sx, sy, sz;
  int n_cells;
};

constexpr int kRayBitmapWords = 8192;  // 32 KiB of shared memory = 262 144 cells

// Mark linear cell `lin` (already aliased exactly like setIndexUpdated: any int whose unsigned value is
// below n_cells hits a cell, whatever the per-axis indices were).
template <bool SMEM>
__device__ __forceinline__ void mark_lin(unsigned* bm, int n_cells, int lin) {
  if ((unsigned)lin < (unsigned)n_cells) {
    const unsigned w = (unsigned)lin >> 5, bit = 1u << ((unsigned)lin & 31u);
    if (SMEM) {
      if (!(bm[w] & bit)) atomicOr(&bm[w], bit);
    } else {
      if (!(__ldcg(&bm[w]) & bit)) atomicOr(&bm[w], bit);
    }
  }
}

// Everything about a depth frame's rays that does not depend on the pixel, computed once on the host
// with the same IEEE operations the per-ray code of the reference performs (RayCaster constructor,
// ray_caster_impl.h:26-53: start = origin / block_size, floor, shifted start).
struct RaycastFrame {
  Pose T_L_C;
  Cam cam;
  const float* depth;
  int rows, cols;
  float block_size, max_dist, behind;
  int sub;
  ViewGrid g;
  float s[3];        // ray start in block units: (T_L_C.t / block_size) / 1
  int start[3];      // floor(s)
  float shifted[3];  // s - start
  int lin0;          // linear (aliased) grid index of the start block
  int tiles_x, n_tiles;
  int seq;           // profile builds: frame number for the timeline stamps (feature frames enqueued so far)
};

template <bool SMEM>
__global__ void __launch_bounds__(256) k_raycast_mark(RaycastFrame f, unsigned* gbits, int* entry_count,
                                                      int flush_early, unsigned flush_seq, const unsigned* clean_seq) {
  // SMEM: the rays are marched BEFORE griddepcontrol.wait, i.e. while the previous kernels of the stream (the
  // last frame's feature gather) drain.  Until the wait this kernel reads only its arguments and the caller's
  // depth image and writes only shared memory; the depth image was produced by an operation that is not one of
  // this library's programmatic launches (a copy, a torch kernel), and such an operation is fully ordered with
  // respect to the launches on either side of it.  Every global write happens after the wait.
  pdl_trigger();
#ifdef NVBX_PROFILE_COUNTERS
  const unsigned long long prof_entry = prof_now();
#endif
  if (!SMEM) pdl_wait();
  extern __shared__ unsigned s_bits[];
  const ViewGrid& g = f.g;
  const int n_words = (g.n_cells + 31) >> 5;
  unsigned* bm = SMEM ? s_bits : gbits;
  if (SMEM) {
    for (int w = threadIdx.x; w < n_words; w += 256) s_bits[w] = 0u;
    __syncthreads();
  }
  if (!SMEM && blockIdx.x == 0 && threadIdx.x == 0) *entry_count = 0;  // consumed by k_view_compact_alloc (next launch)
  const int stride_y = g.sx, stride_z = g.sx * g.sy;

  for (int tile = blockIdx.x; tile < f.n_tiles; tile += gridDim.x) {
    const int ray_col = (tile % f.tiles_x) * 16 + (threadIdx.x & 15);
    const int ray_row = (tile / f.tiles_x) * 16 + (threadIdx.x >> 4);
    int prow = ray_row * f.sub, pcol = ray_col * f.sub;
    if (prow >= f.rows + f.sub - 1 || pcol >= f.cols + f.sub - 1) continue;
    if (prow >= f.rows) prow = f.rows - 1;
    if (pcol >= f.cols) pcol = f.cols - 1;
    float d = __ldg(f.depth + (size_t)prow * f.cols + pcol);
    if (d <= 0.0f) continue;
    if (f.max_dist > 0.0f && d > f.max_dist) d = f.max_dist;
    const V3 ray = ray_from_image_plane(f.cam, (float)pcol + 0.5f, (float)prow + 0.5f);
    const float len = d + f.behind;
    V3 p_C;
    p_C.x = len * ray.x;
    p_C.y = len * ray.y;
    p_C.z = len * ray.z;
    const V3 p_L = dev_xform(f.T_L_C, p_C);
    // ray end in block units; its floor is also the block index of the end point (view_calculator.cu:231-236)
    const float e[3] = {(p_L.x / f.block_size) / 1.0f, (p_L.y / f.block_size) / 1.0f, (p_L.z / f.block_size) / 1.0f};
    const int end[3] = {(int)floorf(e[0]), (int)floorf(e[1]), (int)floorf(e[2])};
    mark_lin<SMEM>(bm, g.n_cells, (end[0] - g.mn.x) + (end[1] - g.mn.y) * stride_y + (end[2] - g.mn.z) * stride_z);

    // RayCaster(origin / bs, p_L / bs), scale 1 -- the per-pixel half of the constructor
    int dlin[3];
    float t_next[3], t_step[3];
    int steps = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      steps += abs(end[i] - f.start[i]);
      const float r = e[i] - f.s[i];
      const int sign = (r > 0.0f) ? 1 : ((r < 0.0f) ? -1 : 0);
      const float dist = (float)max(sign, 0) - f.shifted[i];
      t_next[i] = dist / r;
      t_step[i] = (float)sign / r;
      dlin[i] = sign * (i == 0 ? 1 : (i == 1 ? stride_y : stride_z));
    }
    int lin = f.lin0;
    for (int step = 0; step <= steps; ++step) {
      mark_lin<SMEM>(bm, g.n_cells, lin);
      // Eigen minCoeff: first minimum wins; NaNs (0/0 on a degenerate axis) never compare smaller
      const bool y_lt_x = t_next[1] < t_next[0];
      const float mxy = y_lt_x ? t_next[1] : t_next[0];
      const bool z_min = t_next[2] < mxy;
      const bool y_min = y_lt_x && !z_min;
      const bool x_min = !y_lt_x && !z_min;
      lin += z_min ? dlin[2] : (y_min ? dlin[1] : dlin[0]);
      t_next[0] += x_min ? t_step[0] : 0.0f;
      t_next[1] += y_min ? t_step[1] : 0.0f;
      t_next[2] += z_min ? t_step[2] : 0.0f;
    }
  }
#ifdef NVBX_PROFILE_COUNTERS
  const unsigned long long prof_marched = prof_now();
#endif
#ifdef NVBX_PROFILE_COUNTERS
  if (threadIdx.x == 0 && blockIdx.x < 256) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* r = g_prof_cta[f.seq & 3][blockIdx.x];
    r[0] = prof_entry;
    r[1] = prof_marched;
    r[4] = smid;
  }
#endif
  if (SMEM) {
    // The marks go to one third of a triple-buffered bitmap (flush_early): this third was last read two depth
    // frames ago and is wiped by the TSDF kernel of the PREVIOUS depth frame, which publishes Ctrl::bitmap_clean_seq
    // when it is done.  If that has happened (practically always: the wipe is the first thing that kernel does) the
    // marks are flushed right here, before the wait, under the tail of the preceding kernels; if not, after it.
    __shared__ int s_early;
    if (threadIdx.x == 0) {
      int ok = 0;
      if (flush_early) {  // wrap-safe: the counter runs for the life of the map
        ok = (int)(*reinterpret_cast<const volatile unsigned*>(clean_seq) - flush_seq) >= 0;
        __threadfence();
      }
      s_early = ok;
    }
    __syncthreads();
    const bool early = s_early != 0;
    if (early) {
      for (int w = threadIdx.x; w < n_words; w += 256) {
        const unsigned v = s_bits[w];
        if (v && (__ldcg(&gbits[w]) & v) != v) atomicOr(&gbits[w], v);
      }
      // This CTA is done.  Only CTA 0 stays to sit out the wait: the grid then still completes after its predecessor
      // (the next kernel's own wait relies on that order), while every other CTA hands its registers and thread
      // slots back at once instead of parking on them, and CTAs that did not fit on the machine at launch get
      // their turn long before the wait is over.
#ifdef NVBX_PROFILE_COUNTERS
      if (threadIdx.x == 0) {  // pre-wait row of the timeline: first entry .. last flush over ALL CTAs
        prof_begin(f.seq, kProfRaycastPre, prof_entry);
        prof_end(f.seq, kProfRaycastPre, prof_now());
      }
#endif
      if (blockIdx.x != 0) return;
    }
    pdl_wait();
#ifdef NVBX_PROFILE_COUNTERS
    if (threadIdx.x == 0 && blockIdx.x < 256) g_prof_cta[f.seq & 3][blockIdx.x][2] = prof_now();
#endif
    PROF_BEGIN(f.seq, kProfRaycast);
#ifdef NVBX_PROFILE_COUNTERS
    if (threadIdx.x == 0) {
      prof_begin(prof_seq_, kProfRaycastPre, prof_entry);
      prof_end(prof_seq_, kProfRaycastPre, prof_marched);
    }
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0) *entry_count = 0;
    if (!early) {
      for (int w = threadIdx.x; w < n_words; w += 256) {
        const unsigned v = s_bits[w];
        if (v && (__ldcg(&gbits[w]) & v) != v) atomicOr(&gbits[w], v);
      }
    }
    PROF_END(kProfRaycast);
  }
}

// Number of marked cells (only launched while the block arena is still growing, see ensure_slots()).
__global__ void __launch_bounds__(256) k_count_marked(const unsigned* __restrict__ gbits, int n_words, int* out) {
  pdl_prologue();
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) c += __popc(gbits[i]);
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane_id() == 0 && c) atomicAdd(out, c);
}

// Give `slot` a TSDF layer if it has none; returns true if the payload is uninitialised (the TSDF update
// that follows in stream order writes all 512 voxels of such a block, see k_tsdf_update).
__device__ __forceinline__ bool ensure_tsdf_layer(const MapDev& m, int slot) {
  const uint8_t layers = m.blk_layers[slot];
  if (layers & kLayerTsdfBit) return false;
  m.blk_layers[slot] = layers | kLayerTsdfBit;
  atomicAdd(&m.ctrl->n_tsdf, 1);
  return true;
}

// Scan the marked bitmap -> block indices (appended to the viewpoint-cache entry), find-or-allocate each
// block in the map, emit the slot list for the TSDF update, and clear the bitmap for the next frame.
// Replaces the D2H copy + CPU scan + CPU allocation + H2D pointer tables of view_calculator.cu:313-323 /
// layer_impl.h:130-162 / integrators_common_impl.h:60-121.  One 32-cell word per warp, one cell per lane.
__global__ void __launch_bounds__(256) k_view_compact_alloc(MapDev m, unsigned* gbits, ViewGrid g, int3* entry_idx,
                                                            int* entry_count, int* view_slots) {
  pdl_prologue();
  const int n_words = (g.n_cells + 31) >> 5;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int lane = lane_id();
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps_total) {
    const unsigned bits = gbits[w];
    if (!bits) continue;
    __syncwarp();
    if (lane == 0) gbits[w] = 0u;  // self-cleaning: the bitmap is all-zero again when this kernel ends
    int pos0 = 0;
    if (lane == 0) pos0 = atomicAdd(entry_count, __popc(bits));
    pos0 = __shfl_sync(0xffffffffu, pos0, 0);
    bool fresh = false;
    if ((bits >> lane) & 1u) {
      const int i = (w << 5) + lane;
      const int x = i % g.sx + g.mn.x;
      const int y = (i / g.sx) % g.sy + g.mn.y;
      const int z = i / (g.sx * g.sy) + g.mn.z;
      const int pos = pos0 + __popc(bits & ((1u << lane) - 1u));
      entry_idx[pos] = make_int3(x, y, z);
      bool is_new;
      int slot = acquire_slot(m, x, y, z, &is_new);
      if (slot >= 0 && ensure_tsdf_layer(m, slot)) {
        slot |= kNewFlag;
        fresh = true;
      }
      view_slots[pos] = slot;
    }
    const unsigned fb = __ballot_sync(0xffffffffu, fresh);
    if (fb && lane == 0) count_add(m, kCntTsdfBlocksAllocated, __popc(fb));
  }
}

// ================================================================================================
// a4. TSDF update.  integrateBlocksKernel<TsdfVoxel> (projective_integrator_impl.cuh:58-103) +
// UpdateTsdfVoxelFunctor (projective_tsdf_integrator.cu:25-99).  One CTA of 512 threads per block,
// thread t owns voxel t of the [x][y][z] array (z fastest -> 32 lanes read 256 contiguous bytes).
// A block allocated by this frame (kNewFlag) is never read: its voxels start from (0, 0) and all 512
// are written, which replaces the reference's per-block cudaMemset (blox_impl.h:92-97).
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
  const V3 pl = dev_voxel_center(block_size, bi, vx, vy, vz);
  const V3 pc = dev_xform(T_C_L, pl);
  if (!dev_project(cam, pc, u, v)) return false;
  *vd = pc.z;
  if (max_depth > 0.0f && *vd > max_depth) return false;
  return true;
}

// One voxel of one block (thread t of the CTA), in two steps so that the depth sample -- which only needs the
// block's INDEX -- can be in flight while the block's slot is still being looked up or allocated:
//   tsdf_measure : project the voxel centre, read the depth pixel (and mask) it falls on;
//   tsdf_fuse    : blend with the stored voxel; returns 1 if the voxel was fused with a measurement.  *is_free: the
//                  voxel was written by this frame and now holds (+truncation distance, weight > 1e-4).
struct VoxMeas {
  float meas, vd;
  bool have;    // a depth sample (not NaN) was read
  bool active;  // mask
};
__device__ __forceinline__ VoxMeas tsdf_measure(const DepthFrame& f, float block_size, const int3 b, int t) {
  const int vx = t >> 6, vy = (t >> 3) & 7, vz = t & 7;
  VoxMeas r;
  r.meas = r.vd = 0.0f;
  r.have = false;
  r.active = true;
  float u, v;
  if (!project_voxel(f.cam, f.T_C_L, block_size, f.max_depth, b, vx, vy, vz, &u, &v, &r.vd)) return r;
  const int ui = (int)floorf(u), vi = (int)floorf(v);
  if (ui < 0 || vi < 0 || ui >= f.cols || vi >= f.rows) return r;
  r.meas = __ldg(f.depth + (size_t)vi * f.cols + ui);
  if (isnan(r.meas)) return r;
  r.active = (f.mask == nullptr) || __ldg(f.mask + (size_t)vi * f.cols + ui);
  r.have = true;
  return r;
}
__device__ __forceinline__ unsigned tsdf_fuse(const MapDev& m, const DepthFrame& f, int slot, bool is_new, int t,
                                              const VoxMeas& vm, bool* is_free) {
  float2* vox = tsdf_block(m, slot) + t;
  bool write = is_new;
  float2 out = make_float2(0.0f, 0.0f);
  unsigned updated = 0;
  // the stored voxel travels while the measurement is being judged (an existing block's payload is always readable)
  float2 cur = make_float2(0.0f, 0.0f);
  if (vm.have && !is_new) cur = *vox;
  do {
    if (!vm.have) break;
    const float meas = vm.meas, vd = vm.vd;
    if (meas <= 0.0f) {
      if (f.invalid_decay >= 0.0f && !is_new) {
        out = cur;
        out.y *= f.invalid_decay;
        write = true;
      }
      break;
    }
    const float sdf = meas - vd;
    if (sdf < -f.trunc) break;
    if (!vm.active && sdf < f.trunc) break;
    const float w_m = weighting(f.weighting_mode, meas, vd, f.trunc);
    float fused = fmaf(cur.x, cur.y, sdf * w_m) / (w_m + cur.y);   // FMUL + FFMA in the reference
    if (fused > 0.0f)
      fused = fminf(f.trunc, fused);
    else
      fused = fmaxf(-f.trunc, fused);
    out = make_float2(fused, fminf(w_m + cur.y, f.max_weight));
    write = true;
    updated = 1;
  } while (false);
  if (write) *vox = out;
  *is_free = write && out.x == f.trunc && out.y > 1e-4f;
  return updated;
}
__device__ __forceinline__ unsigned tsdf_update_voxel(const MapDev& m, const DepthFrame& f, int slot, bool is_new,
                                                      int t, bool* is_free) {
  const VoxMeas vm = tsdf_measure(f, m.block_size, m.blk_index[slot], t);
  return tsdf_fuse(m, f, slot, is_new, t, vm, is_free);
}
// All 512 voxels free -> kBlockFreeBit of the slot (one barrier; every thread of the CTA calls this).
__device__ __forceinline__ void publish_block_free(const MapDev& m, int slot, bool voxel_free) {
  const int all = __syncthreads_and(voxel_free);
  if (threadIdx.x == 0) {
    const uint8_t l = m.blk_layers[slot];
    m.blk_layers[slot] = all ? (l | kBlockFreeBit) : (l & (uint8_t)~kBlockFreeBit);
  }
}

// Where the list of blocks to update comes from.
//   kViewFromSlots  : slot list written by k_view_compact_alloc (views whose bitmap exceeds kFusedBitmapWords)
//   kViewFromBitmap : the marked bitmap itself -- compaction, find-or-allocate and the viewpoint-cache entry
//                     are folded into this kernel (every CTA ranks the <= 32 768 cells redundantly in shared
//                     memory and serves ranks blockIdx.x, blockIdx.x + gridDim.x, ...); the bitmap the depth
//                     frame after next will mark (triple buffer) is wiped.  One launch less per depth frame.
//   kViewFromEntry  : viewpoint-cache hit, the cached index list (view_calculator.cu:256-265); blocks are
//                     (re-)allocated where required (projective_integrator_impl.cuh:288-291)
enum ViewMode { kViewFromSlots = 0, kViewFromBitmap = 1, kViewFromEntry = 2 };
constexpr int kFusedBitmapWords = 1024;

struct ViewSource {
  const int* view_slots;  // kViewFromSlots
  int* entry_count;       // list length (read in kViewFromSlots / kViewFromEntry, written in kViewFromBitmap)
  int3* entry_idx;        // cache entry index list (written in kViewFromBitmap, read in kViewFromEntry)
  const unsigned* bits;   // kViewFromBitmap
  unsigned* clean_bits;   // kViewFromBitmap: the third of the triple buffer that the depth frame after next marks
  unsigned clean_seq;     // kViewFromBitmap: value published in Ctrl::bitmap_clean_seq once clean_bits is wiped
  ViewGrid g;
};

template <int MODE>
__global__ void __launch_bounds__(512, 4) k_tsdf_update(MapDev m, ViewSource src, DepthFrame f) {
  pdl_prologue();
  PROF_BEGIN(m.seq, kProfTsdf);
  __shared__ int s_off[513];
  __shared__ int s_warp[17];
  __shared__ int s_slot;
  __shared__ int3 s_b;
  const int t = threadIdx.x;
  unsigned updated = 0;
  int n = 0;

  if (MODE == kViewFromSlots) {
    n = *src.entry_count;
    for (int bi = blockIdx.x; bi < n; bi += gridDim.x) {
      const int raw = src.view_slots[bi];
      if (raw < 0) continue;
      const int slot = raw & kSlotMask;
      if (t == 0) m.blk_dirty[slot] = kDirtyAll;  // blocks_to_update_tracker_.addBlocksToUpdate (mapper.cpp:406)
      bool vfree;
      updated += tsdf_update_voxel(m, f, slot, (raw & kNewFlag) != 0, t, &vfree);
      publish_block_free(m, slot, vfree);
    }
  } else {
    unsigned w0 = 0u, w1 = 0u;
    if (MODE == kViewFromBitmap) {
      // rank the marked cells: thread t owns words 2t and 2t + 1
      const int n_words = (src.g.n_cells + 31) >> 5;
      if (2 * t < n_words) w0 = src.bits[2 * t];
      if (2 * t + 1 < n_words) w1 = src.bits[2 * t + 1];
      if (blockIdx.x == 0) {  // wipe the bitmap of the depth frame after next, then say so (k_raycast_mark)
        for (int w = t; w < kFusedBitmapWords; w += 512) src.clean_bits[w] = 0u;
        __threadfence();
        __syncthreads();
        if (t == 0) *reinterpret_cast<volatile unsigned*>(&m.ctrl->bitmap_clean_seq) = src.clean_seq;
      }
      const int c = __popc(w0) + __popc(w1);
      int inc = c;
      const int lane = t & 31, warp = t >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) s_warp[warp + 1] = inc;
      __syncthreads();
      if (t == 0) {
        int acc = 0;
        s_warp[0] = 0;
        for (int w = 1; w <= 16; ++w) {
          acc += s_warp[w];
          s_warp[w] = acc;
        }
      }
      __syncthreads();
      s_off[t] = s_warp[warp] + inc - c;
      if (t == 511) s_off[512] = s_warp[16];
      __syncthreads();
      n = s_off[512];
      if (blockIdx.x == 0 && t == 0) *src.entry_count = n;
    } else {
      n = *src.entry_count;
    }
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
      bool owner;
      int x = 0, y = 0, z = 0;
      if (MODE == kViewFromBitmap) {
        owner = (s_off[t] <= r) && (r < s_off[t + 1]);
        if (owner) {
          const int k = r - s_off[t], c0 = __popc(w0);
          const int i = (k < c0) ? (64 * t + (int)__fns(w0, 0, k + 1)) : (64 * t + 32 + (int)__fns(w1, 0, k - c0 + 1));
          x = i % src.g.sx + src.g.mn.x;
          y = (i / src.g.sx) % src.g.sy + src.g.mn.y;
          z = i / (src.g.sx * src.g.sy) + src.g.mn.z;
          src.entry_idx[r] = make_int3(x, y, z);
        }
      } else {
        owner = (t == 0);
        if (owner) {
          const int3 b = src.entry_idx[r];
          x = b.x;
          y = b.y;
          z = b.z;
        }
      }
      if (owner) s_b = make_int3(x, y, z);
      __syncthreads();
      // every thread: the depth sample of its voxel (needs the block's index only) flies while the owner thread
      // finds or allocates the block's slot
      const VoxMeas vm = tsdf_measure(f, m.block_size, s_b, t);
      if (owner) {
        bool is_new;
        int slot = acquire_slot(m, x, y, z, &is_new);
        if (slot >= 0) {
          if (ensure_tsdf_layer(m, slot)) {
            slot |= kNewFlag;
            count_add(m, kCntTsdfBlocksAllocated, 1);
          }
          m.blk_dirty[slot & kSlotMask] = kDirtyAll;  // blocks_to_update_tracker_.addBlocksToUpdate (mapper.cpp:406)
        }
        s_slot = slot;
      }
      __syncthreads();
      const int raw = s_slot;
      if (raw >= 0) {  // uniform per CTA
        bool vfree;
        updated += tsdf_fuse(m, f, raw & kSlotMask, (raw & kNewFlag) != 0, t, vm, &vfree);
        publish_block_free(m, raw & kSlotMask, vfree);
      }
      __syncthreads();  // s_slot / s_b are rewritten by the next rank
    }
  }
  // accounting: one atomic per warp
  for (int o = 16; o; o >>= 1) updated += __shfl_xor_sync(0xffffffffu, updated, o);
  if (lane_id() == 0 && updated) count_add(m, kCntTsdfVoxelsUpdated, updated);
  if (blockIdx.x == 0 && t == 0) {
    count_add(m, kCntTsdfBlocksInView, (unsigned long long)n);
    count_add(m, kCntDepthFrames, 1);
  }
  PROF_END(kProfTsdf);
}

// ================================================================================================
// a6. Feature candidates.  getBlocksInViewPlanes (view_calculator.cu:392-470) +
// reduceBlocksToThoseInTruncationBand (projective_appearance_integrator.cu:374-477) + feature block
// allocation (:120-123) in ONE kernel.  Lanes test 32 cells of the frustum AABB in parallel (block
// centre against the normalised viewport widened by 10 px, then one index lookup); the warp then scans
// each surviving TSDF block (4 KiB, all eight 128-bit loads of a lane in flight at once) for a voxel
// with w > 0 and |d| < trunc, and pops a feature slot for band blocks that have none.
// `view` carries the pose / camera / AABB the viewpoint cache holds for this call (Q6): on a cache hit
// those are the CACHED ones, which reproduces the reference re-using its cached index list.
// ================================================================================================
struct PlanesView {
  ViewGrid g;
  Pose T_C_L;
  float vmin_x, vmin_y, vmax_x, vmax_y;
};
constexpr int kBandCountOnly = -2;  // band_select_tile mode: count the band blocks that have no feature block yet

// One tile = `tile_cells` consecutive cells of the AABB (a multiple of 32, <= 256): phase 1 tests one cell
// per thread and compacts the surviving slots into shared memory, phase 2 deals them to the CTA's warps.
// Small tiles spread the few hundred candidates of a mindmap workspace over many SMs.
// color_parity < 0: feature frame (band blocks get a feature slot).  color_parity = 0 / 1: colour frame -- band
// blocks get the colour layer bit (their ColorVoxel payload lives at the slot id, like the TSDF payload) and are
// appended to the colour band list counted by ctrl->cband_count[color_parity].
template <int ILP>
__device__ __forceinline__ void band_select_tile(const MapDev& m, const PlanesView& view, float trunc,
                                                 int* band_slots, int* newfeat_slots, int tile, int tile_cells,
                                                 int* s_cand, int* s_ncand, int* s_band, int* s_nband,
                                                 int color_parity) {
  const ViewGrid& g = view.g;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    *s_ncand = 0;
    *s_nband = 0;
  }
  __syncthreads();
  const int i = tile * tile_cells + threadIdx.x;
  int slot = -1;
  if (threadIdx.x < tile_cells && i < g.n_cells) {
    const int x = i % g.sx + g.mn.x;
    const int y = (i / g.sx) % g.sy + g.mn.y;
    const int z = i / (g.sx * g.sy) + g.mn.z;
    V3 c;
    c.x = m.block_size * ((float)x + 0.5f);
    c.y = m.block_size * ((float)y + 0.5f);
    c.z = m.block_size * ((float)z + 0.5f);
    const V3 r = rotate(view.T_C_L, c);
    const float px = r.x + view.T_C_L.t[0], py = r.y + view.T_C_L.t[1], pz = r.z + view.T_C_L.t[2];
    if (pz > 1e-6f) {
      const float un = px / pz, vn = py / pz;
      if ((view.vmin_x <= un) && (view.vmin_y <= vn) && (un <= view.vmax_x) && (vn <= view.vmax_y)) {
        slot = find_slot(m, x, y, z);
        if (slot >= 0 && !(m.blk_layers[slot] & kLayerTsdfBit)) slot = -1;
      }
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, slot >= 0);
  if (ballot) {
    int base = 0;
    if (lane == 0) base = atomicAdd(s_ncand, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (slot >= 0) s_cand[base + __popc(ballot & ((1u << lane) - 1u))] = slot;
  }
  __syncthreads();
  const int n_cand = *s_ncand;
  // phase 2a: one warp per candidate, two candidates (sixteen 128-bit loads per lane) in flight; band blocks are
  // collected in shared memory so that nothing below waits on a global atomic per block
  const int nwarps = blockDim.x >> 5;
  for (int k = warp; k < n_cand; k += ILP * nwarps) {
    const int s0 = s_cand[k];
    const bool two = ILP == 2 && k + nwarps < n_cand;
    const int s1 = two ? s_cand[k + nwarps] : s0;
    const float4* p0 = reinterpret_cast<const float4*>(tsdf_block(m, s0));
    const float4* p1 = reinterpret_cast<const float4*>(tsdf_block(m, s1));
    float4 q0[8], q1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) q0[j] = p0[j * 32 + lane];  // two voxels each: (d, w, d, w)
    if (two) {
#pragma unroll
      for (int j = 0; j < 8; ++j) q1[j] = p1[j * 32 + lane];
    }
    bool hit0 = false, hit1 = false;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      hit0 |= (q0[j].y > 0.0f && fabsf(q0[j].x) < trunc) || (q0[j].w > 0.0f && fabsf(q0[j].z) < trunc);
    if (two) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        hit1 |= (q1[j].y > 0.0f && fabsf(q1[j].x) < trunc) || (q1[j].w > 0.0f && fabsf(q1[j].z) < trunc);
    }
    const bool any0 = __any_sync(0xffffffffu, hit0), any1 = __any_sync(0xffffffffu, hit1);
    if (lane == 0) {
      if (any0) s_band[atomicAdd(s_nband, 1)] = s0;
      if (any1) s_band[atomicAdd(s_nband, 1)] = s1;
    }
  }
  __syncthreads();
  // phase 2b: one THREAD per band block -- layer bookkeeping and feature-slot allocation of all blocks of the
  // tile proceed in parallel, and each warp appends its blocks to the global list with one atomic
  const int n_band = *s_nband;
  for (int j0 = 0; j0 < n_band; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    bool has = false;
    int out = 0;
    if (j < n_band) {
      const int s = s_band[j];
      if (color_parity == kBandCountOnly) {  // k_band_count: how many feature blocks would this frame allocate?
        if (m.blk_feat[s] < 0) atomicAdd(&m.ctrl->list_count, 1);
      } else if (color_parity >= 0) {
        int flag = 0;
        const uint8_t layers = m.blk_layers[s];
        if (!(layers & kLayerColorBit)) {
          m.blk_layers[s] = layers | kLayerColorBit;
          atomicAdd(&m.ctrl->n_color, 1);
          count_add(m, kCntColorBlocksAllocated, 1);
          flag = kNewFlag;  // k_color_update writes all 512 voxels of such a block (gray, weight 0 where not fused)
        }
        has = true;
        out = s | flag;
      } else {
        int flag = 0;
        int fs = m.blk_feat[s];
        if (fs < 0) {
          fs = pop_id(&m.ctrl->feat_free_top, &m.ctrl->feat_high, m.feat_free, m.feat_capacity, &m.ctrl->overflow);
          if (fs >= 0) {
            m.blk_feat[s] = fs;
            m.blk_layers[s] |= kLayerFeatBit;
            atomicAdd(&m.ctrl->n_feat, 1);
            count_add(m, kCntFeatBlocksAllocated, 1);
            newfeat_slots[atomicAdd(&m.ctrl->newfeat_count[m.fp], 1)] = fs;
            flag = kNewFlag;  // zero-filled (cooperatively, by all CTAs) in k_feature_geometry
          }
        }
        has = fs >= 0;
        out = s | flag;
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, has);
    if (ballot) {
      int* counter = color_parity >= 0 ? &m.ctrl->cband_count[color_parity] : &m.ctrl->band_count[m.fp];
      int base = 0;
      if (lane == 0) base = atomicAdd(counter, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (has) band_slots[base + __popc(ballot & ((1u << lane) - 1u))] = out;
    }
  }
  if (threadIdx.x == 0 && n_cand && color_parity == -1) count_add(m, kCntFeatCandidateBlocks, (unsigned)n_cand);
  __syncthreads();  // s_cand / s_ncand are reused by the CTA's next tile
}

// ================================================================================================
// a7. Synthetic depth by sphere tracing.  sphereTracingKernel + cast (sphere_tracer.cu:31-131,191-236).
// One thread per ray.  Block lookups go through the direct-mapped workspace grid (one L1-resident load);
// the last block is cached in registers so consecutive samples inside one block cost no lookup; a ray
// that has left the workspace grid moving away from it can never see a valid sample again, so it is
// terminated at once (the reference keeps stepping by the truncation distance until 7 m / 100 steps and
// then reports the same miss).
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
  float free_dist;  // the TSDF truncation distance: what every voxel of a kBlockFreeBit block holds
  int use_free;     // tuning knob (NVBX_TRACE_FREE): consult kBlockFreeBit
  int team;         // 1: eight lanes per ray (sphere_trace_team), trace CTAs cover 8 x 4 rays; 0: one thread per ray
  int cache_blocks; // 0: scalar march; 2 / 4: sphere_trace_ray_cached with that many TSDF blocks per warp in (dynamic)
                    // shared memory (8 warps x cache_blocks x 4 KB + tags)
  int march;        // marching threads per trace CTA: 256 (16 x 16 rays) or 128 / 64 / 32 (8 columns x march / 8 rows;
                    // the CTA's other warps only help to stage the workspace table): spreads the rays over more SMs
};

// floor(p / block_size) exactly as block_index_from_position computes it, without the IEEE division on
// the common path: q = p * (1/bs) is within |q| * 2^-22 of the true quotient, so when q sits farther than
// 1e-4 from an integer (and |q| < 256) its floor is the floor of the correctly rounded quotient too.
__device__ __forceinline__ int floor_div_exact(float p, float bs, float bs_inv) {
  const float q = p * bs_inv;
  const float fl = floorf(q);
  const float fr = q - fl;
  if (fr > 1e-4f && fr < 1.0f - 1e-4f && fabsf(q) < 256.0f) return (int)fl;
  return (int)floorf(p / bs);
}

constexpr int kTraceSmemCells = 4096;  // workspace grids up to this many cells are staged in shared memory

template <int kSpec>
__device__ __forceinline__ int sphere_trace_ray(const MapDev& m, const TraceParams& tp, int r, int c,
                                                float* __restrict__ image, const int* s_ws) {
  const float pu = (float)(c * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const float pv = (float)(r * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const V3 ray = ray_from_image_plane(tp.cam, pu, pv);
  const float sq = fmaf(ray.x, ray.x, fmaf(ray.y, ray.y, ray.z * ray.z));
  V3 dc = ray;
  if (sq > 0.0f) {
    const float nrm = sqrtf(sq);
    dc.x = ray.x / nrm;
    dc.y = ray.y / nrm;
    dc.z = ray.z / nrm;
  }
  const V3 dl = dev_rotate(tp.T_L_C, dc);
  const float ox = tp.T_L_C.t[0], oy = tp.T_L_C.t[1], oz = tp.T_L_C.t[2];
  // every block lives inside the workspace grid: leaving it for good ends the march
  const bool closed_world = (m.ws_sx > 0) && (m.ctrl->n_hash == 0);
  const I3 ws_mx = {m.ws_mn.x + m.ws_sx - 1, m.ws_mn.y + m.ws_sy - 1, m.ws_mn.z + m.ws_sz - 1};
  const float bs = m.block_size, bs_inv = 1.0f / m.block_size;
  float2* const slab0 = m.tsdf_slabs[0];

  // One sample of the march: which TSDF voxel does position o + t * d read (nullptr: no block there), and has the
  // ray left a closed world for good?  getBlockAndVoxelIndexFromPositionInLayer (indexing_impl.h:37-49).
  struct Loc {
    const float2* addr;
    bool gone;
    bool free;  // the sample lies in a block of observed free space: its value is (free_dist, valid), no load needed
  };
  auto locate = [&](float tt) -> Loc {
    V3 p;
    p.x = fmaf(tt, dl.x, ox);
    p.y = fmaf(tt, dl.y, oy);
    p.z = fmaf(tt, dl.z, oz);
    I3 b, v;
    b.x = floor_div_exact(p.x, bs, bs_inv);
    b.y = floor_div_exact(p.y, bs, bs_inv);
    b.z = floor_div_exact(p.z, bs, bs_inv);
    v.x = min((int)(fmaf(-(float)b.x, bs, p.x) * m.voxel_size_inv), 7);
    v.y = min((int)(fmaf(-(float)b.y, bs, p.y) * m.voxel_size_inv), 7);
    v.z = min((int)(fmaf(-(float)b.z, bs, p.z) * m.voxel_size_inv), 7);
    // Block lookup on every sample (no per-thread block cache: with the table in shared memory the lookup is
    // cheaper than the divergence a cache introduces).  A slot without a TSDF layer has an all-zero TSDF
    // payload (k_allocate_one), i.e. reads as unobserved, so the layer bits need not be consulted here.
    const int cell = ws_cell(m, b.x, b.y, b.z);
    int slot = -1;
    Loc l;
    l.gone = false;
    l.free = false;
    if (cell >= 0) {
      slot = s_ws ? s_ws[cell] : m.ws_slot[cell];
      if (s_ws && slot >= 0) {  // staged table: bit 30 = the block is observed free space (kBlockFreeBit)
        l.free = (slot & kNewFlag) != 0;
        slot &= kSlotMask;
      }
    } else if (!closed_world) {
      slot = hash_find(m, b.x, b.y, b.z);
    } else {
      // t only grows (steps are >= 0), so each coordinate of p moves monotonically along sign(dl)
      l.gone = (b.x > ws_mx.x && dl.x >= 0.0f) || (b.x < m.ws_mn.x && dl.x <= 0.0f) ||
               (b.y > ws_mx.y && dl.y >= 0.0f) || (b.y < m.ws_mn.y && dl.y <= 0.0f) ||
               (b.z > ws_mx.z && dl.z >= 0.0f) || (b.z < m.ws_mn.z && dl.z <= 0.0f);
    }
    l.addr = nullptr;
    if (slot >= 0)
      l.addr = ((slot < (1 << kTsdfSlabShift)) ? slab0 + (size_t)slot * kVoxelsPerBlock : tsdf_block(m, slot)) +
               ((v.x * 8 + v.y) * 8 + v.z);
    return l;
  };

  // The reference marches one dependent TSDF read per step (sphere_tracer.cu:31-131): <= 24 steps x one L2 round
  // trip each on the slowest rays, which set the kernel's duration.  Here each round trip carries kSpec samples:
  // the true position plus kSpec - 1 positions predicted with the last step length (the truncation distance in
  // unobserved space, the TSDF distance -- nearly constant in observed free space -- otherwise).  A predicted
  // sample is consumed only if the TRUE next position reads the very same voxel address, so the sequence of
  // (t, value) pairs, the step count and the result are exactly those of the one-read-per-step march.
  int first = 0;
  float t = 0.0f;
  bool ok = false;
  int n_steps = 0;
  int i = 0;
  float guess = tp.trunc;
  bool run = (i < tp.max_steps) && (t < tp.max_ray_length);
  while (run) {
    float tk[kSpec];
    Loc loc[kSpec];
    float2 val[kSpec];
    tk[0] = t;
#pragma unroll
    for (int j = 1; j < kSpec; ++j) tk[j] = tk[j - 1] + guess;
#pragma unroll
    for (int j = 0; j < kSpec; ++j) loc[j] = locate(tk[j]);
#pragma unroll
    for (int j = 0; j < kSpec; ++j)
      val[j] = loc[j].free ? make_float2(tp.free_dist, 1.0f) : (loc[j].addr ? *loc[j].addr : make_float2(0.0f, 0.0f));
#pragma unroll
    for (int k = 0; k < kSpec; ++k) {
      Loc cur = loc[k];
      if (k > 0) {
        cur = locate(t);
        if (cur.addr != nullptr && cur.addr != loc[k].addr) break;  // predicted another voxel: refill from t
      }
      n_steps = i + 1;
      if (cur.gone) {  // miss
        run = false;
        break;
      }
      bool valid = false;
      float dist = 0.0f;
      if (cur.addr) {
        const float2 q = val[k];
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
          run = false;  // left observed space: fail
          break;
        }
      } else {
        if (first == 0) first = (dist >= 0.0f) ? 1 : -1;
        if (first == 1) {
          if (dist < tp.eps) {
            t += dist;
            ok = true;
            run = false;
            break;
          }
          step = dist;
        } else {
          if (dist > -tp.eps) {
            t -= dist;
            ok = true;
            run = false;
            break;
          }
          step = -dist;
        }
      }
      t += step;
      ++i;
      guess = step;
      if (!((i < tp.max_steps) && (t < tp.max_ray_length))) {
        run = false;
        break;
      }
    }
  }
  image[(size_t)r * tp.cols + c] = ok ? t * dc.z : -1.0f;
  return n_steps;
}

// ---- cached march: the TSDF blocks a warp's rays cross are staged in shared memory ------------------------------
// The scalar march above pays one L2 round trip per sample that falls into a block near the surface, and those are
// the samples of the slow rays (the step shrinks with the distance, 10 .. 24 dependent round trips, each ~1 us while
// the previous frame's gather saturates the memory system).  The 32 rays of a warp (an 8 x 4 patch of the synthetic
// image) cross the same one or two blocks there.  Here every warp keeps E whole TSDF blocks (4 KB each) in shared
// memory.  A block is staged when a lane samples it twice in a row (the ray lingers: it is near the surface), by
// cp.async in the background of that step's ordinary loads, so staging never adds a round trip; every later sample of
// any lane into that block is a shared-memory read, and a step in which all lanes hit costs no round trip at all.
// Same positions, same values, same decisions as the scalar march: only where a voxel is read from changes.
// Warp-synchronous (no CTA barrier): all 32 lanes stay in the loop until the last ray of the warp is done.
template <int E>
__device__ __forceinline__ void sphere_trace_ray_cached(const MapDev& m, const TraceParams& tp, int r, int c,
                                                        bool in_image, float* __restrict__ image, const int* s_ws,
                                                        float2* cache, int* tags) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const float pu = (float)(c * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const float pv = (float)(r * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const V3 ray = ray_from_image_plane(tp.cam, pu, pv);
  const float sq = fmaf(ray.x, ray.x, fmaf(ray.y, ray.y, ray.z * ray.z));
  V3 dc = ray;
  if (sq > 0.0f) {
    const float nrm = sqrtf(sq);
    dc.x = ray.x / nrm;
    dc.y = ray.y / nrm;
    dc.z = ray.z / nrm;
  }
  const V3 dl = dev_rotate(tp.T_L_C, dc);
  const float ox = tp.T_L_C.t[0], oy = tp.T_L_C.t[1], oz = tp.T_L_C.t[2];
  const bool closed_world = (m.ws_sx > 0) && (m.ctrl->n_hash == 0);
  const I3 ws_mx = {m.ws_mn.x + m.ws_sx - 1, m.ws_mn.y + m.ws_sy - 1, m.ws_mn.z + m.ws_sz - 1};
  const float bs = m.block_size, bs_inv = 1.0f / m.block_size;

  if (lane < E) tags[lane] = -1;
  __syncwarp();
  int victim = 0;      // round-robin replacement (warp-uniform)
  int prev_slot = -1;  // the block this lane's previous sample fell into

  int first = 0;
  float t = 0.0f;
  bool ok = false;
  int i = 0;
  bool run = in_image && (i < tp.max_steps) && (t < tp.max_ray_length);
  while (__any_sync(kFull, run)) {
    // which voxel does o + t * d read?  (getBlockAndVoxelIndexFromPositionInLayer, indexing_impl.h:37-49)
    int slot = -1, vox = 0;
    bool gone = false, free_blk = false;
    if (run) {
      V3 p;
      p.x = fmaf(t, dl.x, ox);
      p.y = fmaf(t, dl.y, oy);
      p.z = fmaf(t, dl.z, oz);
      I3 b, v;
      b.x = floor_div_exact(p.x, bs, bs_inv);
      b.y = floor_div_exact(p.y, bs, bs_inv);
      b.z = floor_div_exact(p.z, bs, bs_inv);
      v.x = min((int)(fmaf(-(float)b.x, bs, p.x) * m.voxel_size_inv), 7);
      v.y = min((int)(fmaf(-(float)b.y, bs, p.y) * m.voxel_size_inv), 7);
      v.z = min((int)(fmaf(-(float)b.z, bs, p.z) * m.voxel_size_inv), 7);
      vox = (v.x * 8 + v.y) * 8 + v.z;
      const int cell = ws_cell(m, b.x, b.y, b.z);
      if (cell >= 0) {
        slot = s_ws ? s_ws[cell] : m.ws_slot[cell];
        if (s_ws && slot >= 0) {
          free_blk = (slot & kNewFlag) != 0;
          slot &= kSlotMask;
        }
      } else if (!closed_world) {
        slot = hash_find(m, b.x, b.y, b.z);
      } else {
        gone = (b.x > ws_mx.x && dl.x >= 0.0f) || (b.x < m.ws_mn.x && dl.x <= 0.0f) ||
               (b.y > ws_mx.y && dl.y >= 0.0f) || (b.y < m.ws_mn.y && dl.y <= 0.0f) ||
               (b.z > ws_mx.z && dl.z >= 0.0f) || (b.z < m.ws_mn.z && dl.z <= 0.0f);
      }
    }
    // The voxel's value.  A block already staged for this warp is read from shared memory; any other voxel is loaded
    // straight from global memory exactly like the scalar march does (all lanes' loads in one round trip), and a block
    // a lane samples for the SECOND time in a row -- the ray lingers there: it is close to the surface -- is staged in
    // the background (cp.async, in flight together with those loads) for the samples still to come.
    const bool need = run && slot >= 0 && !free_blk;
    float2 q = make_float2(0.0f, 0.0f);
    int hit = -1;
    if (need) {
#pragma unroll
      for (int k = 0; k < E; ++k)
        if (tags[k] == slot) hit = k;
    }
    const float2* gaddr = nullptr;
    if (need) {
      if (hit >= 0)
        q = cache[(size_t)hit * kVoxelsPerBlock + vox];
      else
        gaddr = tsdf_block(m, slot) + vox;
    }
    __syncwarp();  // the staged entries have been read: they may be replaced now
    unsigned want = __ballot_sync(kFull, need && hit < 0 && slot == prev_slot);
    bool staged_any = false;
    int n_staged = 0;  // at most E blocks per step: two copies in flight must never target the same entry
    while (want && n_staged < E) {
      const int s = __shfl_sync(kFull, slot, __ffs(want) - 1);
      bool present = false;
#pragma unroll
      for (int k = 0; k < E; ++k) present |= (tags[k] == s);
      if (!present) {
        const int e = victim;
        victim = (victim + 1 == E) ? 0 : victim + 1;
        const uint4* src = reinterpret_cast<const uint4*>(tsdf_block(m, s));
        const unsigned dst = (unsigned)__cvta_generic_to_shared(cache + (size_t)e * kVoxelsPerBlock);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(lane + 32 * j) * 16u),
                       "l"(src + lane + 32 * j));
        __syncwarp();
        if (lane == 0) tags[e] = s;
        staged_any = true;
        ++n_staged;
      }
      want &= ~__ballot_sync(kFull, need && slot == s);
      __syncwarp();
    }
    if (gaddr) q = *gaddr;
    prev_slot = run ? slot : -1;
    if (staged_any) {  // warp-uniform
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    if (run) {  // sphereTracingKernel's step (sphere_tracer.cu:31-131), exactly as in sphere_trace_ray
      if (gone) {
        run = false;
      } else {
        bool valid = false;
        float dist = 0.0f;
        if (slot >= 0) {
          if (free_blk) q = make_float2(tp.free_dist, 1.0f);
          if (q.y > 1e-4f) {
            valid = true;
            dist = q.x;
          }
        }
        float step = 0.0f;
        if (!valid) {
          if (first == 0) {
            step = tp.trunc;
          } else {
            run = false;  // left observed space: fail
          }
        } else {
          if (first == 0) first = (dist >= 0.0f) ? 1 : -1;
          if (first == 1) {
            if (dist < tp.eps) {
              t += dist;
              ok = true;
              run = false;
            } else {
              step = dist;
            }
          } else {
            if (dist > -tp.eps) {
              t -= dist;
              ok = true;
              run = false;
            } else {
              step = -dist;
            }
          }
        }
        if (run) {
          t += step;
          ++i;
          if (!((i < tp.max_steps) && (t < tp.max_ray_length))) run = false;
        }
      }
    }
  }
  if (in_image) image[(size_t)r * tp.cols + c] = ok ? t * dc.z : -1.0f;
}

// ---- team march: G lanes share one ray ---------------------------------------------------------------------
// The scalar march above is a chain of <= ~24 dependent (address arithmetic + L2 round trip) steps per ray, and the
// slowest rays set the kernel's duration.  Here a team of G lanes works on ONE ray: in every round lane j samples
// the position the march WOULD reach after j more steps of the last step length (t_j = t + guess + ... + guess,
// the same sequence of float additions the scalar march performs when its steps stay equal), all G loads fly at
// once, and the samples are then consumed in order for as long as the true position equals the sampled one BIT FOR
// BIT -- i.e. for as long as every consumed step really had the predicted length.  Through unobserved space (step =
// truncation distance) and observed free space (step = the clamped TSDF value) that is G steps per round trip; next
// to the surface, where the steps shrink, it degenerates to the scalar march.  The sequence of (t, value) pairs, the
// termination rule and the result are exactly the scalar march's (and the reference's, sphere_tracer.cu:31-131).
template <int G>
__device__ __forceinline__ void sphere_trace_team(const MapDev& m, const TraceParams& tp, int r, int c, bool in_image,
                                                  float* __restrict__ image, const int* s_ws) {
  const int tl = threadIdx.x & (G - 1);  // lane within the team
  const float pu = (float)(c * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const float pv = (float)(r * tp.sub) + 0.5f * (float)tp.sub * 1.0f;
  const V3 ray = ray_from_image_plane(tp.cam, pu, pv);
  const float sq = fmaf(ray.x, ray.x, fmaf(ray.y, ray.y, ray.z * ray.z));
  V3 dc = ray;
  if (sq > 0.0f) {
    const float nrm = sqrtf(sq);
    dc.x = ray.x / nrm;
    dc.y = ray.y / nrm;
    dc.z = ray.z / nrm;
  }
  const V3 dl = dev_rotate(tp.T_L_C, dc);
  const float ox = tp.T_L_C.t[0], oy = tp.T_L_C.t[1], oz = tp.T_L_C.t[2];
  const bool closed_world = (m.ws_sx > 0) && (m.ctrl->n_hash == 0);
  const I3 ws_mx = {m.ws_mn.x + m.ws_sx - 1, m.ws_mn.y + m.ws_sy - 1, m.ws_mn.z + m.ws_sz - 1};
  const float bs = m.block_size, bs_inv = 1.0f / m.block_size;
  float2* const slab0 = m.tsdf_slabs[0];

  int first = 0, i = 0;
  float t = 0.0f, guess = tp.trunc;
  bool ok = false;
  bool run = in_image && (i < tp.max_steps) && (t < tp.max_ray_length);
  while (__any_sync(0xffffffffu, run)) {
    // this lane's speculative sample position: t advanced by tl steps of the predicted length
    float tj = t;
#pragma unroll
    for (int k = 0; k < G - 1; ++k)
      if (k < tl) tj += guess;
    // flags: 1 = has a voxel address, 2 = ray has left a closed world for good, 4 = block of observed free space
    int flags = 0;
    float2 val = make_float2(0.0f, 0.0f);
    if (run && (i + tl < tp.max_steps) && (tj < tp.max_ray_length)) {
      V3 p;
      p.x = fmaf(tj, dl.x, ox);
      p.y = fmaf(tj, dl.y, oy);
      p.z = fmaf(tj, dl.z, oz);
      I3 b, v;
      b.x = floor_div_exact(p.x, bs, bs_inv);
      b.y = floor_div_exact(p.y, bs, bs_inv);
      b.z = floor_div_exact(p.z, bs, bs_inv);
      v.x = min((int)(fmaf(-(float)b.x, bs, p.x) * m.voxel_size_inv), 7);
      v.y = min((int)(fmaf(-(float)b.y, bs, p.y) * m.voxel_size_inv), 7);
      v.z = min((int)(fmaf(-(float)b.z, bs, p.z) * m.voxel_size_inv), 7);
      const int cell = ws_cell(m, b.x, b.y, b.z);
      int slot = -1;
      if (cell >= 0) {
        slot = s_ws ? s_ws[cell] : m.ws_slot[cell];
        if (s_ws && slot >= 0) {
          if (slot & kNewFlag) flags |= 4;
          slot &= kSlotMask;
        }
      } else if (!closed_world) {
        slot = hash_find(m, b.x, b.y, b.z);
      } else if ((b.x > ws_mx.x && dl.x >= 0.0f) || (b.x < m.ws_mn.x && dl.x <= 0.0f) ||
                 (b.y > ws_mx.y && dl.y >= 0.0f) || (b.y < m.ws_mn.y && dl.y <= 0.0f) ||
                 (b.z > ws_mx.z && dl.z >= 0.0f) || (b.z < m.ws_mn.z && dl.z <= 0.0f)) {
        flags |= 2;
      }
      if (slot >= 0) {
        flags |= 1;
        if (flags & 4) {
          val = make_float2(tp.free_dist, 1.0f);
        } else {
          const float2* addr =
              ((slot < (1 << kTsdfSlabShift)) ? slab0 + (size_t)slot * kVoxelsPerBlock : tsdf_block(m, slot)) +
              ((v.x * 8 + v.y) * 8 + v.z);
          val = *addr;
        }
      }
    }
    // consume the samples in order (every lane of the team replays the same state machine)
    bool open = run;           // samples of this round may still be consumed
    float next_guess = guess;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const float tj_b = __shfl_sync(0xffffffffu, tj, j, G);
      const float vx_b = __shfl_sync(0xffffffffu, val.x, j, G);
      const float vy_b = __shfl_sync(0xffffffffu, val.y, j, G);
      const int fl_b = __shfl_sync(0xffffffffu, flags, j, G);
      if (open) {
        if (__float_as_int(tj_b) != __float_as_int(t)) {
          open = false;        // an earlier step was not of the predicted length: lanes >= j sampled elsewhere
        } else if (!((i < tp.max_steps) && (t < tp.max_ray_length))) {
          open = run = false;  // ran out of steps or length: fail
        } else if (fl_b & 2) {
          open = run = false;  // miss
        } else {
          const bool valid = (fl_b & 1) && (vy_b > 1e-4f);
          float step = 0.0f;
          if (!valid) {
            if (first == 0) {
              step = tp.trunc;
            } else {
              open = run = false;  // left observed space: fail
            }
          } else {
            if (first == 0) first = (vx_b >= 0.0f) ? 1 : -1;
            if (first == 1) {
              if (vx_b < tp.eps) {
                t += vx_b;
                ok = true;
                open = run = false;
              } else {
                step = vx_b;
              }
            } else {
              if (vx_b > -tp.eps) {
                t -= vx_b;
                ok = true;
                open = run = false;
              } else {
                step = -vx_b;
              }
            }
          }
          if (open) {
            t += step;
            ++i;
            next_guess = step;
          }
        }
      }
    }
    guess = next_guess;
    if (run && !((i < tp.max_steps) && (t < tp.max_ray_length))) run = false;
  }
  if (in_image && tl == 0) image[(size_t)r * tp.cols + c] = ok ? t * dc.z : -1.0f;
}

// The two independent, latency-bound preparations of a feature frame in ONE launch: CTAs
// [0, n_trace_ctas) sphere-trace 16x16 tiles of the synthetic depth image, the remaining CTAs run
// band_select_tile.  Both only read the TSDF layer; they overlap instead of queueing.
template <int SPEC, int ILP>
__global__ void __launch_bounds__(256) k_trace_and_band(MapDev m, TraceParams tp, float* __restrict__ image,
                                                        int trace_tiles_x, int n_trace_ctas, PlanesView view,
                                                        float trunc, int* band_slots, int* newfeat_slots,
                                                        int tile_cells, int n_tiles, int color_parity) {
  pdl_prologue();
  __shared__ int s_cand[256], s_band[256];
  __shared__ int s_ncand, s_nband;
  __shared__ int s_ws[kTraceSmemCells];
#ifdef NVBX_PROFILE_COUNTERS
  const long long t0 = clock64();
#endif
  PROF_BEGIN(m.seq, kProfTrace);
  // this ring slot's item list: last read by the gather of frame i - kFrameRing, complete before we were enqueued
  if (blockIdx.x == 0 && threadIdx.x == 0 && color_parity == -1) m.ctrl->item_count[m.fp] = 0;
  // n_trace_ctas == 0: the synthetic depth image of this pose / camera / TSDF state is already in `image`
  if ((int)blockIdx.x < n_trace_ctas) {
    const int c = (blockIdx.x % trace_tiles_x) * 16 + (threadIdx.x & 7) + ((threadIdx.x >> 7) << 3);
    const int r = (blockIdx.x / trace_tiles_x) * 16 + ((threadIdx.x >> 3) & 15);
    const bool stage_ws = m.ws_cells > 0 && m.ws_cells <= kTraceSmemCells;
    if (stage_ws) {
      // four cells per thread and round: the slot loads of a round fly together, then the layer-byte loads that depend
      // on them (two L2 round trips per round instead of two per cell)
      for (int i0 = threadIdx.x; i0 < m.ws_cells; i0 += 4 * 256) {
        int sl[4];
        uint8_t ly[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * 256;
          sl[k] = i < m.ws_cells ? m.ws_slot[i] : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) ly[k] = (tp.use_free && sl[k] >= 0) ? m.blk_layers[sl[k]] : (uint8_t)0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * 256;
          if (i < m.ws_cells) s_ws[i] = (ly[k] & kBlockFreeBit) ? (sl[k] | kNewFlag) : sl[k];  // bit 30: observed free space
        }
      }
      __syncthreads();
    }
    if (tp.team) {  // 32 rays (8 x 4) per CTA, eight lanes per ray
      const int ray = threadIdx.x >> 3;
      const int tc = (blockIdx.x % trace_tiles_x) * 8 + (ray & 7);
      const int tr = (blockIdx.x / trace_tiles_x) * 4 + (ray >> 3);
      sphere_trace_team<8>(m, tp, tr, tc, tr < tp.rows && tc < tp.cols, image, stage_ws ? s_ws : nullptr);
      PROF_END(kProfTrace);
      return;
    }
    if (tp.cache_blocks) {  // whole TSDF blocks staged per warp (dynamic shared memory)
      extern __shared__ uint4 s_dyn[];
      const int warp = threadIdx.x >> 5, nb = tp.cache_blocks;
      float2* cache = reinterpret_cast<float2*>(s_dyn) + (size_t)warp * nb * kVoxelsPerBlock;
      int* tags = reinterpret_cast<int*>(reinterpret_cast<float2*>(s_dyn) + (size_t)8 * nb * kVoxelsPerBlock) + warp * nb;
      const bool in_image = r < tp.rows && c < tp.cols;
      if (nb == 4)
        sphere_trace_ray_cached<4>(m, tp, r, c, in_image, image, stage_ws ? s_ws : nullptr, cache, tags);
      else
        sphere_trace_ray_cached<2>(m, tp, r, c, in_image, image, stage_ws ? s_ws : nullptr, cache, tags);
      PROF_END(kProfTrace);
      return;
    }
    [[maybe_unused]] int n_steps = 0;
    if (tp.march < 256) {
      if ((int)threadIdx.x >= tp.march) return;
      const int tc = (blockIdx.x % trace_tiles_x) * 8 + (threadIdx.x & 7);
      const int tr = (blockIdx.x / trace_tiles_x) * (tp.march >> 3) + (threadIdx.x >> 3);
      if (tr < tp.rows && tc < tp.cols) n_steps = sphere_trace_ray<SPEC>(m, tp, tr, tc, image, stage_ws ? s_ws : nullptr);
    } else if (r < tp.rows && c < tp.cols) n_steps = sphere_trace_ray<SPEC>(m, tp, r, c, image, stage_ws ? s_ws : nullptr);
#ifdef NVBX_PROFILE_COUNTERS
    {  // tuning aid: steps (sum / max over rays) and the slowest warp's cycles, one atomic set per warp
      const long long dt = clock64() - t0;
      int sum = n_steps, mx = n_steps;
      for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (lane_id() == 0) {
        atomicAdd(&m.ctrl->counters[kCntProfile0], (unsigned long long)sum);
        atomicMax(&m.ctrl->counters[kCntProfile0 + 1], (unsigned long long)mx);
        atomicMax(&m.ctrl->counters[kCntProfile0 + 2], (unsigned long long)dt);
      }
    }
#else
    (void)n_steps;
#endif
    PROF_END(kProfTrace);
    return;
  }
  for (int tile = (int)blockIdx.x - n_trace_ctas; tile < n_tiles; tile += (int)gridDim.x - n_trace_ctas)
    band_select_tile<ILP>(m, view, trunc, band_slots, newfeat_slots, tile, tile_cells, s_cand, &s_ncand, s_band,
                          &s_nband, color_parity);
#ifdef NVBX_PROFILE_COUNTERS
  if (threadIdx.x == 0) atomicMax(&m.ctrl->counters[kCntProfile0 + 3], (unsigned long long)(clock64() - t0));
#endif
  PROF_END(kProfTrace);
}

// The read-only twin of the band selection: ctrl->list_count += band blocks of this view that have no feature
// block yet.  Launched only when the pessimistic bound of the frame (every candidate block gets a feature
// block) would not fit the feature arena and is too large to simply allocate (ensure_feats): the arena then
// grows by the exact figure.
__global__ void __launch_bounds__(256) k_band_count(MapDev m, PlanesView view, float trunc, int tile_cells,
                                                    int n_tiles) {
  pdl_prologue();
  __shared__ int s_cand[256], s_band[256];
  __shared__ int s_ncand, s_nband;
  for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x)
    band_select_tile<1>(m, view, trunc, nullptr, nullptr, tile, tile_cells, s_cand, &s_ncand, s_band, &s_nband,
                        kBandCountOnly);
}

// ================================================================================================
// a8. Feature integration -- THE hot path, in two kernels.
//
// Reference: integrateBlocksKernel<UpdateAppearanceVoxelFunctor<FeatureVoxel>> (projective_integrator_
// impl.cuh:156-214) runs one THREAD per voxel that serially walks all C channels through three
// 1.5 KB per-thread arrays and a 2-byte-aligned 1538-byte AoS record.  Here:
//
//   k_feature_geometry (thread = voxel, CTA = band block; fp32 geometry only): project, bilinear
//     synthetic depth, band test, image bounds, mask, old weight -> 16-byte work items appended to ONE
//     global list (one atomicAdd per block).  Only ~1/3 of a band block's voxels survive and there are
//     fewer band blocks than SMs on the mindmap tasks, so doing the byte-moving per block would leave
//     most of the machine idle; the list is what balances it.
//   k_feature_gather (warp = 512-byte chunk of one item's channel vector; persistent, grid = SMs x 4):
//     four 128-bit read-only loads per lane (the 4 bilinear neighbours, each C contiguous halves in the
//     HWC image), fp16 interpolation / blend in the reference's operation order on half2 lanes, one
//     128-bit store into the voxel's 16-byte-aligned row.  Units are dealt round-robin over all warps
//     of the grid and every lane keeps two units (eight loads) in flight.
//
// Algorithmic bytes per updated voxel: 2C x (distinct pixels, <= 4) read + 2(C+1) written
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
  int stream_loads;           // 1: pixel rows are loaded without allocating in L1 (k_feature_gather_dyn)
};

__device__ __forceinline__ __half2 interp_h2(__half2 x, __half2 y, __half2 xy, __half2 f00, __half2 f01, __half2 f10,
                                             __half2 f11) {
  // the reference's SASS: HFMA2(x, dx, f00); HFMA2(y, f01 - f00, .); HFMA2(x*y, f11 - f01 - dx, .)
  const __half2 dx = __hsub2(f10, f00);
  const __half2 t2 = __hfma2(x, dx, f00);
  const __half2 t5 = __hfma2(y, __hsub2(f01, f00), t2);
  return __hfma2(xy, __hsub2(__hsub2(f11, f01), dx), t5);
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
  for (int i = 0; i < 4; ++i) pr[i] = __hfma2(po[i], w1, __hmul2_rn(pm[i], w2));   // HMUL2 + HFMA2
  return o;
}
__device__ __forceinline__ uint4 ldg_nc(const uint4* p) { return __ldg(p); }
// The same read-only load without allocating the line in L1: the gather streams every pixel row through once, and while
// it shares the SMs with the next frame's depth path (frame pipelining) those lines only push that path's TSDF / table
// lines out of L1.
__device__ __forceinline__ uint4 ldg_nc_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// One voxel to update (16 bytes, moved as one uint4).
struct __align__(16) FeatItem {
  int pix;                // (ly * cols + lx): offset of the top-left neighbour in pixels
  unsigned short hx, hy;  // half bits of the interpolation offsets
  unsigned short wnew;    // half bits of the new weight
  unsigned short first;   // 1: first observation (copy, no blend)
  int row;                // feature voxel row: feature_slot * 512 + voxel
};

__global__ void __launch_bounds__(512) k_feature_geometry(MapDev m, const int* __restrict__ band_slots,
                                                          const int* __restrict__ newfeat_slots, FeatFrame f,
                                                          FeatItem* __restrict__ items, int block_begin,
                                                          int block_end) {
  pdl_prologue();
  __shared__ int s_warp_base[17];
  __shared__ int s_base;
  const int n = min(m.ctrl->band_count[m.fp], block_end);
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int vx = t >> 6, vy = (t >> 3) & 7, vz = t & 7;
  const int C = m.C;
  unsigned long long n_updated = 0;
  if (blockIdx.x == 0 && t == 0) m.ctrl->gather_ticket[m.fp] = 0;  // consumed by this frame's k_feature_gather_dyn
  PROF_BEGIN(m.seq, kProfGeometry);

  // Zero-fill the feature blocks allocated by this frame (blox_impl.h:92-97), every CTA taking an equal
  // slice of each, so that a 794 KB block costs each SM a few KB; the gather kernel runs after us.
  if (block_begin == 0) {
    // The weight vector of each row (the uint4 behind the C feature halves) is NOT touched here: the CTA that
    // owns the block writes all 512 of them below, possibly before another CTA's slice of the fill gets there.
    const int n_new = m.ctrl->newfeat_count[m.fp];
    const int row_vecs_z = m.row >> 3, wvec = C >> 3;
    const int vec_per_block = kVoxelsPerBlock * row_vecs_z;
    const int per_cta = (vec_per_block + gridDim.x - 1) / gridDim.x;
    const int k0 = blockIdx.x * per_cta, k1 = min(vec_per_block, k0 + per_cta);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int j = 0; j < n_new; ++j) {
      uint4* p = reinterpret_cast<uint4*>(feat_block(m, newfeat_slots[j]));
      for (int k = k0 + t; k < k1; k += 512)
        if (k % row_vecs_z != wvec) p[k] = z;
    }
  }

  for (int bi = block_begin + blockIdx.x; bi < n; bi += gridDim.x) {
    const int raw = band_slots[bi];
    const bool is_new = (raw & kNewFlag) != 0;
    const int slot = raw & kSlotMask;
    const int3 b = m.blk_index[slot];
    const int fs = m.blk_feat[slot];
    __half* blk = feat_block(m, fs);
    if (t == 0) m.blk_dirty[slot] = kDirtyAll;  // mapper.cpp:462

    bool active = false;
    FeatItem it;
    it.pix = 0;
    it.hx = it.hy = it.wnew = it.first = 0;
    it.row = fs * kVoxelsPerBlock + t;
    do {
      float u, v, vd;
      if (!project_voxel(f.cam, f.T_C_L, m.block_size, f.max_depth, b, vx, vy, vz, &u, &v, &vd)) break;
      // occlusion test against the synthetic depth (bilinear, no validity check)
      const float ud = u / (float)f.sub, vdp = v / (float)f.sub;
      const float uc = ud - 0.5f, vc = vdp - 0.5f;
      const int lx = (int)floorf(uc), ly = (int)floorf(vc);
      if (lx < 0 || ly < 0 || (lx + 1) > (f.scols - 1) || (ly + 1) > (f.srows - 1)) break;
      const float* sp = f.synth + (size_t)ly * f.scols + lx;
      const float surface = dev_interp_float(uc - (float)lx, vc - (float)ly, sp[0], sp[f.scols], sp[1], sp[f.scols + 1]);
      if (fabsf(surface - vd) > f.trunc) break;
      const float fu = u - 0.5f, fv = v - 0.5f;
      const int px = (int)floorf(fu), py = (int)floorf(fv);
      if (px < 0 || py < 0 || (px + 1) > (f.cols - 1) || (py + 1) > (f.rows - 1)) break;
      if (f.mask != nullptr && !__ldg(f.mask + (size_t)((int)v) * f.cols + (int)u)) break;
      const float w_cur = is_new ? 0.0f : __half2float(blk[(size_t)t * m.row + C]);
      it.pix = py * f.cols + px;
      it.hx = __half_as_ushort(__float2half_rn(fu - (float)px));
      it.hy = __half_as_ushort(__float2half_rn(fv - (float)py));
      it.wnew = __half_as_ushort(__float2half_rn(fminf(f.alpha + w_cur, f.max_weight)));
      it.first = (w_cur == 0.0f) ? 1 : 0;
      active = true;
    } while (false);

    // The voxel's new weight (+ the zero padding of its row) is written HERE, not by the gather: the next frame's
    // geometry reads it, and must not depend on a gather that may still be running on another stream.
    if (active || is_new)
      *(reinterpret_cast<uint4*>(blk + (size_t)t * m.row) + (C >> 3)) = make_uint4((unsigned)it.wnew, 0u, 0u, 0u);
    // block-level compaction: ballot + 16-entry scan, one atomicAdd per block on the global list
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
      s_base = acc ? atomicAdd(&m.ctrl->item_count[m.fp], acc) : 0;
      n_updated += (unsigned long long)acc;
    }
    __syncthreads();
    if (active)
      *reinterpret_cast<uint4*>(&items[s_base + s_warp_base[warp] + __popc(ballot & ((1u << lane) - 1u))]) =
          *reinterpret_cast<const uint4*>(&it);
    __syncthreads();  // s_warp_base / s_base are reused by the next block
  }
  if (t == 0) {
    if (n_updated) count_add(m, kCntFeatVoxelsUpdated, n_updated);
    if (blockIdx.x == 0 && block_begin == 0) {
      count_add(m, kCntFeatBandBlocks, (unsigned long long)m.ctrl->band_count[m.fp]);
      count_add(m, kCntFeatureFrames, 1);
      m.ctrl->last_band_count = m.ctrl->band_count[m.fp];  // debug / parity hook (nvbx_debug_last_block_list)
    }
  }
  PROF_END(kProfGeometry);
}

// CH: 512-byte chunks (32 lanes x 16 B) per channel vector, C = 256 * CH; 0 = any C (multiple of 8).
template <int CH, int U, int CTAS>  // U: units in flight per warp; CTAS: resident CTAs per SM (register budget)
__global__ void __launch_bounds__(256, CTAS) k_feature_gather(MapDev m, const FeatItem* __restrict__ items,
                                                              FeatFrame f, int last_chunk) {
  pdl_prologue();
  PROF_BEGIN(m.seq, kProfGather);
  const int n_items = m.ctrl->item_count[m.fp];
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int C = m.C;
  const int nvec = C >> 3;
  const int ch_per_item = CH > 0 ? CH : ((nvec + 31) >> 5);
  const long long n_units = (long long)n_items * ch_per_item;
  const __half2 w1 = __half2half2(__ushort_as_half(f.h_w1));
  const __half2 w2 = __half2half2(__ushort_as_half(f.h_w2));
  const size_t row_vecs = (size_t)(m.row >> 3);

  for (long long q0 = warp; q0 < n_units; q0 += (long long)U * warps_total) {
    uint4 a00[U], a01[U], a10[U], a11[U], old[U];
    FeatItem it[U];
    int cvec[U];
    bool live[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long q = q0 + (long long)k * warps_total;
      live[k] = q < n_units;
      const int item = live[k] ? (int)(q / ch_per_item) : 0;
      cvec[k] = (int)(q - (long long)item * ch_per_item) * 32 + lane;
      if (CH == 0 && cvec[k] >= nvec) live[k] = false;
      *reinterpret_cast<uint4*>(&it[k]) = __ldg(reinterpret_cast<const uint4*>(items + item));
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (live[k]) {
        const uint4* p00 = reinterpret_cast<const uint4*>(f.img + (size_t)it[k].pix * C) + cvec[k];
        const uint4* p01 = p00 + (size_t)f.cols * nvec;
        a00[k] = ldg_nc(p00);
        a10[k] = ldg_nc(p00 + nvec);
        a01[k] = ldg_nc(p01);
        a11[k] = ldg_nc(p01 + nvec);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (live[k]) {
        const int fslot = it[k].row >> 9, vox = it[k].row & 511;
        uint4* dst = reinterpret_cast<uint4*>(feat_block(m, fslot)) + (size_t)vox * row_vecs;
        const bool blend = (!it[k].first) && f.read_old;
        if (blend) old[k] = dst[cvec[k]];
        const __half hx = __ushort_as_half(it[k].hx), hy = __ushort_as_half(it[k].hy);
        uint4 o = interp_vec(__half2half2(hx), __half2half2(hy), __half2half2(__hmul_rn(hx, hy)), a00[k], a01[k],
                             a10[k], a11[k]);
        if (blend) o = blend_vec(old[k], o, w1, w2);
        dst[cvec[k]] = o;
      }
    }
  }
  if (last_chunk && blockIdx.x == 0 && threadIdx.x == 0) {
    // The band list of this frame was consumed by its geometry kernel (complete: we run behind it); the next writer
    // of this ring slot's counters is the band selection of frame i + kFrameRing, enqueued only after this gather.
    m.ctrl->band_count[m.fp] = 0;
    m.ctrl->newfeat_count[m.fp] = 0;
  }
  PROF_END(kProfGather);
}

// Same units, same arithmetic, different SCHEDULE: the first `n_static` units are dealt round-robin as above,
// the rest are handed out through an atomic ticket (`tk` units per grab) so that no warp idles while others
// still drain their share -- under a saturated HBM queue the per-warp finishing times of a static deal spread
// by a few microseconds, which is a sixth of a 24 us kernel.  The next unit's work item (and, in the dynamic
// part, its ticket) is fetched while the current unit's four pixel loads are in flight, and the first item
// load does not wait for item_count (clamped to the list's capacity, discarded if past the end).
// `ticket` is m.ctrl->gather_ticket, zeroed by k_feature_geometry (the launch before us in stream order).
template <int CH, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k_feature_gather_dyn(MapDev m, const FeatItem* __restrict__ items,
                                                                      int items_cap, FeatFrame f, int last_chunk,
                                                                      int dyn_permille, int tk) {
  pdl_prologue();
  PROF_BEGIN(m.seq, kProfGather);
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (THREADS >> 5);
  const int warp = blockIdx.x * (THREADS >> 5) + (threadIdx.x >> 5);
  const int C = m.C;
  const int nvec = C >> 3;
  const int ch_per_item = CH > 0 ? CH : ((nvec + 31) >> 5);
  FeatItem cur;  // speculative: issued together with the item_count load
  *reinterpret_cast<uint4*>(&cur) =
      __ldg(reinterpret_cast<const uint4*>(items + min(warp / ch_per_item, items_cap - 1)));
  const int n_items = m.ctrl->item_count[m.fp];
  const long long n_units = (long long)n_items * ch_per_item;
  // dyn_permille < 0: CTA-blocked deal -- CTA c owns the contiguous units [c * n / G, (c + 1) * n / G), i.e. work
  // items of neighbouring voxels (the list is in block / voxel order), whose bilinear footprints overlap: the
  // repeated pixel rows then hit in this SM's L1 instead of travelling from L2 again.
  const bool blocked = dyn_permille < 0;
  const long long rounds = (n_units * (1000 - max(dyn_permille, 0)) / 1000) / warps_total;
  const long long n_static = dyn_permille <= 0 ? n_units : rounds * warps_total;  // 0: no ticket at all
  const __half2 w1 = __half2half2(__ushort_as_half(f.h_w1));
  const __half2 w2 = __half2half2(__ushort_as_half(f.h_w2));
  const size_t row_vecs = (size_t)(m.row >> 3);
  int* ticket = &m.ctrl->gather_ticket[m.fp];

  long long q = warp;
  long long stride = warps_total, end = n_static;
  int left = 0;  // units of the current ticket not yet started
  if (blocked) {
    const long long lo = ((long long)blockIdx.x * n_units) / gridDim.x;
    end = ((long long)(blockIdx.x + 1) * n_units) / gridDim.x;
    stride = THREADS >> 5;
    q = lo + (threadIdx.x >> 5);
    if (q < end) {
      *reinterpret_cast<uint4*>(&cur) = __ldg(reinterpret_cast<const uint4*>(items + q / ch_per_item));
    } else {
      q = n_units;
    }
  } else if (q >= n_static) {
    int t = 0;
    if (lane == 0) t = atomicAdd(ticket, tk);
    q = n_static + __shfl_sync(0xffffffffu, t, 0);
    left = tk - 1;
    if (q < n_units) *reinterpret_cast<uint4*>(&cur) = __ldg(reinterpret_cast<const uint4*>(items + q / ch_per_item));
  }
  while (q < n_units) {
    const int item = (int)(q / ch_per_item);
    const int cvec = (int)(q - (long long)item * ch_per_item) * 32 + lane;
    const bool live = CH != 0 || cvec < nvec;
    uint4 a00, a01, a10, a11;
    if (live) {
      const uint4* p00 = reinterpret_cast<const uint4*>(f.img + (size_t)cur.pix * C) + cvec;
      const uint4* p01 = p00 + (size_t)f.cols * nvec;
      if (f.stream_loads) {  // warp-uniform
        a00 = ldg_nc_stream(p00);
        a10 = ldg_nc_stream(p00 + nvec);
        a01 = ldg_nc_stream(p01);
        a11 = ldg_nc_stream(p01 + nvec);
      } else {
        a00 = ldg_nc(p00);
        a10 = ldg_nc(p00 + nvec);
        a01 = ldg_nc(p01);
        a11 = ldg_nc(p01 + nvec);
      }
    }
    // next unit (and its item) while the pixel loads fly
    long long qn;
    if (q + stride < end) {
      qn = q + stride;
    } else if (dyn_permille <= 0) {
      qn = n_units;
    } else if (q >= n_static && left > 0) {
      qn = q + 1;
      --left;
    } else {
      int t = 0;
      if (lane == 0) t = atomicAdd(ticket, tk);
      qn = n_static + __shfl_sync(0xffffffffu, t, 0);
      left = tk - 1;
    }
    FeatItem nxt;
    *reinterpret_cast<uint4*>(&nxt) = make_uint4(0u, 0u, 0u, 0u);
    if (qn < n_units) *reinterpret_cast<uint4*>(&nxt) = __ldg(reinterpret_cast<const uint4*>(items + qn / ch_per_item));
    if (live) {
      const int fslot = cur.row >> 9, vox = cur.row & 511;
      uint4* dst = reinterpret_cast<uint4*>(feat_block(m, fslot)) + (size_t)vox * row_vecs;
      const bool blend = (!cur.first) && f.read_old;
      uint4 old;
      if (blend) old = dst[cvec];
      const __half hx = __ushort_as_half(cur.hx), hy = __ushort_as_half(cur.hy);
      uint4 o = interp_vec(__half2half2(hx), __half2half2(hy), __half2half2(__hmul_rn(hx, hy)), a00, a01, a10, a11);
      if (blend) o = blend_vec(old, o, w1, w2);
      dst[cvec] = o;
    }
    cur = nxt;
    q = qn;
  }
  if (last_chunk && blockIdx.x == 0 && threadIdx.x == 0) {
    // The band list of this frame was consumed by its geometry kernel (complete: we run behind it); the next writer
    // of this ring slot's counters is the band selection of frame i + kFrameRing, enqueued only after this gather.
    m.ctrl->band_count[m.fp] = 0;
    m.ctrl->newfeat_count[m.fp] = 0;
  }
  PROF_END(kProfGather);
}

// ---- bulk-copy (TMA 1-D) variant of the gather --------------------------------------------------------
// The two pixel PAIRS of a voxel's bilinear footprint are 2 x 2C contiguous bytes each in the HWC frame (px and
// px + 1 of one image row).  Here they travel as two cp.async.bulk copies (3 KB each at C = 768) straight into
// shared memory, completion counted by an mbarrier; every warp runs its own ring of stages, the elected lane
// keeps NST voxels (NST x 6 KB) in flight per warp without holding a register for them, and the lanes read the
// landed rows back as conflict-free 128-bit shared loads.  Same arithmetic and stores as k_feature_gather.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

constexpr int kTmaWarps = 8;
template <int CH>
struct TmaGather {
  static constexpr int kRowBytes = 512 * CH;          // one pixel: C halves
  static constexpr int kStageBytes = 4 * kRowBytes;   // (px, px + 1) of the top row, then of the bottom row
  static constexpr int kStages = (24576 / kStageBytes) < 8 ? (24576 / kStageBytes) : 8;
  static constexpr int kSmemBytes = kTmaWarps * kStages * kStageBytes;
};

template <int CH>
__global__ void __launch_bounds__(kTmaWarps * 32, 1) k_feature_gather_tma(MapDev m, const FeatItem* __restrict__ items,
                                                                          FeatFrame f, int last_chunk) {
  using G = TmaGather<CH>;
  constexpr int NST = G::kStages;
  extern __shared__ __align__(128) unsigned char s_stage[];
  __shared__ __align__(8) unsigned long long s_bar[kTmaWarps][NST];
  pdl_prologue();
  PROF_BEGIN(m.seq, kProfGather);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warps_total = (long long)gridDim.x * kTmaWarps;
  const long long warp = (long long)blockIdx.x * kTmaWarps + wid;
  const int n_items = m.ctrl->item_count[m.fp];
  const int nvec = m.C >> 3;
  const size_t row_vecs = (size_t)(m.row >> 3);
  const __half2 w1 = __half2half2(__ushort_as_half(f.h_w1));
  const __half2 w2 = __half2half2(__ushort_as_half(f.h_w2));
  unsigned char* wstage = s_stage + (size_t)wid * NST * G::kStageBytes;
  const unsigned stage0 = smem_u32(wstage), bar0 = smem_u32(&s_bar[wid][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(bar0 + 8u * s, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  FeatItem it[NST];
  long long j = warp;  // next item to put in flight
  auto issue = [&](int s, FeatItem& slot_item) {
    if (j < n_items) {
      *reinterpret_cast<uint4*>(&slot_item) = __ldg(reinterpret_cast<const uint4*>(items + j));
      if (lane == 0) {
        const unsigned bar = bar0 + 8u * s, dst = stage0 + (unsigned)(s * G::kStageBytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(f.img) + (size_t)slot_item.pix * G::kRowBytes;
        mbar_expect_tx(bar, (unsigned)G::kStageBytes);
        bulk_g2s(dst, src, 2u * G::kRowBytes, bar);
        bulk_g2s(dst + 2u * G::kRowBytes, src + (size_t)f.cols * G::kRowBytes, 2u * G::kRowBytes, bar);
      }
    }
    j += warps_total;
  };
#pragma unroll
  for (int s = 0; s < NST; ++s) issue(s, it[s]);

  unsigned parity = 0;
  for (long long k = warp; k < n_items;) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      if (k < n_items) {  // warp-uniform
        mbar_wait(bar0 + 8u * s, parity);
        const FeatItem cur = it[s];
        const uint4* st = reinterpret_cast<const uint4*>(wstage + (size_t)s * G::kStageBytes);
        const int fslot = cur.row >> 9, vox = cur.row & 511;
        uint4* dst = reinterpret_cast<uint4*>(feat_block(m, fslot)) + (size_t)vox * row_vecs;
        const bool blend = (!cur.first) && f.read_old;
        const __half hx = __ushort_as_half(cur.hx), hy = __ushort_as_half(cur.hy);
        const __half2 x2 = __half2half2(hx), y2 = __half2half2(hy), xy2 = __half2half2(__hmul_rn(hx, hy));
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int v = c * 32 + lane;
          const uint4 a00 = st[v], a10 = st[nvec + v], a01 = st[2 * nvec + v], a11 = st[3 * nvec + v];
          uint4 o = interp_vec(x2, y2, xy2, a00, a01, a10, a11);
          if (blend) o = blend_vec(dst[v], o, w1, w2);
          dst[v] = o;
        }
        __syncwarp();  // every lane has read stage s: it may be overwritten
        issue(s, it[s]);
      }
      k += warps_total;
    }
    parity ^= 1u;
  }
  if (last_chunk && blockIdx.x == 0 && threadIdx.x == 0) {
    // The band list of this frame was consumed by its geometry kernel (complete: we run behind it); the next writer
    // of this ring slot's counters is the band selection of frame i + kFrameRing, enqueued only after this gather.
    m.ctrl->band_count[m.fp] = 0;
    m.ctrl->newfeat_count[m.fp] = 0;
  }
  PROF_END(kProfGather);
}

// ================================================================================================
// Host-resident feature frames (nvbx_integrate_frame_host).  The reference only accepts CUDA tensors, so a
// caller with a host frame pays `.cuda()` on all H*W*C halves (384 MiB at 512^2 x 768) although the frame's
// work items read ~20 % of the pixels.  With a pinned (device-mapped) host frame the pixels cross PCIe
// sparsely instead: k_pixel_mark sets one bit per pixel any work item reads, k_pixel_fetch copies exactly
// those 2C-byte rows from the mapped host pointer to their natural place in the device frame buffer (every
// distinct pixel crosses the bus once), and k_feature_gather then runs unchanged on the device buffer.  The
// pixels that are not fetched keep stale bytes that no work item reads.
// ================================================================================================
__global__ void __launch_bounds__(256) k_pixel_mark(MapDev m, const FeatItem* __restrict__ items,
                                                    unsigned* __restrict__ bitmap, int cols) {
  pdl_prologue();
  const int n = m.ctrl->item_count[m.fp];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int pix = __ldg(&items[i].pix);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int p0 = pix + r * cols, p1 = p0 + 1;
      const unsigned b0 = 1u << (p0 & 31), b1 = 1u << (p1 & 31);
      if ((p0 >> 5) == (p1 >> 5)) {
        if ((bitmap[p0 >> 5] & (b0 | b1)) != (b0 | b1)) atomicOr(&bitmap[p0 >> 5], b0 | b1);
      } else {
        if (!(bitmap[p0 >> 5] & b0)) atomicOr(&bitmap[p0 >> 5], b0);
        if (!(bitmap[p1 >> 5] & b1)) atomicOr(&bitmap[p1 >> 5], b1);
      }
    }
  }
}

// warp = one 32-pixel bitmap word; lanes stride over the pixel row in 16-byte vectors.  NV > 0: the row is
// exactly NV * 32 vectors (C = 256 * NV) and all NV loads of a pixel are in flight together.
template <int NV>
__global__ void __launch_bounds__(256) k_pixel_fetch(MapDev m, unsigned* __restrict__ bitmap, int n_words,
                                                     const uint4* __restrict__ host_img, uint4* __restrict__ dev_img,
                                                     int nvec) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  unsigned fetched = 0;
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_words; w += warps_total) {
    unsigned bits = bitmap[w];
    if (!bits) continue;
    __syncwarp();
    if (lane == 0) bitmap[w] = 0u;  // clean for the next frame
    fetched += __popc(bits);
    while (bits) {
      const int b0 = __ffs(bits) - 1;
      bits &= bits - 1;
      const size_t o0 = (size_t)(w * 32 + b0) * nvec;
      if (NV > 0) {
        // two pixels (2 * NV independent 512-byte warp loads) in flight per iteration
        const bool two = bits != 0;
        const int b1 = two ? __ffs(bits) - 1 : b0;
        bits &= bits - 1;
        const size_t o1 = (size_t)(w * 32 + b1) * nvec;
        uint4 r0[NV > 0 ? NV : 1], r1[NV > 0 ? NV : 1];
#pragma unroll
        for (int k = 0; k < NV; ++k) r0[k] = __ldcs(host_img + o0 + k * 32 + lane);
        if (two) {
#pragma unroll
          for (int k = 0; k < NV; ++k) r1[k] = __ldcs(host_img + o1 + k * 32 + lane);
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) dev_img[o0 + k * 32 + lane] = r0[k];
        if (two) {
#pragma unroll
          for (int k = 0; k < NV; ++k) dev_img[o1 + k * 32 + lane] = r1[k];
        }
      } else {
        for (int v = lane; v < nvec; v += 32) dev_img[o0 + v] = __ldcs(host_img + o0 + v);
      }
    }
  }
  if (lane == 0 && fetched) count_add(m, kCntHostPixelsFetched, (unsigned long long)fetched);
}

// ================================================================================================
// N1. Colour integration: integrateBlocksKernel<UpdateAppearanceVoxelFunctor<ColorVoxel>>
// (projective_integrator_impl.cuh:156-214, projective_appearance_integrator.cu:277-353).  Same geometry
// as the feature path (projection, bilinear synthetic depth, band test, image bounds, mask); the payload
// is the reference's 8-byte ColorVoxel (r, g, b, pad | float weight), moved as one uint2 per voxel.
// interpolatePixels(Color) (interpolation_2d_impl.h:50-60) is the fp32 bilinear formula per channel followed
// by std::round; the blend weights arrive already rounded through binary16 (`__float2half(weight)` at
// :301-302); the first-observation test looks at the weight rounded to binary16 (:328).
// One CTA of 512 threads per band block, thread = voxel; 12 bytes of image per updated voxel -- negligible
// next to the feature gather, so no work-item indirection.
// ================================================================================================
struct ColorFrame {
  const uint8_t* img;
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
  float w1, w2;  // float(half((1-alpha)/total)), float(half(alpha/total))
};

constexpr unsigned kGrayVoxel = 0x007f7f7fu;  // Color::Gray() + zeroed padding byte (blox.cu:32-54)

__global__ void __launch_bounds__(512) k_color_update(MapDev m, const int* __restrict__ band_slots, ColorFrame f,
                                                      int parity) {
  pdl_prologue();
  const int n = m.ctrl->cband_count[parity];
  const int t = threadIdx.x;
  const int vx = t >> 6, vy = (t >> 3) & 7, vz = t & 7;
  unsigned updated = 0;
  for (int bi = blockIdx.x; bi < n; bi += gridDim.x) {
    const int raw = band_slots[bi];
    const bool is_new = (raw & kNewFlag) != 0;
    const int slot = raw & kSlotMask;
    const int3 b = m.blk_index[slot];
    uint2* vox = color_block(m, slot) + t;
    if (t == 0) m.blk_dirty[slot] = kDirtyAll;  // mapper.cpp:448
    bool write = is_new;
    uint2 out = make_uint2(kGrayVoxel, 0u);
    do {
      float u, v, vd;
      if (!project_voxel(f.cam, f.T_C_L, m.block_size, f.max_depth, b, vx, vy, vz, &u, &v, &vd)) break;
      const float ud = u / (float)f.sub, vdp = v / (float)f.sub;
      const float uc = ud - 0.5f, vc = vdp - 0.5f;
      const int lx = (int)floorf(uc), ly = (int)floorf(vc);
      if (lx < 0 || ly < 0 || (lx + 1) > (f.scols - 1) || (ly + 1) > (f.srows - 1)) break;
      const float* sp = f.synth + (size_t)ly * f.scols + lx;
      const float surface = dev_interp_float(uc - (float)lx, vc - (float)ly, sp[0], sp[f.scols], sp[1], sp[f.scols + 1]);
      if (fabsf(surface - vd) > f.trunc) break;
      const float fu = u - 0.5f, fv = v - 0.5f;
      const int px = (int)floorf(fu), py = (int)floorf(fv);
      if (px < 0 || py < 0 || (px + 1) > (f.cols - 1) || (py + 1) > (f.rows - 1)) break;
      if (f.mask != nullptr && !__ldg(f.mask + (size_t)((int)v) * f.cols + (int)u)) break;
      const uint2 cur = is_new ? make_uint2(kGrayVoxel, 0u) : *vox;
      const float w_cur = __uint_as_float(cur.y);
      const bool first = __half2float(__float2half_rn(w_cur)) == 0.0f;
      const float ox = fu - (float)px, oy = fv - (float)py;
      const uint8_t* p00 = f.img + ((size_t)py * f.cols + px) * 3;
      const uint8_t* p01 = p00 + (size_t)f.cols * 3;
      unsigned rgb = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float m00 = (float)__ldg(p00 + c), m10 = (float)__ldg(p00 + 3 + c);
        const float m01 = (float)__ldg(p01 + c), m11 = (float)__ldg(p01 + 3 + c);
        unsigned ch = (unsigned)(uint8_t)roundf(dev_interp_float(ox, oy, m00, m01, m10, m11));
        if (!first) {
          const float old = (float)((cur.x >> (8 * c)) & 0xffu);
          ch = (unsigned)(uint8_t)roundf(fmaf(f.w1, old, f.w2 * (float)ch));   // FMUL + FFMA in the reference
        }
        rgb |= ch << (8 * c);
      }
      out = make_uint2(rgb, __float_as_uint(fminf(f.alpha + w_cur, f.max_weight)));
      write = true;
      updated = 1;
    } while (false);
    if (write) *vox = out;
  }
  for (int o = 16; o; o >>= 1) updated += __shfl_xor_sync(0xffffffffu, updated, o);
  if (lane_id() == 0 && updated) count_add(m, kCntColorVoxelsUpdated, updated);
  if (blockIdx.x == 0 && t == 0) {
    count_add(m, kCntColorBandBlocks, (unsigned long long)n);
    count_add(m, kCntColorFrames, 1);
    m.ctrl->last_cband_count = n;
    m.ctrl->cband_count[parity ^ 1] = 0;  // ready for the next colour frame's band_select_tile
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
  pdl_prologue();
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
    // a decay that may change distances, or push weights to the sphere tracer's validity threshold (1e-4), voids
    // what kBlockFreeBit asserts; the default decay (floor 1e-3, distances untouched) does not
    if (threadIdx.x == 0 && (dp.set_free || dp.threshold <= 1e-4f)) m.blk_layers[slot] = layers & (uint8_t)~kBlockFreeBit;
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
        if (layers & kLayerColorBit) atomicAdd(&m.ctrl->n_color, -1);
        m.blk_mesh[slot] = make_int4(0, 0, 0, 0);
        m.blk_cmesh[slot] = make_int4(0, 0, 0, 0);
        m.blk_dirty[slot] = 0;
        unindex_slot(m, slot);
        push_id(&m.ctrl->slot_free_top, m.slot_free, slot);
        count_add(m, kCntBlocksDeallocated, 1);
      } else {
        m.blk_dirty[slot] = kDirtyAll;  // decayTsdf marks every TSDF block "to update" (mapper.cpp:469-471)
      }
    }
  }
}

// Overflow-hash rebuild after releases of hash-resident blocks (three tiny launches, all no-ops unless
// ctrl->rebuild is set; blocks inside the workspace grid never set it).
__global__ void __launch_bounds__(256) k_hash_clear(MapDev m, int force) {
  pdl_prologue();
  if (!force && !m.ctrl->rebuild) return;
  const unsigned n = m.hash_mask + 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m.keys[i] = kEmptyKey;
}
__global__ void __launch_bounds__(256) k_hash_reinsert(MapDev m, int force) {
  pdl_prologue();
  if (!force && !m.ctrl->rebuild) return;
  const int n = m.ctrl->slot_high;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (m.blk_layers[s]) {
      const int3 b = m.blk_index[s];
      if (ws_cell(m, b.x, b.y, b.z) < 0) hash_insert(m, b.x, b.y, b.z, s);
    }
  }
}
__global__ void k_hash_rebuild_done(MapDev m) {
  pdl_prologue(); m.ctrl->rebuild = 0; }

// Every TSDF block "to update" for every mesh layer: update*Mesh(UpdateFullLayer::kYes), as Mapper::loadMap does
// after swapping the layer cake (mapper.cpp:884-900).
__global__ void __launch_bounds__(256) k_mark_all_dirty(MapDev m) {
  pdl_prologue();
  const int n = m.ctrl->slot_high;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    if (m.blk_layers[s] & kLayerTsdfBit) m.blk_dirty[s] = kDirtyAll;
}

// Mapper::clear (py_mapper.cu:286-306): drop every block of every layer.
__global__ void __launch_bounds__(256) k_clear_all(MapDev m) {
  pdl_prologue();
  const unsigned n = m.hash_mask + 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m.keys[i] = kEmptyKey;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.ws_cells; i += gridDim.x * blockDim.x) m.ws_slot[i] = -1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    Ctrl* c = m.ctrl;
    c->n_hash = 0;
    for (int r = 0; r < kFrameRing; ++r) c->band_count[r] = c->newfeat_count[r] = c->item_count[r] = 0;
    c->slot_free_top = 0;
    c->slot_high = 0;
    c->feat_free_top = 0;
    c->feat_high = 0;
    c->n_tsdf = 0;
    c->n_feat = 0;
    c->n_color = 0;
    c->cband_count[0] = 0;
    c->cband_count[1] = 0;
    c->rebuild = 0;
    c->mesh_total_v = 0;
    c->mesh_total_t = 0;
  }
}

// ================================================================================================
// a13. Layer views and point queries.
// ================================================================================================
// A TSDF block view is about to be handed to the caller, who may write through it: no block is known to be free
// space any more (kBlockFreeBit).
__global__ void __launch_bounds__(256) k_clear_free_bits(MapDev m) {
  pdl_prologue();
  const int n = m.ctrl->slot_high;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const uint8_t l = m.blk_layers[s];
    if (l & kBlockFreeBit) m.blk_layers[s] = l & (uint8_t)~kBlockFreeBit;
  }
}

// getAllBlocks / getAllBlockIndices (py_layer.cpp:177-198) in one pass: slot id, block index AND payload pointer of
// every block of a layer (the host orders the result by slot id).
__global__ void __launch_bounds__(256) k_collect_blocks(MapDev m, int layer, int3* out_idx, unsigned long long* out_ptr,
                                                        int* out_slot, int capacity) {
  pdl_prologue();
  const int n = m.ctrl->slot_high;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    unsigned long long p = 0ull;
    if (layer == 0) {
      if (m.blk_layers[s] & kLayerTsdfBit) p = (unsigned long long)tsdf_block(m, s);
    } else if (layer == 2) {
      if (m.blk_layers[s] & kLayerColorBit) p = (unsigned long long)color_block(m, s);
    } else {
      if ((m.blk_layers[s] & kLayerFeatBit) && m.blk_feat[s] >= 0) p = (unsigned long long)feat_block(m, m.blk_feat[s]);
    }
    if (p) {
      const int pos = atomicAdd(&m.ctrl->list_count, 1);
      if (pos < capacity) {
        out_idx[pos] = m.blk_index[s];
        out_ptr[pos] = p;
        out_slot[pos] = s;
      }
    }
  }
}

// single-thread helpers for allocate_block_at_index / get_block_at_index
__global__ void k_allocate_one(MapDev m, int x, int y, int z, int layer, int* newfeat_slot_out) {
  pdl_prologue();
  bool is_new;
  const int slot = acquire_slot(m, x, y, z, &is_new);
  *newfeat_slot_out = -1;
  if (slot < 0) return;
  // invariant relied on by k_sphere_trace: a slot's TSDF payload is zero unless it holds a TSDF layer
  if (is_new || (layer == 0 && !(m.blk_layers[slot] & kLayerTsdfBit))) {
    float2* p = tsdf_block(m, slot);
    for (int i = 0; i < kVoxelsPerBlock; ++i) p[i] = make_float2(0.f, 0.f);
  }
  if (layer == 0) {
    ensure_tsdf_layer(m, slot);
  } else if (layer == 2) {
    if (!(m.blk_layers[slot] & kLayerColorBit)) {
      m.blk_layers[slot] |= kLayerColorBit;
      atomicAdd(&m.ctrl->n_color, 1);
      uint2* c = color_block(m, slot);
      for (int i = 0; i < kVoxelsPerBlock; ++i) c[i] = make_uint2(kGrayVoxel, 0u);
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
  pdl_prologue();
  if (*fslot < 0) return;
  uint4* p = reinterpret_cast<uint4*>(feat_block(m, *fslot));
  const int vec_per_block = (kVoxelsPerBlock * m.row) / 8;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < vec_per_block; k += gridDim.x * blockDim.x)
    p[k] = make_uint4(0, 0, 0, 0);
}
__global__ void k_find_one(MapDev m, int x, int y, int z, int layer, unsigned long long* ptr_out) {
  pdl_prologue();
  const int slot = find_slot(m, x, y, z);
  *ptr_out = 0ull;
  if (slot < 0) return;
  if (layer == 0) {
    if (m.blk_layers[slot] & kLayerTsdfBit) *ptr_out = (unsigned long long)tsdf_block(m, slot);
  } else if (layer == 2) {
    if (m.blk_layers[slot] & kLayerColorBit) *ptr_out = (unsigned long long)color_block(m, slot);
  } else {
    if (m.blk_feat[slot] >= 0) *ptr_out = (unsigned long long)feat_block(m, m.blk_feat[slot]);
  }
}

// queryTSDFKernel (NT/cpp/src/sdf_query.cu:240-270): one thread per query, rows of misses untouched.
__global__ void __launch_bounds__(128) k_query_tsdf(MapDev m, const float* __restrict__ xyz, long long n,
                                                    float2* __restrict__ out) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 p;
  p.x = xyz[3 * i];
  p.y = xyz[3 * i + 1];
  p.z = xyz[3 * i + 2];
  I3 b, v;
  dev_block_and_voxel_from_position(m.block_size, m.voxel_size_inv, p, &b, &v);
  const int slot = find_slot(m, b.x, b.y, b.z);
  if (slot < 0 || !(m.blk_layers[slot] & kLayerTsdfBit)) return;
  out[i] = tsdf_block(m, slot)[(v.x * 8 + v.y) * 8 + v.z];
}
// queryFeatureKernel (sdf_query.cu:206-238): one WARP per query; the output row has C+1 halves
// (2-byte aligned), so lanes write single halves, coalesced.
__global__ void __launch_bounds__(128) k_query_features(MapDev m, const float* __restrict__ xyz, long long n,
                                                        __half* __restrict__ out) {
  pdl_prologue();
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n) return;
  V3 p;
  p.x = xyz[3 * q];
  p.y = xyz[3 * q + 1];
  p.z = xyz[3 * q + 2];
  I3 b, v;
  dev_block_and_voxel_from_position(m.block_size, m.voxel_size_inv, p, &b, &v);
  const int slot = find_slot(m, b.x, b.y, b.z);
  if (slot < 0 || m.blk_feat[slot] < 0) return;
  const __half* row = feat_block(m, m.blk_feat[slot]) + (size_t)((v.x * 8 + v.y) * 8 + v.z) * m.row;
  __half* o = out + q * (long long)(m.C + 1);
  for (int c = lane_id(); c <= m.C; c += 32) o[c] = row[c];
}

}  // namespace nvbx
