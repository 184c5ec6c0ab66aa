// nvbx_map.cuh -- device-resident sparse map store.
//
// Replaces the reference's CPU std::unordered_map<Index3D, unified_ptr<Block>> master + stdgpu mirror +
// one cudaMallocAsync per block (NB/include/nvblox/map/internal/impl/layer_impl.h:106-305,
// block_memory_pool_impl.h:22-73, gpu_hash/internal/cuda/impl/gpu_layer_view_impl.cuh:36-174) with ONE
// structure that lives in HBM and is only ever touched by kernels:
//
//   * a two-level block index.  Level 1 is a DIRECT-MAPPED grid over the workspace bounding box (mindmap
//     always sets one: nvblox_mapping_helpers.py:53-61; <= 405 / 3 740 cells for its tasks): block index
//     -> slot in ONE load, no probing, release = one store.  Level 2 is an open-addressing hash (64-bit
//     packed key -> 32-bit slot, linear probing, load <= 0.25, no tombstones: rebuilt from the slot table
//     after a decay that freed hash-resident blocks) for indices outside the box and for unbounded maps;
//   * a slot table (struct-of-arrays): block index, layer bits, feature-slot id, mesh-dirty flag and the
//     block's extent in the current mesh arena;
//   * slab arenas for voxel payloads: TSDF float2[512] per slot (slot id == payload id), feature
//     fp16[512][C+8] per feature slot; slabs are appended, never moved, so block views stay valid;
//   * free-list stacks + high-water marks for both payload kinds, all driven by atomics on the device.
//
// Nothing here needs the host in steady state: allocation, lookup, release all happen inside kernels.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "nvbx_math.cuh"

namespace nvbx {

constexpr int kVoxelsPerBlock = 512;
constexpr int kFrameRing = 4;  // feature frames whose per-frame lists may be alive at once (frame pipelining)
constexpr int kTsdfSlabShift = 10;  // 1024 blocks  (4 MiB) per TSDF slab
constexpr int kFeatSlabShift = 4;   // 16 blocks (12.1 MiB at C=768) per feature slab
constexpr int kMaxSlabs = 1 << 15;
constexpr unsigned long long kEmptyKey = ~0ull;

constexpr uint8_t kLayerTsdfBit = 1;
constexpr uint8_t kLayerFeatBit = 2;
constexpr uint8_t kLayerColorBit = 4;
// Not a layer: "every one of this block's 512 TSDF voxels holds distance == +truncation distance with a weight above
// the sphere tracer's validity threshold", i.e. the block is observed free space.  Set / cleared by k_tsdf_update for
// the blocks it updates, cleared by whatever else may change TSDF voxels; lets the sphere tracer step through such a
// block without reading it (the value it would read is known exactly).
constexpr uint8_t kBlockFreeBit = 8;
// blk_dirty bits: one "to update" set per mesh layer (BlocksToUpdateTracker, blocks_to_update_tracker.cpp:32-60:
// every addBlocksToUpdate feeds all consumer sets, every consumer clears only its own)
constexpr uint8_t kDirtyFeatMesh = 1;
constexpr uint8_t kDirtyColorMesh = 2;
constexpr uint8_t kDirtyAll = 3;
// slot lists carry "freshly allocated, payload not initialised yet" in bit 30 (slot ids are < 2^30)
constexpr int kNewFlag = 0x40000000;
constexpr int kSlotMask = 0x3fffffff;

// counter slots (unsigned long long each) -- mirror nvbx_counters
enum CounterId {
  kCntDepthFrames = 0,
  kCntFeatureFrames,
  kCntTsdfBlocksInView,
  kCntTsdfVoxelsUpdated,
  kCntTsdfBlocksAllocated,
  kCntFeatCandidateBlocks,
  kCntFeatBandBlocks,
  kCntFeatVoxelsUpdated,
  kCntFeatBlocksAllocated,
  kCntBlocksDeallocated,
  kCntMeshBlocksRemeshed,
  kCntMeshVertices,
  kCntColorFrames,
  kCntColorBandBlocks,
  kCntColorVoxelsUpdated,
  kCntColorBlocksAllocated,
  kCntProfile0 = 16,  // NVBX_PROFILE_COUNTERS builds: 16..19
  kCntHostPixelsFetched = 20,  // nvbx_integrate_frame_host: feature pixels read from the mapped host frame
  kCntNum = 24
};

// Device-side control block (one per map).
struct Ctrl {
  int slot_free_top;   // number of entries on the slot free stack
  int slot_high;       // slots [0, slot_high) have been handed out at least once
  int feat_free_top;
  int feat_high;
  int n_tsdf;          // live blocks per layer
  int n_feat;
  int overflow;        // set when a pool ran dry (host sizing bug -- reported as an error)
  int rebuild;         // hash must be rebuilt (blocks were released)
  int view_count;      // length of the current TSDF view list
  int cand_count;      // length of the feature candidate list
  // Per-feature-frame lists live in the ring slot MapDev::fp = (feature frame number) mod kFrameRing.  Frame i
  // appends its band list to band_count[fp] (band selection), its geometry kernel reads it and fills item_count[fp]
  // items, its gather consumes them and clears the band counters; item_count[fp] is cleared by the band-selection
  // kernel of the NEXT frame that uses the slot (frame i + kFrameRing).  Before the host enqueues that frame it makes
  // sure -- with an event query, in practice always already satisfied -- that the gather of frame i is complete, so
  // the kernels of frame i that run on the map's gather stream (nvbx_set_pipelining) share no list with the kernels of
  // the following frames, and no wait is ever inserted between two kernels of the caller's stream.
  int band_count[kFrameRing];     // length of the feature band list (k_trace_and_band -> k_feature_geometry)
  int newfeat_count[kFrameRing];  // feature blocks allocated this frame (to be zero-filled by k_feature_geometry)
  int list_count;      // generic compaction counter (block index export)
  int mesh_total_v;    // totals of the mesh being built
  int mesh_total_t;
  int item_count[kFrameRing];   // length of the feature work-item list of the current chunk (geometry -> gather)
  int last_band_count; // band_count of the last completed feature frame (debug / parity hook)
  int n_hash;          // blocks resident in the overflow hash (0: every block is in the workspace grid)
  int n_color;         // live colour blocks
  int cband_count[2];  // colour band list length, double-buffered by frame parity (the consumer of one frame
                       // clears the other half, so no kernel both reads and resets the same counter)
  int last_cband_count;
  int gather_ticket[kFrameRing];  // k_feature_gather_dyn's work ticket; zeroed by k_feature_geometry
  unsigned bitmap_clean_seq;  // small-view bitmaps (triple buffer): thirds of depth frames <= this number are wiped
  unsigned long long counters[kCntNum];
};

struct MapDev {
  // level-1 index: direct-mapped workspace grid (ws_slot == nullptr / ws_sx == 0: disabled)
  int* ws_slot;
  I3 ws_mn;
  int ws_sx, ws_sy, ws_sz;
  int ws_cells;
  // level-2 index: overflow hash
  unsigned long long* keys;
  int* vals;
  unsigned int hash_mask;
  // slot table
  int3* blk_index;
  uint8_t* blk_layers;
  int* blk_feat;
  uint8_t* blk_dirty;
  int4* blk_mesh;  // (vertex offset, vertex count, triangle-index offset, triangle-index count)
  int4* blk_cmesh; // the same for the colour mesh layer
  int slot_capacity;
  int feat_capacity;
  // payload arenas
  float2* const* tsdf_slabs;
  __half* const* feat_slabs;
  uint2* const* color_slabs;  // ColorVoxel records (r, g, b, pad | float weight), slot-indexed like the TSDF slabs;
                              // allocated on the first colour frame
  // free lists
  int* slot_free;
  int* feat_free;
  Ctrl* ctrl;
  // geometry / layout
  float block_size;
  float voxel_size;
  float voxel_size_inv;
  int C;    // feature channels
  int row;  // halves per feature voxel row (C + 8)
  int seq;  // feature frames enqueued so far (host count; labels the timeline stamps of profile builds)
  int fp;   // ring slot of the current feature frame (frame number mod kFrameRing): which copy of the per-frame Ctrl
            // counters / lists it uses
};

__device__ __forceinline__ float2* tsdf_block(const MapDev& m, int slot) {
  return m.tsdf_slabs[slot >> kTsdfSlabShift] + (size_t)(slot & ((1 << kTsdfSlabShift) - 1)) * kVoxelsPerBlock;
}
__device__ __forceinline__ uint2* color_block(const MapDev& m, int slot) {
  return m.color_slabs[slot >> kTsdfSlabShift] + (size_t)(slot & ((1 << kTsdfSlabShift) - 1)) * kVoxelsPerBlock;
}
__device__ __forceinline__ __half* feat_block(const MapDev& m, int fslot) {
  return m.feat_slabs[fslot >> kFeatSlabShift] +
         (size_t)(fslot & ((1 << kFeatSlabShift) - 1)) * (size_t)kVoxelsPerBlock * (size_t)m.row;
}

__host__ __device__ __forceinline__ bool key_in_range(int x, int y, int z) {
  const int lim = 1 << 20;
  return x >= -lim && x < lim && y >= -lim && y < lim && z >= -lim && z < lim;
}
__host__ __device__ __forceinline__ unsigned long long pack_key(int x, int y, int z) {
  const unsigned long long bx = (unsigned long long)(unsigned)(x + (1 << 20)) & 0x1fffffull;
  const unsigned long long by = (unsigned long long)(unsigned)(y + (1 << 20)) & 0x1fffffull;
  const unsigned long long bz = (unsigned long long)(unsigned)(z + (1 << 20)) & 0x1fffffull;
  return (bx << 42) | (by << 21) | bz;
}
__host__ __device__ __forceinline__ unsigned int mix_key(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned int)k;
}

__device__ __forceinline__ int hash_find(const MapDev& m, int x, int y, int z) {
  if (!key_in_range(x, y, z)) return -1;
  const unsigned long long key = pack_key(x, y, z);
  unsigned int h = mix_key(key) & m.hash_mask;
  // bounded probe: a table without an empty key (a host sizing bug) must not hang the GPU
  for (unsigned int n = 0; n <= m.hash_mask; ++n) {
    const unsigned long long k = m.keys[h];
    if (k == key) return m.vals[h];
    if (k == kEmptyKey) return -1;
    h = (h + 1) & m.hash_mask;
  }
  return -1;
}
// Insert a key known to be absent and inserted by exactly one thread.
__device__ __forceinline__ void hash_insert(const MapDev& m, int x, int y, int z, int slot) {
  const unsigned long long key = pack_key(x, y, z);
  unsigned int h = mix_key(key) & m.hash_mask;
  for (unsigned int n = 0; n <= m.hash_mask; ++n) {
    const unsigned long long prev = atomicCAS(&m.keys[h], kEmptyKey, key);
    if (prev == kEmptyKey) {
      m.vals[h] = slot;
      return;
    }
    h = (h + 1) & m.hash_mask;
  }
  atomicExch(&m.ctrl->overflow, 1);  // table full: reported as an error by the next read_ctrl (never hangs)
}

// Cell of the workspace grid holding block (x, y, z), or -1 when the index lies outside the box (or the
// grid is disabled: ws_sx == 0 makes the unsigned comparison fail).
__device__ __forceinline__ int ws_cell(const MapDev& m, int x, int y, int z) {
  const unsigned lx = (unsigned)(x - m.ws_mn.x), ly = (unsigned)(y - m.ws_mn.y), lz = (unsigned)(z - m.ws_mn.z);
  if (lx < (unsigned)m.ws_sx && ly < (unsigned)m.ws_sy && lz < (unsigned)m.ws_sz)
    return (int)(lx + (unsigned)m.ws_sx * (ly + (unsigned)m.ws_sy * lz));
  return -1;
}
// Block index -> slot (or -1): one load inside the workspace grid, a hash probe outside.
__device__ __forceinline__ int find_slot(const MapDev& m, int x, int y, int z) {
  const int cell = ws_cell(m, x, y, z);
  if (cell >= 0) return m.ws_slot[cell];
  return hash_find(m, x, y, z);
}

// Pop a slot id: recycled ids first, then fresh ones.  Returns -1 (and raises ctrl->overflow) when the
// arena is exhausted -- the host sizes arenas before every launch so this is an internal error.
__device__ __forceinline__ int pop_id(int* free_top, int* high, const int* free_stack, int capacity, int* overflow) {
  const int top = atomicSub(free_top, 1);
  if (top > 0) return free_stack[top - 1];
  atomicAdd(free_top, 1);
  const int id = atomicAdd(high, 1);
  if (id >= capacity) {
    atomicAdd(high, -1);
    atomicExch(overflow, 1);
    return -1;
  }
  return id;
}
__device__ __forceinline__ void push_id(int* free_top, int* free_stack, int id) {
  const int top = atomicAdd(free_top, 1);
  free_stack[top] = id;
}

// Find-or-create the slot of a block index and make sure the TSDF layer bit is set.  The caller
// guarantees that no other thread handles the same index concurrently.  *created_tsdf tells the caller
// to zero the TSDF payload.
__device__ __forceinline__ int acquire_slot(const MapDev& m, int x, int y, int z, bool* is_new_slot) {
  *is_new_slot = false;
  const int cell = ws_cell(m, x, y, z);
  int slot;
  if (cell >= 0) {
    slot = m.ws_slot[cell];
  } else {
    if (!key_in_range(x, y, z)) return -1;  // |index| >= 2^20 blocks: outside the packed-key range
    slot = hash_find(m, x, y, z);
  }
  if (slot >= 0) return slot;
  slot = pop_id(&m.ctrl->slot_free_top, &m.ctrl->slot_high, m.slot_free, m.slot_capacity, &m.ctrl->overflow);
  if (slot < 0) return -1;
  m.blk_index[slot] = make_int3(x, y, z);
  m.blk_layers[slot] = 0;
  m.blk_feat[slot] = -1;
  m.blk_dirty[slot] = 0;
  m.blk_mesh[slot] = make_int4(0, 0, 0, 0);
  m.blk_cmesh[slot] = make_int4(0, 0, 0, 0);
  if (cell >= 0) {
    m.ws_slot[cell] = slot;
  } else {
    hash_insert(m, x, y, z, slot);
    atomicAdd(&m.ctrl->n_hash, 1);
  }
  *is_new_slot = true;
  return slot;
}

// Drop a block index from the index (its slot id is pushed back by the caller).  Grid-resident blocks
// cost one store; hash-resident ones raise ctrl->rebuild (no tombstones).
__device__ __forceinline__ void unindex_slot(const MapDev& m, int slot) {
  const int3 b = m.blk_index[slot];
  const int cell = ws_cell(m, b.x, b.y, b.z);
  if (cell >= 0) {
    m.ws_slot[cell] = -1;
  } else {
    atomicAdd(&m.ctrl->n_hash, -1);
    m.ctrl->rebuild = 1;
  }
}

}  // namespace nvbx
