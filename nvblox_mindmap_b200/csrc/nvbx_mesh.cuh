// nvbx_mesh.cuh -- surface extraction.
//
// Reference pipeline per update (mesh_integrator.cu:64-103,491-688, mesh_integrator_appearance.cu:
// 290-340, mesh_serializer_gpu.cu:27-70): meshable test kernel -> D2H -> CPU filter -> 8 CPU hash probes
// per block -> table-index kernel writing 136 B/voxel of scratch -> D2H sizes -> per-block host vector
// resize -> vertex kernel -> cub weld kernel -> D2H sizes -> resize -> 3 cudaMallocs -> closest-voxel
// appearance kernel -> 3 serialisation kernels into PINNED HOST memory.
//
// Here: one CTA per voxel block does meshable test + marching cubes + weld entirely in shared memory,
// twice (count pass, emit pass) around a device-side exclusive scan, and the emit pass writes the
// final serialised layout ([N,3] f32 vertices, [N,C] f16 features, [M,3] i32 global triangle ids)
// straight into a device arena.  Blocks whose "to update" flag is clear keep their previous mesh (it is
// copied from the previous arena), which is exactly the reference's incremental semantics
// (BlocksToUpdateTracker, mapper.cpp:580-614).
//
// Determinism: the reference's vertex order inside a block depends on atomicAdd arrival order
// (marching_cubes_impl.cuh:31-33) and its weld keeps whichever duplicate the radix sort saw first.  We
// fix the canonical order the CPU oracle uses: voxels in memory order, triangles in table order, weld
// keeps the first duplicate and orders survivors by ascending key.
#pragma once
#include "nvbx_kernels.cuh"

namespace nvbx {

constexpr int kWeldLimit = 128 * 20;  // weldVerticesCubKernel<128,20>: blocks with >= 2560 vertices are not welded
constexpr int kSortCapacity = 4096;

struct MeshParams {
  float min_weight;
  float cutoff;  // cutoff_distance_vox * voxel_size
  int weld;
};

struct MeshArena {
  float* verts;     // [cap_v * 3]
  __half* feats;    // [cap_v * C]   (feature mesh)
  int* tris;        // [cap_t]
  uint8_t* colors;  // [cap_v * 3]   (colour mesh)
};

// The reference keeps one mesh layer per appearance type (MeshBlockLayer<FeatureArray>, MeshBlockLayer<Color>),
// each with its own "to update" set; KIND selects which one a launch works on.
enum MeshKind { kMeshFeature = 0, kMeshColor = 1 };
template <int KIND>
__device__ __forceinline__ int4* mesh_extents(const MapDev& m) {
  return KIND == kMeshFeature ? m.blk_mesh : m.blk_cmesh;
}
template <int KIND>
__device__ __forceinline__ uint8_t mesh_dirty_bit() {
  return KIND == kMeshFeature ? kDirtyFeatMesh : kDirtyColorMesh;
}

// shared-memory carve-up of the mesh CTA
struct MeshSmem {
  float2 tsdf[729];  // 9x9x9 corner samples (distance, weight); weight < 0: neighbour block missing
  int nb_slot[8];
  int warp_sum[17];
  int flags[4];
  float vx[kWeldLimit], vy[kWeldLimit], vz[kWeldLimit];
  unsigned long long key[kSortCapacity];
  unsigned short idx[kSortCapacity];
  unsigned short tri[kWeldLimit];
  unsigned short head[kWeldLimit];
};

// exclusive scan of one int per thread across a 512-thread CTA; returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_exclusive_scan_512(int v, int* warp_sum /*[17]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_sum[warp + 1] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    warp_sum[0] = 0;
    for (int w = 1; w <= 16; ++w) {
      acc += warp_sum[w];
      warp_sum[w] = acc;
    }
  }
  __syncthreads();
  const int excl = warp_sum[warp] + inc - v;
  *total = warp_sum[16];
  __syncthreads();  // warp_sum may be reused immediately
  return excl;
}

__device__ __forceinline__ unsigned long long weld_key(float x, float y, float z) {  // Index3DHash(int(v*1000))
  const int ix = (int)(x * 1000), iy = (int)(y * 1000), iz = (int)(z * 1000);
  return (unsigned long long)(long long)ix + (unsigned long long)(long long)iy * 17191ull +
         (unsigned long long)(long long)iz * (17191ull * 17191ull);
}

__device__ __forceinline__ void interp_vertex(const float* a, const float* b, float s1, float s2, float* o) {
  const float diff = s1 - s2;
  if (fabsf(diff) >= 1e-4f) {
    const float t = s1 / diff;
    o[0] = fmaf(t, b[0] - a[0], a[0]);   // vertex1 + t * (vertex2 - vertex1): one FFMA per axis in the reference
    o[1] = fmaf(t, b[1] - a[1], a[1]);
    o[2] = fmaf(t, b[2] - a[2], a[2]);
  } else {
    o[0] = 0.5f * (a[0] + b[0]);
    o[1] = 0.5f * (a[1] + b[1]);
    o[2] = 0.5f * (a[2] + b[2]);
  }
}

// Closest-voxel feature of a vertex (updateAppearanceBlockByClosestVoxel, mesh_integrator_appearance.cu:
// 97-146): one warp copies the C-half row with 128-bit accesses; zeros when the block has no features
// (updateAppearanceBlocksConstant :148-159).
__device__ __forceinline__ void warp_paint_vertex(const MapDev& m, const __half* fblk, const int3 b, float vs,
                                                  float x, float y, float z, __half* dst) {
  const int nvec = m.C >> 3;
  uint4* d = reinterpret_cast<uint4*>(dst);
  const int lane = threadIdx.x & 31;
  if (fblk == nullptr) {
    for (int c = lane; c < nvec; c += 32) d[c] = make_uint4(0, 0, 0, 0);
    return;
  }
  // p_L_V - block_size * float(block index): mul.f32 + sub.f32 in the reference's PTX, one FFMA in its SASS
  const float bs = m.block_size;
  int ix = (int)(fmaf(-bs, (float)b.x, x) / vs), iy = (int)(fmaf(-bs, (float)b.y, y) / vs),
      iz = (int)(fmaf(-bs, (float)b.z, z) / vs);
  ix = max(min(ix, 7), 0);
  iy = max(min(iy, 7), 0);
  iz = max(min(iz, 7), 0);
  const uint4* s = reinterpret_cast<const uint4*>(fblk + (size_t)((ix * 8 + iy) * 8 + iz) * m.row);
  for (int c = lane; c < nvec; c += 32) d[c] = s[c];
}

// One CTA (512 threads) meshes one block.  EMIT=false: only sizes.  EMIT=true: writes the block's
// vertices / triangles / features at (voff, toff) of `out`.
template <bool EMIT, int KIND>
__device__ void mesh_one_block(const MapDev& m, const MeshParams& mp, MeshSmem& s, int slot, int voff, int toff,
                               const MeshArena& out, int* n_verts_out, int* n_tris_out) {
  const int t = threadIdx.x;
  const int3 b = m.blk_index[slot];
  const float vs = m.voxel_size;
  const float origin[3] = {m.block_size * (float)b.x, m.block_size * (float)b.y, m.block_size * (float)b.z};

  // neighbour slots (+x,+y,+z directions), directionFromNeighborIndex: j -> ((j>>2)&1, (j>>1)&1, j&1)
  if (t < 8) {
    int ns = (t == 0) ? slot : find_slot(m, b.x + ((t >> 2) & 1), b.y + ((t >> 1) & 1), b.z + (t & 1));
    if (ns >= 0 && !(m.blk_layers[ns] & kLayerTsdfBit)) ns = -1;
    s.nb_slot[t] = ns;
  }
  __syncthreads();
  // stage the 9x9x9 corner lattice
  for (int i = t; i < 729; i += 512) {
    const int cx = i / 81, cy = (i / 9) % 9, cz = i % 9;
    const int j = ((cx >> 3) << 2) | ((cy >> 3) << 1) | (cz >> 3);
    const int ns = s.nb_slot[j];
    float2 q = make_float2(0.f, -1.f);
    if (ns >= 0) q = tsdf_block(m, ns)[((cx & 7) * 8 + (cy & 7)) * 8 + (cz & 7)];
    s.tsdf[i] = q;
  }
  __syncthreads();

  const int vx = t >> 6, vy = (t >> 3) & 7, vz = t & 7;
  // isBlockMeshableKernel (mesh_integrator.cu:308-326) on the block's own voxels
  const float2 own = s.tsdf[vx * 81 + vy * 9 + vz];
  const int meshable = __syncthreads_or(fabsf(own.x) <= mp.cutoff && own.y >= mp.min_weight);
  if (!meshable) {
    *n_verts_out = 0;
    *n_tris_out = 0;
    return;
  }

  // marching cubes configuration of this voxel's cube
  float sdf[8];
  bool skip = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 q = s.tsdf[(vx + kMcCornerOffsets[i][0]) * 81 + (vy + kMcCornerOffsets[i][1]) * 9 +
                            (vz + kMcCornerOffsets[i][2])];
    // a missing neighbour block (weight -1) and an unobserved corner both skip the cube
    if (q.y < mp.min_weight) skip = true;
    sdf[i] = q.x;
  }
  int cfg = 0;
  if (!skip) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (sdf[i] < 0) cfg |= 1 << i;
  }
  const int nv = skip ? 0 : (int)kMcNumVerts[cfg];
  int total;
  const int off = block_exclusive_scan_512(nv, s.warp_sum, &total);
  if (total == 0) {
    *n_verts_out = 0;
    *n_tris_out = 0;
    return;
  }
  const bool weld = mp.weld && total < kWeldLimit;
  if (!EMIT && !weld) {
    *n_verts_out = total;
    *n_tris_out = total;
    return;
  }

  // ---- generate this voxel's vertices ------------------------------------------------------------
  if (nv) {
    const int8_t* row = kMcTriTable[cfg];
    for (int k = 0; k < nv; ++k) {
      // calculateVertices writes (col+2, col+1, col) per triangle
      const int e = row[(k / 3) * 3 + (2 - k % 3)];
      const int ca = kMcEdgePairs[e][0], cb = kMcEdgePairs[e][1];
      float pa[3], pb[3];
      {
        const int ax = vx + kMcCornerOffsets[ca][0], ay = vy + kMcCornerOffsets[ca][1],
                  az = vz + kMcCornerOffsets[ca][2];
        const int bx = vx + kMcCornerOffsets[cb][0], by = vy + kMcCornerOffsets[cb][1],
                  bz = vz + kMcCornerOffsets[cb][2];
        // block_position + voxel_size * ((corner&7) + 0.5 + 8*block_offset)   mesh_integrator.cu:421-424
        pa[0] = fmaf(vs, ((float)(ax & 7) + 0.5f) + (float)(8 * (ax >> 3)), origin[0]);
        pa[1] = fmaf(vs, ((float)(ay & 7) + 0.5f) + (float)(8 * (ay >> 3)), origin[1]);
        pa[2] = fmaf(vs, ((float)(az & 7) + 0.5f) + (float)(8 * (az >> 3)), origin[2]);
        pb[0] = fmaf(vs, ((float)(bx & 7) + 0.5f) + (float)(8 * (bx >> 3)), origin[0]);
        pb[1] = fmaf(vs, ((float)(by & 7) + 0.5f) + (float)(8 * (by >> 3)), origin[1]);
        pb[2] = fmaf(vs, ((float)(bz & 7) + 0.5f) + (float)(8 * (bz >> 3)), origin[2]);
      }
      // corner distances come back from the staged lattice (dynamic register indexing would spill)
      const float sa = s.tsdf[(vx + kMcCornerOffsets[ca][0]) * 81 + (vy + kMcCornerOffsets[ca][1]) * 9 +
                              (vz + kMcCornerOffsets[ca][2])].x;
      const float sb = s.tsdf[(vx + kMcCornerOffsets[cb][0]) * 81 + (vy + kMcCornerOffsets[cb][1]) * 9 +
                              (vz + kMcCornerOffsets[cb][2])].x;
      float p[3];
      interp_vertex(pa, pb, sa, sb, p);
      if (weld) {
        s.vx[off + k] = p[0];
        s.vy[off + k] = p[1];
        s.vz[off + k] = p[2];
      } else {  // EMIT, unwelded: straight to the arena
        float* o = out.verts + (size_t)(voff + off + k) * 3;
        o[0] = p[0];
        o[1] = p[1];
        o[2] = p[2];
        out.tris[toff + off + k] = voff + off + k;
      }
    }
  }
  __syncthreads();

  int n_unique = total;
  if (weld) {
    // ---- weld: sort (key, original id), keep the first of every key run -----------------------------
    int npad = 1;
    while (npad < total) npad <<= 1;
    for (int i = t; i < npad; i += 512) {
      if (i < total) {
        s.key[i] = weld_key(s.vx[i], s.vy[i], s.vz[i]);
        s.idx[i] = (unsigned short)i;
      } else {
        s.key[i] = ~0ull;
        s.idx[i] = 0xffff;
      }
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = t; i < npad; i += 512) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long ka = s.key[i], kb = s.key[ixj];
            const unsigned short ia = s.idx[i], ib = s.idx[ixj];
            const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
            const bool ascending = ((i & k) == 0);
            if (a_gt_b == ascending) {
              s.key[i] = kb;
              s.key[ixj] = ka;
              s.idx[i] = ib;
              s.idx[ixj] = ia;
            }
          }
        }
        __syncthreads();
      }
    }
    // head flags + scan: thread t owns sorted positions [t*chunk, (t+1)*chunk)
    const int chunk = (total + 511) / 512;
    const int p0 = t * chunk, p1 = min(p0 + chunk, total);
    int heads = 0;
    for (int p = p0; p < p1; ++p) heads += (p == 0 || s.key[p] != s.key[p - 1]) ? 1 : 0;
    const int hoff = block_exclusive_scan_512(heads, s.warp_sum, &n_unique);
    int id = hoff - 1;
    for (int p = p0; p < p1; ++p) {
      const bool is_head = (p == 0 || s.key[p] != s.key[p - 1]);
      if (is_head) {
        ++id;
        s.head[id] = s.idx[p];
      }
      s.tri[s.idx[p]] = (unsigned short)id;
    }
    __syncthreads();
    if (EMIT) {
      for (int i = t; i < n_unique; i += 512) {
        const int src = s.head[i];
        float* o = out.verts + (size_t)(voff + i) * 3;
        o[0] = s.vx[src];
        o[1] = s.vy[src];
        o[2] = s.vz[src];
      }
      for (int i = t; i < total; i += 512) out.tris[toff + i] = voff + (int)s.tri[i];
    }
  }

  if (EMIT && KIND == kMeshColor) {
    // ---- vertex colours: closest voxel's colour, Gray without a colour block (AppearanceGetter<ColorVoxel>,
    // mesh_integrator_appearance.cu:43-54), one thread per vertex
    const bool has_color = (m.blk_layers[slot] & kLayerColorBit) != 0;
    const uint2* cblk = has_color ? color_block(m, slot) : nullptr;
    const float vs_paint = m.block_size / 8;
    if (!weld) __syncthreads();
    for (int i = t; i < n_unique; i += 512) {
      float x, y, z;
      if (weld) {
        const int src = s.head[i];
        x = s.vx[src];
        y = s.vy[src];
        z = s.vz[src];
      } else {
        const float* o = out.verts + (size_t)(voff + i) * 3;
        x = o[0];
        y = o[1];
        z = o[2];
      }
      unsigned rgb = kGrayVoxel;
      if (cblk) {
        const float bs = m.block_size;
        int ix = (int)(fmaf(-bs, (float)b.x, x) / vs_paint), iy = (int)(fmaf(-bs, (float)b.y, y) / vs_paint),
            iz = (int)(fmaf(-bs, (float)b.z, z) / vs_paint);
        ix = max(min(ix, 7), 0);
        iy = max(min(iy, 7), 0);
        iz = max(min(iz, 7), 0);
        rgb = cblk[(ix * 8 + iy) * 8 + iz].x;
      }
      uint8_t* d = out.colors + (size_t)(voff + i) * 3;
      d[0] = (uint8_t)(rgb & 0xffu);
      d[1] = (uint8_t)((rgb >> 8) & 0xffu);
      d[2] = (uint8_t)((rgb >> 16) & 0xffu);
    }
  }
  if (EMIT && KIND == kMeshFeature) {
    // ---- vertex features: one warp per vertex ---------------------------------------------------------
    const int fs = m.blk_feat[slot];
    const __half* fblk = fs >= 0 ? feat_block(m, fs) : nullptr;
    const float vs_paint = m.block_size / 8;  // mesh_layer->block_size() / kVoxelsPerSide
    if (!weld) __syncthreads();               // unwelded vertices were written to the arena above
    const int warp = t >> 5;
    for (int i = warp; i < n_unique; i += 16) {
      float x, y, z;
      if (weld) {
        const int src = s.head[i];
        x = s.vx[src];
        y = s.vy[src];
        z = s.vz[src];
      } else {
        const float* o = out.verts + (size_t)(voff + i) * 3;
        x = o[0];
        y = o[1];
        z = o[2];
      }
      warp_paint_vertex(m, fblk, b, vs_paint, x, y, z, out.feats + (size_t)(voff + i) * m.C);
    }
  }
  *n_verts_out = n_unique;
  *n_tris_out = total;
}

// pass 1: per-slot sizes of the NEW mesh
template <int KIND>
__global__ void __launch_bounds__(512, 2) k_mesh_count(MapDev m, MeshParams mp, int* cnt_v, int* cnt_t) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MeshSmem& s = *reinterpret_cast<MeshSmem*>(smem_raw);
  const int n = m.ctrl->slot_high;
  unsigned remeshed = 0;
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    int nv = 0, nt = 0;
    const bool has_tsdf = (m.blk_layers[slot] & kLayerTsdfBit) != 0;
    if (has_tsdf) {
      if (m.blk_dirty[slot] & mesh_dirty_bit<KIND>()) {
        MeshArena none = {nullptr, nullptr, nullptr, nullptr};
        mesh_one_block<false, KIND>(m, mp, s, slot, 0, 0, none, &nv, &nt);
        ++remeshed;
      } else {
        const int4 old = mesh_extents<KIND>(m)[slot];
        nv = old.y;
        nt = old.w;
      }
    }
    if (threadIdx.x == 0) {
      cnt_v[slot] = nv;
      cnt_t[slot] = nt;
    }
    __syncthreads();
  }
  if (KIND == kMeshFeature && threadIdx.x == 0 && remeshed) count_add(m, kCntMeshBlocksRemeshed, remeshed);
}

// device-side exclusive scan over the slot table (single CTA; the table is small)
__global__ void __launch_bounds__(1024) k_mesh_scan(MapDev m, const int* cnt_v, const int* cnt_t, int* off_v,
                                                    int* off_t, int kind) {
  pdl_prologue();
  __shared__ int ws_v[33], ws_t[33];
  __shared__ int carry_v, carry_t;
  const int n = m.ctrl->slot_high;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    carry_v = 0;
    carry_t = 0;
  }
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? cnt_v[i] : 0, tt = i < n ? cnt_t[i] : 0;
    int iv = v, it = tt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, iv, o), b2 = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) {
        iv += a;
        it += b2;
      }
    }
    if (lane == 31) {
      ws_v[warp + 1] = iv;
      ws_t[warp + 1] = it;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int av = carry_v, at = carry_t;
      ws_v[0] = av;
      ws_t[0] = at;
      for (int w = 1; w <= 32; ++w) {
        av += ws_v[w];
        ws_v[w] = av;
        at += ws_t[w];
        ws_t[w] = at;
      }
      carry_v = av;
      carry_t = at;
    }
    __syncthreads();
    if (i < n) {
      off_v[i] = ws_v[warp] + iv - v;
      off_t[i] = ws_t[warp] + it - tt;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    m.ctrl->mesh_total_v = carry_v;
    m.ctrl->mesh_total_t = carry_t;
    if (kind == kMeshFeature) m.ctrl->counters[kCntMeshVertices] = (unsigned long long)carry_v;
  }
}

// pass 2: write the new arena; clean blocks are copied from the previous arena
template <int KIND>
__global__ void __launch_bounds__(512, 2) k_mesh_emit(MapDev m, MeshParams mp, const int* off_v, const int* off_t,
                                                      MeshArena prev, MeshArena out) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MeshSmem& s = *reinterpret_cast<MeshSmem*>(smem_raw);
  const int n = m.ctrl->slot_high;
  const int t = threadIdx.x;
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    if (!(m.blk_layers[slot] & kLayerTsdfBit)) continue;
    const int voff = off_v[slot], toff = off_t[slot];
    int nv = 0, nt = 0;
    if (m.blk_dirty[slot] & mesh_dirty_bit<KIND>()) {
      mesh_one_block<true, KIND>(m, mp, s, slot, voff, toff, out, &nv, &nt);
    } else {
      const int4 old = mesh_extents<KIND>(m)[slot];
      nv = old.y;
      nt = old.w;
      for (int i = t; i < nv * 3; i += 512) out.verts[(size_t)voff * 3 + i] = prev.verts[(size_t)old.x * 3 + i];
      const int shift = voff - old.x;
      for (int i = t; i < nt; i += 512) out.tris[toff + i] = prev.tris[old.z + i] + shift;
      if (KIND == kMeshFeature) {
        const size_t nvecs = (size_t)nv * (size_t)(m.C >> 3);
        const uint4* src = reinterpret_cast<const uint4*>(prev.feats + (size_t)old.x * m.C);
        uint4* dst = reinterpret_cast<uint4*>(out.feats + (size_t)voff * m.C);
        for (size_t i = t; i < nvecs; i += 512) dst[i] = src[i];
      } else {
        for (int i = t; i < nv * 3; i += 512) out.colors[(size_t)voff * 3 + i] = prev.colors[(size_t)old.x * 3 + i];
      }
    }
    __syncthreads();
    if (t == 0) {
      mesh_extents<KIND>(m)[slot] = make_int4(voff, nv, toff, nt);
      m.blk_dirty[slot] &= (uint8_t)~mesh_dirty_bit<KIND>();  // markBlocksAsUpdated (mapper.cpp:611)
    }
    __syncthreads();
  }
}

}  // namespace nvbx
