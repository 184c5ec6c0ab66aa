// nvbx_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch restatement, in plain C++17 with no dependencies, of the algorithm nvblox runs for
// mindmap's reconstruction hot path: depth -> TSDF, feature frame -> feature layer, decay, feature
// mesh.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (libnvbx.so) never links, imports or calls it.
//
// PARITY STATUS: pinned (a) by the reference's own known-answer tests, restated in
// tests/test_oracle_known_answers.py (test_feature_integrator.cpp:131-203, test_ray_caster.cpp,
// test_interpolation_2d.cpp, test_weighting_function / test_tsdf_integrator.cpp:359-474,
// test_mapper_masking.py:35-80,170-198, test_tsdf_decay.cpp, test_mesh.cpp), and (b) BIT FOR BIT by
// reference-compiled vectors: the reference's own device kernels (integrateBlocksKernel for TSDF / feature /
// colour voxels, combinedBlockIndicesInImageKernel, sphereTracingKernel, the marching-cubes and mesh-appearance
// kernels) compiled verbatim with nvblox's nvcc flags (oracle/ref_snippets/) and run on a B200 on seeded
// inputs (tests/golden/make_ref_vectors.py -> tests/golden/ref_nvcc_*.npz; tests/test_ref_vectors.py).
// The whole nvblox library cannot be built offline (Eigen, stdgpu, glog are fetched at CMake time), so the
// reference's HOST code (view AABB, planes view, viewpoint cache, allocation) is pinned by (a) only, and the
// Eigen expressions inside the kernels were compiled against a stand-in that evaluates them in Eigen 3.4's
// order (oracle/ref_snippets/shim/Eigen/Core) -- that residual is stated in DESIGN.md section 5.
//
// Floating-point model (orc_set_fp_model):
//   1 = "nvcc" (DEFAULT): what the reference BINARY computes.  nvblox passes no -fmad flag
//       (NB/cmake/nvblox_targets.cmake:128-152), so nvcc -O2 contracts mul+add pairs in DEVICE code into
//       single-rounding fma (NVVM contracts a*b+c; ptxas fuses the mul.f32 / add.f32 and mul.f16 / add.f16 it
//       is left with).  Every contracted expression below is written as an explicit fmaf / hfma in the
//       operand order read from the PTX / SASS of oracle/ref_snippets (profiles/r02_ref_contraction.md).
//       HOST code of the reference is gcc x86-64 -O2 without -mfma: never contracted.
//   0 = "ieee": every operation individually rounded (round 1's model), kept for comparison.
// fp16 arithmetic is one RNE rounding per operator (__hmul/__hadd/__hsub) or per fused hfma.
// Build with -ffp-contract=off so the compiler adds no contraction of its own.
//
// Every function cites the reference file:line (relative to /root/reference) it follows.
//   NB/ = submodules/nvblox/nvblox/      NT/ = submodules/nvblox/nvblox_torch/

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <set>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/nvbx_c_api.h"

#define NVBX_MC_QUAL static const
#include "../nvblox_mindmap_b200/csrc/mc_tables.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Small math types.  No operator overloading on purpose: the operation order is the specification.
// ------------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
struct I3 {
  int x, y, z;
  bool operator<(const I3& o) const {
    if (x != o.x) return x < o.x;
    if (y != o.y) return y < o.y;
    return z < o.z;
  }
  bool operator==(const I3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Pose {  // Eigen::Isometry3f: p_out = t + R p
  float R[3][3];
  float t[3];
};
struct Cam {  // NB/include/nvblox/sensors/camera.h
  float fu, fv, cu, cv;
  int width, height;
};

// Eigen redux of 3 elements: a0 + (a1 + a2)   (Eigen/src/Core/Redux.h, redux_novec_unroller)
inline float sum3(float a, float b, float c) { return a + (b + c); }

// Transform * Vector3f  (Eigen/src/Geometry/Transform.h transform_right_product_impl: res = t; res += R*v)
inline V3 xform(const Pose& T, const V3& v) {
  V3 o;
  o.x = T.t[0] + sum3(T.R[0][0] * v.x, T.R[0][1] * v.y, T.R[0][2] * v.z);
  o.y = T.t[1] + sum3(T.R[1][0] * v.x, T.R[1][1] * v.y, T.R[1][2] * v.z);
  o.z = T.t[2] + sum3(T.R[2][0] * v.x, T.R[2][1] * v.y, T.R[2][2] * v.z);
  return o;
}
inline V3 rotate(const Pose& T, const V3& v) {
  V3 o;
  o.x = sum3(T.R[0][0] * v.x, T.R[0][1] * v.y, T.R[0][2] * v.z);
  o.y = sum3(T.R[1][0] * v.x, T.R[1][1] * v.y, T.R[1][2] * v.z);
  o.z = sum3(T.R[2][0] * v.x, T.R[2][1] * v.y, T.R[2][2] * v.z);
  return o;
}
// Isometry3f::inverse(): R' = R^T, t' = -(R^T) t   (Eigen/src/Geometry/Transform.h inverse(Isometry))
inline Pose inverse(const Pose& T) {
  Pose o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.R[i][j] = T.R[j][i];
  for (int i = 0; i < 3; ++i)
    o.t[i] = sum3((-o.R[i][0]) * T.t[0], (-o.R[i][1]) * T.t[1], (-o.R[i][2]) * T.t[2]);
  return o;
}
inline Pose pose_from_row_major(const float* m) {  // NT/cpp/src/convert_tensors.cpp:119-129
  Pose p;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.R[i][j] = m[i * 4 + j];
    p.t[i] = m[i * 4 + 3];
  }
  return p;
}

// ------------------------------------------------------------------------------------------------
// Floating-point model switch and the DEVICE-code variants of the geometry (see the header).
// Operand orders are the ones nvcc 12.9 -O2 emits for the reference's expressions
// (oracle/ref_snippets, profiles/r02_ref_contraction.md).
// ------------------------------------------------------------------------------------------------
int g_fp_model = 1;
inline bool nvcc_model() { return g_fp_model != 0; }

// R.row(i) . v on the device: Eigen's a0 + (a1 + a2) becomes fma(v.x, R0, fma(v.y, R1, v.z * R2))
inline float dev_dot3(const float* r, const V3& v) {
  if (!nvcc_model()) return sum3(r[0] * v.x, r[1] * v.y, r[2] * v.z);
  return std::fmaf(v.x, r[0], std::fmaf(v.y, r[1], v.z * r[2]));
}
inline V3 dev_rotate(const Pose& T, const V3& v) { return V3{dev_dot3(T.R[0], v), dev_dot3(T.R[1], v), dev_dot3(T.R[2], v)}; }
// Transform * Vector3f on the device: the translation is added last, unfused (add.f32 of the fma chain)
inline V3 dev_xform(const Pose& T, const V3& v) {
  return V3{T.t[0] + dev_dot3(T.R[0], v), T.t[1] + dev_dot3(T.R[1], v), T.t[2] + dev_dot3(T.R[2], v)};
}
inline float dev_fma(float a, float b, float c) { return nvcc_model() ? std::fmaf(a, b, c) : a * b + c; }

// ------------------------------------------------------------------------------------------------
// Software binary16.  Conversions are round-to-nearest-even like cvt.rn.f16.f32; NaN -> 0x7fff like
// CUDA's __float2half.  +,-,* of two halves computed in fp32 and rounded once to half are correctly
// rounded (24 >= 2*11+2, Figueroa), which is what __hadd/__hsub/__hmul return.
// ------------------------------------------------------------------------------------------------
inline uint16_t f2h(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t absx = x & 0x7fffffffu;
  if (absx > 0x7f800000u) return 0x7fff;                          // NaN
  if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);     // >= 65520 rounds to inf
  if (absx < 0x33000001u) return (uint16_t)sign;                  // <= 2^-25 rounds to zero (tie -> even 0)
  int exp = (int)(absx >> 23) - 127;
  uint32_t man = (absx & 0x7fffffu) | 0x800000u;  // 24-bit significand
  int shift;                                       // bits to drop
  uint32_t hexp;
  if (exp < -14) {  // subnormal half
    shift = 13 + (-14 - exp);
    hexp = 0;
  } else {
    shift = 13;
    hexp = (uint32_t)(exp + 15);
  }
  uint32_t kept = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u);
  const uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (kept & 1u))) kept += 1;
  uint32_t out;
  if (hexp == 0) {
    out = kept;  // may carry into exponent 1: that is the right encoding
  } else {
    out = ((hexp - 1) << 10) + kept;  // kept has the implicit bit at position 10
  }
  return (uint16_t)(sign | out);
}
inline float h2f(uint16_t h) {
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  const uint32_t exp = (h >> 10) & 0x1f;
  const uint32_t man = h & 0x3ffu;
  uint32_t x;
  if (exp == 0) {
    if (man == 0) {
      x = sign;
    } else {
      float f = (float)man * 5.9604644775390625e-8f;  // 2^-24, exact
      std::memcpy(&x, &f, 4);
      x |= sign;
    }
  } else if (exp == 31) {
    x = sign | 0x7f800000u | (man << 13);
  } else {
    x = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}
inline uint16_t hadd(uint16_t a, uint16_t b) { return f2h(h2f(a) + h2f(b)); }
inline uint16_t hsub(uint16_t a, uint16_t b) { return f2h(h2f(a) - h2f(b)); }
inline uint16_t hmul(uint16_t a, uint16_t b) { return f2h(h2f(a) * h2f(b)); }
// fma variant (a*b exact in double; the sum is rounded to double, then to half -- exact except in
// astronomically unlikely double-rounding ties; informational only).
inline uint16_t hfma(uint16_t a, uint16_t b, uint16_t c) {
  const double r = (double)h2f(a) * (double)h2f(b) + (double)h2f(c);
  // round double -> half through float with sticky handling
  float f = (float)r;
  if ((double)f != r && std::isfinite(f)) {  // round-to-odd so the second rounding (to half) is safe
    uint32_t x;
    std::memcpy(&x, &f, 4);
    if (std::fabs((double)f) > std::fabs(r)) x -= 1;  // truncate toward zero
    x |= 1u;                                          // sticky bit
    std::memcpy(&f, &x, 4);
  }
  return f2h(f);
}

// ORC "fused half": follows the floating-point model unless forced (orc_set_fused_half: -1 follow, 0 off, 1 on)
int g_fused_half_override = -1;
inline bool fused_half_on() { return g_fused_half_override < 0 ? nvcc_model() : g_fused_half_override != 0; }

// NB/include/nvblox/interpolation/internal/impl/interpolation_2d_impl.h:33-48 with FloatType = __half:
//   x = half(off.x), y = half(off.y), dx = f10 - f00
//   out = f00 + x*dx + y*(f01 - f00) + x*y*(f11 - f01 - dx)          (C++ precedence, left to right)
inline uint16_t interp_half(uint16_t x, uint16_t y, uint16_t f00, uint16_t f01, uint16_t f10, uint16_t f11) {
  const uint16_t dx = hsub(f10, f00);
  if (!fused_half_on()) {
    const uint16_t t2 = hadd(f00, hmul(x, dx));
    const uint16_t t5 = hadd(t2, hmul(y, hsub(f01, f00)));
    const uint16_t t9 = hmul(hmul(x, y), hsub(hsub(f11, f01), dx));
    return hadd(t5, t9);
  } else {
    // SASS of the reference: HFMA2(x, dx, f00); HFMA2(y, f01 - f00, .); HMUL2 x*y; HFMA2(xy, f11 - f01 - dx, .)
    const uint16_t t2 = hfma(x, dx, f00);
    const uint16_t t5 = hfma(y, hsub(f01, f00), t2);
    return hfma(hmul(x, y), hsub(hsub(f11, f01), dx), t5);
  }
}
// same formula in fp32 (FloatType = float), used for the synthetic depth image
inline float interp_float(float x, float y, float f00, float f01, float f10, float f11) {
  const float dx = f10 - f00;
  if (!nvcc_model()) return ((f00 + x * dx) + y * (f01 - f00)) + (x * y) * ((f11 - f01) - dx);
  // device code (the only caller of the fp32 form is the appearance kernel): three FFMAs
  return std::fmaf(x * y, (f11 - f01) - dx, std::fmaf(y, f01 - f00, std::fmaf(x, dx, f00)));
}

// ------------------------------------------------------------------------------------------------
// Indexing.  NB/include/nvblox/core/internal/impl/indexing_impl.h:22-81
// ------------------------------------------------------------------------------------------------
inline I3 block_index_from_position(float block_size, const V3& p) {  // :32-36
  return I3{(int)std::floor(p.x / block_size), (int)std::floor(p.y / block_size),
            (int)std::floor(p.z / block_size)};
}
inline void block_and_voxel_from_position(float block_size, const V3& p, I3* b, I3* v) {  // :37-49
  const float voxel_size = block_size * (1.0f / 8.0f);
  const float inv = (float)(1.0 / (double)voxel_size);
  *b = block_index_from_position(block_size, p);
  const float rx = (p.x - block_size * (float)b->x) * inv;
  const float ry = (p.y - block_size * (float)b->y) * inv;
  const float rz = (p.z - block_size * (float)b->z) * inv;
  v->x = std::min((int)rx, 7);
  v->y = std::min((int)ry, 7);
  v->z = std::min((int)rz, 7);
}
// device form of the same (sphere tracer, query kernel): p - bs*b is one FFMA
inline void dev_block_and_voxel_from_position(float block_size, const V3& p, I3* b, I3* v) {
  if (!nvcc_model()) return block_and_voxel_from_position(block_size, p, b, v);
  const float voxel_size = block_size * (1.0f / 8.0f);
  const float inv = (float)(1.0 / (double)voxel_size);
  *b = block_index_from_position(block_size, p);
  const float rx = std::fmaf(-(float)b->x, block_size, p.x) * inv;
  const float ry = std::fmaf(-(float)b->y, block_size, p.y) * inv;
  const float rz = std::fmaf(-(float)b->z, block_size, p.z) * inv;
  v->x = std::min((int)rx, 7);
  v->y = std::min((int)ry, 7);
  v->z = std::min((int)rz, 7);
}
// device form of the voxel centre (projectThreadVoxel): fma(bs, 1/16, fma(bs, b, (bs/8) * v))
inline V3 dev_voxel_center(float block_size, const I3& b, const I3& v);
inline V3 voxel_center(float block_size, const I3& b, const I3& v) {  // :51-81
  const float voxel_size = block_size * (1.0f / 8.0f);
  const float half_voxel = block_size * (0.5f / 8.0f);
  V3 p;
  p.x = (block_size * (float)b.x + voxel_size * (float)v.x) + half_voxel;
  p.y = (block_size * (float)b.y + voxel_size * (float)v.y) + half_voxel;
  p.z = (block_size * (float)b.z + voxel_size * (float)v.z) + half_voxel;
  return p;
}

inline V3 dev_voxel_center(float block_size, const I3& b, const I3& v) {
  if (!nvcc_model()) return voxel_center(block_size, b, v);
  const float voxel_size = block_size * (1.0f / 8.0f);
  V3 p;
  p.x = std::fmaf(block_size, 0.5f / 8.0f, std::fmaf(block_size, (float)b.x, voxel_size * (float)v.x));
  p.y = std::fmaf(block_size, 0.5f / 8.0f, std::fmaf(block_size, (float)b.y, voxel_size * (float)v.y));
  p.z = std::fmaf(block_size, 0.5f / 8.0f, std::fmaf(block_size, (float)b.z, voxel_size * (float)v.z));
  return p;
}

// ------------------------------------------------------------------------------------------------
// Camera.  NB/include/nvblox/sensors/internal/impl/camera_impl.h:20-91, NB/src/sensors/camera.cpp:51-116
// ------------------------------------------------------------------------------------------------
inline V3 ray_from_image_plane(const Cam& c, float u, float v) {  // camera_impl.h:71-80
  return V3{(u - c.cu) / c.fu, (v - c.cv) / c.fv, 1.0f};
}
inline bool project(const Cam& c, const V3& p, float* u, float* v) {  // camera_impl.h:25-54
  if (!(p.z >= 1e-6f)) return false;  // projectToNormalizedCoordinates, min_depth default camera.h:158
  float un = p.x / p.z;
  float vn = p.y / p.z;
  un = un * c.fu + c.cu;
  vn = vn * c.fv + c.cv;
  if (un > (float)c.width || vn > (float)c.height || un < 0 || vn < 0) return false;
  *u = un;
  *v = vn;
  return true;
}
// Camera::project inside a kernel: the intrinsics are applied with one FFMA each
inline bool dev_project(const Cam& c, const V3& p, float* u, float* v) {
  if (!nvcc_model()) return project(c, p, u, v);
  if (!(p.z >= 1e-6f)) return false;
  const float un = std::fmaf(p.x / p.z, c.fu, c.cu);
  const float vn = std::fmaf(p.y / p.z, c.fv, c.cv);
  if (un > (float)c.width || vn > (float)c.height || un < 0 || vn < 0) return false;
  *u = un;
  *v = vn;
  return true;
}
struct Aabb {
  float mn[3], mx[3];
  bool empty() const { return mn[0] > mx[0] || mn[1] > mx[1] || mn[2] > mx[2]; }  // Eigen::AlignedBox::isEmpty
};
// Camera::getViewCorners + getViewAABB (camera.cpp:51-103) and Frustum AABB (camera.cpp:153-166) are the
// same min/max over the 8 transformed corners.
inline Aabb view_aabb(const Cam& c, const Pose& T_L_C, float min_depth, float max_depth) {
  const V3 rays[4] = {ray_from_image_plane(c, (float)c.width, (float)c.height),
                      ray_from_image_plane(c, (float)c.width, 0.0f), ray_from_image_plane(c, 0.0f, 0.0f),
                      ray_from_image_plane(c, 0.0f, (float)c.height)};
  Aabb a;
  for (int i = 0; i < 3; ++i) {
    a.mn[i] = std::numeric_limits<float>::max();
    a.mx[i] = std::numeric_limits<float>::lowest();
  }
  for (int k = 0; k < 8; ++k) {
    const float d = k < 4 ? min_depth : max_depth;
    const V3& r = rays[k & 3];
    const V3 pc{d * r.x, d * r.y, d * r.z};
    const V3 pl = xform(T_L_C, pc);
    const float v[3] = {pl.x, pl.y, pl.z};
    for (int i = 0; i < 3; ++i) {
      a.mn[i] = std::min(a.mn[i], v[i]);
      a.mx[i] = std::max(a.mx[i], v[i]);
    }
  }
  return a;
}
// NB/src/geometry/workspace_bounds.cpp:20-61
inline bool apply_workspace_bounds(const nvbx_params& p, Aabb* a) {
  if (p.workspace_bounds_type == NVBX_WORKSPACE_HEIGHT_BOUNDS) {
    a->mn[2] = std::max(a->mn[2], p.workspace_min[2]);
    a->mx[2] = std::min(a->mx[2], p.workspace_max[2]);
  } else if (p.workspace_bounds_type == NVBX_WORKSPACE_BOUNDING_BOX) {
    for (int i = 0; i < 3; ++i) {  // workspace.intersection(input): cwiseMax of mins, cwiseMin of maxes
      a->mn[i] = std::max(p.workspace_min[i], a->mn[i]);
      a->mx[i] = std::min(p.workspace_max[i], a->mx[i]);
    }
  }
  return !a->empty();
}

// arePosesClose, NB/src/geometry/transforms.cpp:20-37 (Eigen::AngleAxisf(R).angle() through a quaternion)
inline bool poses_close(const Pose& A, const Pose& B, float tol_m, float tol_deg) {
  const Pose Ai = inverse(A);
  // T = Ai * B
  float R[3][3];
  float t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = sum3(Ai.R[i][0] * B.R[0][j], Ai.R[i][1] * B.R[1][j], Ai.R[i][2] * B.R[2][j]);
    t[i] = Ai.t[i] + sum3(Ai.R[i][0] * B.t[0], Ai.R[i][1] * B.t[1], Ai.R[i][2] * B.t[2]);
  }
  const float n = std::sqrt(sum3(t[0] * t[0], t[1] * t[1], t[2] * t[2]));
  if (n > tol_m) return false;
  // Eigen quaternion from rotation matrix (Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl)
  float qw, qx, qy, qz;
  float tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0f) {
    tr = std::sqrt(tr + 1.0f);
    qw = 0.5f * tr;
    tr = 0.5f / tr;
    qx = (R[2][1] - R[1][2]) * tr;
    qy = (R[0][2] - R[2][0]) * tr;
    qz = (R[1][0] - R[0][1]) * tr;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    tr = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0f);
    float q[3];
    q[i] = 0.5f * tr;
    tr = 0.5f / tr;
    qw = (R[k][j] - R[j][k]) * tr;
    q[j] = (R[j][i] + R[i][j]) * tr;
    q[k] = (R[k][i] + R[i][k]) * tr;
    qx = q[0];
    qy = q[1];
    qz = q[2];
  }
  // AngleAxis from quaternion: angle = 2*atan2(|vec|, |w|)
  const float vn = std::sqrt(sum3(qx * qx, qy * qy, qz * qz));
  const float angle = 2.0f * std::atan2(vn, std::fabs(qw));
  const float deg = (float)((double)(angle * 180.0f) / M_PI);
  return !(std::fabs(deg) > tol_deg);
}
// areCamerasEqual, NB/src/sensors/camera.cpp:31-49
inline bool cameras_equal(const Cam& a, const Cam& b, const Pose& Ta, const Pose& Tb) {
  const bool ext = poses_close(Ta, Tb, 0.001f, 0.1f);
  bool in = true;
  in &= std::fabs(a.fu - b.fu) <= 0.1;
  in &= std::fabs(a.fv - b.fv) <= 0.1;
  in &= std::fabs(a.cu - b.cu) <= 0.1;
  in &= std::fabs(a.cv - b.cv) <= 0.1;
  in &= a.width == b.width;
  in &= a.height == b.height;
  return ext && in;
}

// ViewpointCache, NB/src/integrators/view_calculator.cu:472-541 (kMaxCacheSize = 2, view_calculator.h:271)
struct ViewCache {
  std::deque<Pose> poses;
  std::deque<Cam> cams;
  std::deque<std::vector<I3>> results;
  const std::vector<I3>* get(const Pose& T, const Cam& c) const {
    for (size_t i = 0; i < cams.size(); ++i)
      if (cameras_equal(c, cams[i], T, poses[i])) return &results[i];
    return nullptr;
  }
  void store(const Pose& T, const Cam& c, const std::vector<I3>& r) {
    if (cams.size() == 2) {
      poses.pop_back();
      cams.pop_back();
      results.pop_back();
    }
    poses.push_front(T);
    cams.push_front(c);
    results.push_front(r);
  }
  void clear() {
    poses.clear();
    cams.clear();
    results.clear();
  }
};

// ------------------------------------------------------------------------------------------------
// RayCaster: 3-D DDA.  NB/include/nvblox/rays/internal/impl/ray_caster_impl.h:26-75 (scale = 1)
// ------------------------------------------------------------------------------------------------
inline int signum(float x) { return (x > 0.0f) ? 1 : ((x < 0.0f) ? -1 : 0); }
struct RayCaster {
  int cur[3];
  int sign[3];
  int steps, step;
  float t_next[3], t_step[3];
  RayCaster(const V3& origin, const V3& dest) {
    const float s[3] = {origin.x / 1.0f, origin.y / 1.0f, origin.z / 1.0f};
    const float e[3] = {dest.x / 1.0f, dest.y / 1.0f, dest.z / 1.0f};
    int end[3];
    for (int i = 0; i < 3; ++i) {
      cur[i] = (int)std::floor(s[i]);
      end[i] = (int)std::floor(e[i]);
    }
    steps = std::abs(end[0] - cur[0]) + std::abs(end[1] - cur[1]) + std::abs(end[2] - cur[2]);
    step = 0;
    for (int i = 0; i < 3; ++i) {
      const float ray = e[i] - s[i];
      sign[i] = signum(ray);
      const int corrected = std::max(sign[i], 0);
      const float shifted = s[i] - (float)cur[i];
      const float dist = (float)corrected - shifted;
      t_next[i] = dist / ray;  // NaN / inf are "fine" per the reference comment
      t_step[i] = (float)sign[i] / ray;
    }
  }
  bool next(int out[3]) {
    if (step++ > steps) return false;
    out[0] = cur[0];
    out[1] = cur[1];
    out[2] = cur[2];
    // Eigen minCoeff visitor: start at 0, replace on strict '<' (NaN never wins, a NaN at 0 never loses)
    int m = 0;
    if (t_next[1] < t_next[m]) m = 1;
    if (t_next[2] < t_next[m]) m = 2;
    cur[m] += sign[m];
    t_next[m] += t_step[m];
    return true;
  }
};

// ------------------------------------------------------------------------------------------------
// Reference-kernel hooks (tests/golden/make_ref_vectors.py only).  When a hook is set, the oracle keeps
// doing the reference's HOST work (view AABB, caches, allocation, candidate lists, band filter) but hands
// the DEVICE stage to the callback, which runs the reference's own kernel (oracle/ref_snippets) on a GPU.
// The maps such a run produces are the reference-compiled golden vectors the un-hooked oracle and the
// CUDA product are compared against.  All pointers are host memory; poses are row-major 4x4; cam = fu,fv,cu,cv.
// ------------------------------------------------------------------------------------------------
extern "C" {
typedef int (*orc_hook_raycast_t)(const float* T_L_C, const float* cam, int W, int H, const float* depth,
                                  float block_size, float max_dist, float behind, int subsampling,
                                  const int* aabb_min, const int* aabb_size, uint8_t* grid);
typedef int (*orc_hook_tsdf_t)(const float* T_C_L, const float* cam, int W, int H, const float* depth,
                               const uint8_t* mask, float block_size, float max_dist, float trunc, float max_weight,
                               float invalid_decay, int mode, const int* block_idx, int n, float* voxels);
typedef int (*orc_hook_trace_t)(const float* T_L_C, const float* cam, int W, int H, const int* all_idx, int n_all,
                                const float* voxels, float trunc, float block_size, int max_steps, float max_len,
                                float eps, int subsample, float* out);
typedef int (*orc_hook_feat_t)(const float* T_C_L, const float* cam, int W, int H, const uint16_t* img,
                               const uint8_t* mask, const float* synth, int sub, float block_size, float max_dist,
                               float trunc, float max_weight, float alpha, const int* block_idx, int n, uint16_t* fvox);
typedef int (*orc_hook_color_t)(const float* T_C_L, const float* cam, int W, int H, const uint8_t* img,
                                const uint8_t* mask, const float* synth, int sub, float block_size, float max_dist,
                                float trunc, float max_weight, float alpha, const int* block_idx, int n, uint8_t* rgb,
                                float* weight);
}
struct RefHooks {
  orc_hook_raycast_t raycast = nullptr;
  orc_hook_tsdf_t tsdf = nullptr;
  orc_hook_trace_t trace = nullptr;
  orc_hook_feat_t feat = nullptr;
  orc_hook_color_t color = nullptr;
} g_hooks;
inline void pose_to_row_major(const Pose& T, float* m) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) m[i * 4 + j] = T.R[i][j];
    m[i * 4 + 3] = T.t[i];
  }
  m[12] = m[13] = m[14] = 0.0f;
  m[15] = 1.0f;
}

// ------------------------------------------------------------------------------------------------
// Map storage
// ------------------------------------------------------------------------------------------------
struct TsdfBlock {
  float d[512];
  float w[512];
  TsdfBlock() {
    std::memset(d, 0, sizeof(d));
    std::memset(w, 0, sizeof(w));
  }
};
inline int vlin(int x, int y, int z) { return (x * 8 + y) * 8 + z; }  // voxels[x][y][z], blox.h:28-67

// ColorVoxel {Color color = Gray (127,127,127); float weight = 0}, NB/include/nvblox/map/voxels.h:77-83,
// gray initialisation NB/src/map/blox.cu:32-54
struct ColorBlock {
  uint8_t c[512][3];
  float w[512];
  ColorBlock() {
    std::memset(c, 127, sizeof(c));
    std::memset(w, 0, sizeof(w));
  }
};

struct MeshBlock {
  std::vector<V3> verts;
  std::vector<int> tris;             // block-local vertex ids, size = unwelded vertex count
  std::vector<uint16_t> feats;       // verts.size() * C     (feature mesh layer)
  std::vector<uint8_t> colors;       // verts.size() * 3     (colour mesh layer)
};

struct Counters {
  int64_t v[16];
};

struct Oracle {
  float voxel_size;
  float block_size;
  int C;
  nvbx_params p;
  std::map<I3, TsdfBlock> tsdf;
  std::map<I3, std::vector<uint16_t>> feat;  // 512 * (C+1): [voxel][0..C) feature, [C] weight
  std::map<I3, ColorBlock> color;
  std::map<I3, MeshBlock> mesh;    // MeshBlockLayer<FeatureArray>
  std::map<I3, MeshBlock> cmesh;   // MeshBlockLayer<Color> -- a separate layer with its own geometry
  std::set<I3> mesh_dirty;   // BlocksToUpdateTracker::feature_mesh_blocks_to_update_
  std::set<I3> cmesh_dirty;  // BlocksToUpdateTracker::color_mesh_blocks_to_update_
  // Mapper::Mapper makes its TSDF, colour and feature integrators SHARE one raycasting and one planes viewpoint
  // cache (shareViewpointCaches, mapper.cpp:56-58, mapper_common_impl.h:20-34): a colour frame and a feature frame
  // at (nearly) the same pose see each other's cached block list.
  ViewCache raycast_cache, planes_cache;
  std::vector<I3> last_tsdf_list, last_feat_list, last_color_list;
  std::vector<float> synth;
  int synth_rows = 0, synth_cols = 0;
  nvbx_counters cnt;
  std::unordered_set<uint64_t> pixel_seen;  // distinct feature pixels of the last feature frame
  int64_t last_distinct_pixels = 0;
  int64_t last_trace_steps = 0, last_trace_max_steps = 0;  // sphere-tracing work of the last feature frame
  int64_t last_n_upd = 0;
};

inline float trunc_tsdf(const Oracle& o) { return o.p.truncation_distance_vox * o.voxel_size; }

// ------------------------------------------------------------------------------------------------
// a2: blocks in view by ray casting.  NB/src/integrators/view_calculator.cu:155-248,250-331,355-390
// ------------------------------------------------------------------------------------------------
std::vector<I3> blocks_in_view_raycast(Oracle& o, const float* depth, int rows, int cols, const Pose& T_L_C,
                                       const Cam& cam) {
  if (o.p.cache_last_viewpoint) {
    if (const std::vector<I3>* hit = o.raycast_cache.get(T_L_C, cam)) return *hit;
  }
  const float max_dist = o.p.max_integration_distance_m;
  const float behind = trunc_tsdf(o);
  Aabb aabb = view_aabb(cam, T_L_C, 0.0f, max_dist);
  if (!apply_workspace_bounds(o.p, &aabb)) return {};  // note: not cached (view_calculator.cu:275-279)
  const float bs = o.block_size;
  const I3 mn = block_index_from_position(bs, V3{aabb.mn[0], aabb.mn[1], aabb.mn[2]});
  const I3 mx = block_index_from_position(bs, V3{aabb.mx[0], aabb.mx[1], aabb.mx[2]});
  const int sx = mx.x - mn.x + 1, sy = mx.y - mn.y + 1, sz = mx.z - mn.z + 1;
  const size_t linear_size = (size_t)(sx * sy * sz);
  std::vector<uint8_t> grid(linear_size, 0);
  // setIndexUpdated (:171-180): the linear index is computed in int and converted to size_t; only
  // `lin < linear_size` is tested, so out-of-box indices with an in-range linear index ALIAS.
  auto mark = [&](int x, int y, int z) {
    const int lx = x - mn.x, ly = y - mn.y, lz = z - mn.z;
    const size_t lin = (size_t)(int)(lx + ly * sx + lz * sx * sy);
    if (lin < linear_size) grid[lin] = 1;
  };
  const int s = o.p.raycast_subsampling_factor;
  const int n_rows = (int)std::ceil((float)(rows + 1) / (float)s);
  const int n_cols = (int)std::ceil((float)(cols + 1) / (float)s);
  // launch covers ceil(n/16)*16 threads per dimension; threads beyond n_rows/n_cols may still pass the
  // in-kernel test `pixel >= rows + s - 1` (:211-214), so iterate over the launched extent.
  const int launched_rows = (int)std::ceil(n_rows / 16.0f) * 16;
  const int launched_cols = (int)std::ceil(n_cols / 16.0f) * 16;
  const V3 t_L{T_L_C.t[0], T_L_C.t[1], T_L_C.t[2]};
  const V3 origin_scaled{t_L.x / bs, t_L.y / bs, t_L.z / bs};
  if (g_hooks.raycast) {
    float Tm[16];
    pose_to_row_major(T_L_C, Tm);
    const float cm[4] = {cam.fu, cam.fv, cam.cu, cam.cv};
    const int amn[3] = {mn.x, mn.y, mn.z}, asz[3] = {sx, sy, sz};
    if (g_hooks.raycast(Tm, cm, cols, rows, depth, bs, max_dist, behind, s, amn, asz, grid.data()))
      std::fprintf(stderr, "orc: raycast hook failed\n");
  } else
  for (int rr = 0; rr < launched_rows; ++rr) {
    for (int rc = 0; rc < launched_cols; ++rc) {
      int prow = rr * s, pcol = rc * s;
      if (prow >= rows + s - 1 || pcol >= cols + s - 1) continue;
      if (prow >= rows) prow = rows - 1;
      if (pcol >= cols) pcol = cols - 1;
      float d = depth[(size_t)prow * cols + pcol];
      if (d <= 0.0f) continue;  // NaN compares false -> NaN depth proceeds, as in the reference
      if (max_dist > 0.0f && d > max_dist) d = max_dist;
      const V3 ray = ray_from_image_plane(cam, (float)pcol + 0.5f, (float)prow + 0.5f);
      const float len = d + behind;
      const V3 p_C{len * ray.x, len * ray.y, len * ray.z};
      const V3 p_L = dev_xform(T_L_C, p_C);   // device code (K1)
      const I3 b = block_index_from_position(bs, p_L);
      mark(b.x, b.y, b.z);
      RayCaster rc3(origin_scaled, V3{p_L.x / bs, p_L.y / bs, p_L.z / bs});
      int idx[3];
      while (rc3.next(idx)) mark(idx[0], idx[1], idx[2]);
    }
  }
  std::vector<I3> out;
  for (size_t i = 0; i < linear_size; ++i) {
    if (grid[i]) {
      const int ix = (int)(i % (size_t)sx), iy = (int)((i / (size_t)sx) % (size_t)sy), iz = (int)(i / (size_t)(sx * sy));
      out.push_back(I3{ix + mn.x, iy + mn.y, iz + mn.z});
    }
  }
  if (o.p.cache_last_viewpoint) o.raycast_cache.store(T_L_C, cam, out);
  return out;
}

// ------------------------------------------------------------------------------------------------
// a4: TSDF update.  projective_integrator_impl.cuh:58-103, projective_integrators_common_impl.cuh:21-55,
// NB/src/integrators/projective_tsdf_integrator.cu:30-90, weighting_function_impl.h:29-117
// ------------------------------------------------------------------------------------------------
inline float weight_dropoff(float measured, float voxel_depth, float trunc) {
  if (trunc <= 1e-2f) return 0.0f;
  if (voxel_depth > measured) {
    const float behind = voxel_depth - measured;
    if (behind > trunc) return 0.0f;
    return (trunc - behind) / trunc;
  }
  return 1.0f;
}
inline float weight_inverse_square(float measured, float voxel_depth, float trunc) {
  if (voxel_depth <= 1e-2f) return 1.0f;
  if (voxel_depth - measured >= trunc) return 0.0f;
  return 1.0f / (voxel_depth * voxel_depth);
}
inline float weight_distance_penalty(float measured, float voxel_depth, float trunc) {
  const float d = measured - voxel_depth;
  if (std::fabs(d) >= trunc) return 0.1f;
  return 1.0f;
}
inline float weighting(int mode, float measured, float voxel_depth, float trunc) {
  switch (mode) {
    case NVBX_WEIGHT_CONSTANT:
      return 1.0f;
    case NVBX_WEIGHT_CONSTANT_DROPOFF:
      return 1.0f * weight_dropoff(measured, voxel_depth, trunc);
    case NVBX_WEIGHT_INVERSE_SQUARE:
      return weight_inverse_square(measured, voxel_depth, trunc);
    case NVBX_WEIGHT_INVERSE_SQUARE_DROPOFF:
      return weight_inverse_square(measured, voxel_depth, trunc) * weight_dropoff(measured, voxel_depth, trunc);
    case NVBX_WEIGHT_INVERSE_SQUARE_TSDF_DISTANCE_PENALTY:
      return weight_inverse_square(measured, voxel_depth, trunc) * weight_distance_penalty(measured, voxel_depth, trunc);
    case NVBX_WEIGHT_LINEAR_WITH_MAX:
      return voxel_depth > 1.0f ? 1.0f / voxel_depth : 1.0f;
  }
  return 0.0f;
}

// projectThreadVoxel
inline bool project_voxel(const Oracle& o, const I3& b, const I3& v, const Cam& cam, const Pose& T_C_L, float* u,
                          float* vv, float* depth) {
  const V3 pl = dev_voxel_center(o.block_size, b, v);
  const V3 pc = dev_xform(T_C_L, pl);
  if (!dev_project(cam, pc, u, vv)) return false;
  *depth = pc.z;
  const float max_depth = o.p.max_integration_distance_m;
  if (max_depth > 0.0f && *depth > max_depth) return false;
  return true;
}

void integrate_depth(Oracle& o, const float* depth, int rows, int cols, const uint8_t* mask, const Pose& T_L_C,
                     const Cam& cam) {
  o.cnt.depth_frames++;
  const std::vector<I3> blocks = blocks_in_view_raycast(o, depth, rows, cols, T_L_C, cam);
  o.last_tsdf_list = blocks;
  if (blocks.empty()) return;
  for (const I3& b : blocks) {
    if (!o.tsdf.count(b)) {
      o.tsdf[b];
      o.cnt.tsdf_blocks_allocated++;
    }
  }
  const Pose T_C_L = inverse(T_L_C);
  const float trunc = trunc_tsdf(o);
  std::vector<TsdfBlock*> blk_ptrs(blocks.size());
  for (size_t i = 0; i < blocks.size(); ++i) blk_ptrs[i] = &o.tsdf[blocks[i]];
  int64_t n_updated = 0;
  if (g_hooks.tsdf) {
    float Tm[16];
    pose_to_row_major(T_C_L, Tm);
    const float cm[4] = {cam.fu, cam.fv, cam.cu, cam.cv};
    std::vector<int> idx(blocks.size() * 3);
    std::vector<float> vox(blocks.size() * 1024);
    for (size_t i = 0; i < blocks.size(); ++i) {
      idx[3 * i] = blocks[i].x;
      idx[3 * i + 1] = blocks[i].y;
      idx[3 * i + 2] = blocks[i].z;
      for (int l = 0; l < 512; ++l) {
        vox[i * 1024 + 2 * l] = blk_ptrs[i]->d[l];
        vox[i * 1024 + 2 * l + 1] = blk_ptrs[i]->w[l];
      }
    }
    if (g_hooks.tsdf(Tm, cm, cols, rows, depth, mask, o.block_size, o.p.max_integration_distance_m, trunc,
                     o.p.max_weight, o.p.invalid_depth_decay_factor, o.p.weighting_mode, idx.data(), (int)blocks.size(),
                     vox.data()))
      std::fprintf(stderr, "orc: tsdf hook failed\n");
    for (size_t i = 0; i < blocks.size(); ++i)
      for (int l = 0; l < 512; ++l) {
        blk_ptrs[i]->d[l] = vox[i * 1024 + 2 * l];
        blk_ptrs[i]->w[l] = vox[i * 1024 + 2 * l + 1];
      }
  } else
  // blocks are independent: OpenMP over blocks does not change any result
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_updated)
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    const I3 b = blocks[bi];
    TsdfBlock& blk = *blk_ptrs[bi];
    for (int x = 0; x < 8; ++x)
      for (int y = 0; y < 8; ++y)
        for (int z = 0; z < 8; ++z) {
          float u, v, vd;
          if (!project_voxel(o, b, I3{x, y, z}, cam, T_C_L, &u, &v, &vd)) continue;
          // interpolate2DClosest<PixelNotNan> (interpolation_2d_impl.h:130-155)
          const int ui = (int)std::floor(u), vi = (int)std::floor(v);
          if (ui < 0 || vi < 0 || ui >= cols || vi >= rows) continue;
          const float meas = depth[(size_t)vi * cols + ui];
          if (std::isnan(meas)) continue;
          const bool active = (mask == nullptr) || mask[(size_t)vi * cols + ui];
          const int l = vlin(x, y, z);
          // UpdateTsdfVoxelFunctor
          if (meas <= 0.0f) {
            if (o.p.invalid_depth_decay_factor >= 0.0f) blk.w[l] *= o.p.invalid_depth_decay_factor;
            continue;
          }
          const float sdf = meas - vd;
          if (sdf < -trunc) continue;
          if (!active && sdf < trunc) continue;
          const float d_cur = blk.d[l];
          const float w_cur = blk.w[l];
          const float w_m = weighting(o.p.weighting_mode, meas, vd, trunc);
          // (sdf*w_m + d_cur*w_cur) / (w_m + w_cur); contracted: FMUL sdf*w_m, FFMA(d_cur, w_cur, .)
          float fused = (nvcc_model() ? std::fmaf(d_cur, w_cur, sdf * w_m) : (sdf * w_m + d_cur * w_cur)) / (w_m + w_cur);
          if (fused > 0.0f)
            fused = std::fmin(trunc, fused);
          else
            fused = std::fmax(-trunc, fused);
          const float w_new = std::fmin(w_m + w_cur, o.p.max_weight);
          blk.d[l] = fused;
          blk.w[l] = w_new;
          n_updated++;
        }
  }
  o.cnt.tsdf_voxels_updated += n_updated;
  o.cnt.tsdf_blocks_in_view += (int64_t)blocks.size();
  for (const I3& b : blocks) {  // mapper.cpp:406 -> blocks_to_update_tracker.cpp:32-60 (every consumer set)
    o.mesh_dirty.insert(b);
    o.cmesh_dirty.insert(b);
  }
}

// ------------------------------------------------------------------------------------------------
// a6: planes view + band filter.  view_calculator.cu:392-470, bounding_boxes_impl.h:28-52,
// projective_appearance_integrator.cu:374-477
// ------------------------------------------------------------------------------------------------
std::vector<I3> blocks_in_view_planes(Oracle& o, ViewCache& cache, const Pose& T_L_C, const Cam& cam,
                                      float max_distance) {
  if (o.p.cache_last_viewpoint) {
    if (const std::vector<I3>* hit = cache.get(T_L_C, cam)) return *hit;
  }
  Aabb aabb = view_aabb(cam, T_L_C, 1e-6f, max_distance);
  if (!apply_workspace_bounds(o.p, &aabb)) return {};
  const float bs = o.block_size;
  const I3 mn = block_index_from_position(bs, V3{aabb.mn[0], aabb.mn[1], aabb.mn[2]});
  const I3 mx = block_index_from_position(bs, V3{aabb.mx[0], aabb.mx[1], aabb.mx[2]});
  // Camera::getNormalizedViewport(10) camera.cpp:105-116
  const V3 vmin = ray_from_image_plane(cam, -10.0f, -10.0f);
  const V3 vmax = ray_from_image_plane(cam, (float)cam.width + 10.0f, (float)cam.height + 10.0f);
  const Pose T_C_L = inverse(T_L_C);
  std::vector<I3> out;
  for (int x = mn.x; x <= mx.x; ++x)
    for (int y = mn.y; y <= mx.y; ++y)
      for (int z = mn.z; z <= mx.z; ++z) {
        // getCenterPositionFromBlockIndex: block_size * (float(idx) + 0.5)
        const V3 c{bs * ((float)x + 0.5f), bs * ((float)y + 0.5f), bs * ((float)z + 0.5f)};
        const V3 r = rotate(T_C_L, c);
        const V3 pc{r.x + T_C_L.t[0], r.y + T_C_L.t[1], r.z + T_C_L.t[2]};
        if (pc.z > 1e-6f) {
          const float un = pc.x / pc.z, vn = pc.y / pc.z;
          if (vmin.x <= un && vmin.y <= vn && un <= vmax.x && vn <= vmax.y) out.push_back(I3{x, y, z});
        }
      }
  if (o.p.cache_last_viewpoint) cache.store(T_L_C, cam, out);
  return out;
}

// ------------------------------------------------------------------------------------------------
// a7: sphere tracing.  NB/src/rays/sphere_tracer.cu:26-131,191-236,421-480
// ------------------------------------------------------------------------------------------------
inline bool sphere_cast(const Oracle& o, const V3& origin, const V3& dir, float trunc, float* t_out,
                        int* steps_out) {
  const float eps = o.p.sphere_tracing_surface_epsilon_vox * o.voxel_size;
  int first = 0;  // 0 unknown, 1 positive, -1 negative
  float t = 0.0f;
  for (int i = 0; (i < o.p.sphere_tracing_max_steps) && (t < o.p.sphere_tracing_max_ray_length_m); ++i) {
    *steps_out = i + 1;
    const V3 p{dev_fma(t, dir.x, origin.x), dev_fma(t, dir.y, origin.y), dev_fma(t, dir.z, origin.z)};
    I3 b, v;
    dev_block_and_voxel_from_position(o.block_size, p, &b, &v);
    float step;
    auto it = o.tsdf.find(b);
    bool valid = false;
    float dist = 0.0f;
    if (it != o.tsdf.end()) {
      const int l = vlin(v.x, v.y, v.z);
      if (it->second.w[l] > 1e-4f) {
        valid = true;
        dist = it->second.d[l];
      }
    }
    if (!valid) {
      if (first == 0) {
        step = trunc;
      } else {
        *t_out = t;
        return false;
      }
    } else {
      if (first == 0) first = (dist >= 0.0f) ? 1 : -1;
      if (first == 1) {
        if (dist < eps) {
          *t_out = t + dist;
          return true;
        }
        step = dist;
      } else {
        if (dist > -eps) {
          *t_out = t - dist;
          return true;
        }
        step = -dist;
      }
    }
    t += step;
  }
  *t_out = t;
  return false;
}

void render_synthetic_depth(Oracle& o, const Cam& cam, const Pose& T_L_C, float trunc) {
  const int s = o.p.sphere_tracing_subsampling;
  const int rows = cam.height / s, cols = cam.width / s;
  o.synth.assign((size_t)rows * cols, 0.0f);
  o.synth_rows = rows;
  o.synth_cols = cols;
  const V3 origin{T_L_C.t[0], T_L_C.t[1], T_L_C.t[2]};
  int64_t steps_total = 0, steps_max = 0;
  if (g_hooks.trace) {
    float Tm[16];
    pose_to_row_major(T_L_C, Tm);
    const float cm[4] = {cam.fu, cam.fv, cam.cu, cam.cv};
    std::vector<int> idx;
    std::vector<float> vox;
    idx.reserve(o.tsdf.size() * 3);
    vox.reserve(o.tsdf.size() * 1024);
    for (const auto& kv : o.tsdf) {
      idx.push_back(kv.first.x);
      idx.push_back(kv.first.y);
      idx.push_back(kv.first.z);
      for (int l = 0; l < 512; ++l) {
        vox.push_back(kv.second.d[l]);
        vox.push_back(kv.second.w[l]);
      }
    }
    if (g_hooks.trace(Tm, cm, cam.width, cam.height, idx.data(), (int)o.tsdf.size(), vox.data(), trunc, o.block_size,
                      o.p.sphere_tracing_max_steps, o.p.sphere_tracing_max_ray_length_m,
                      o.p.sphere_tracing_surface_epsilon_vox * o.voxel_size, s, o.synth.data()))
      std::fprintf(stderr, "orc: trace hook failed\n");
    return;
  }
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : steps_total) reduction(max : steps_max)
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float pu = (float)(c * s) + 0.5f * (float)s * 1.0f;
      const float pv = (float)(r * s) + 0.5f * (float)s * 1.0f;
      const V3 ray = ray_from_image_plane(cam, pu, pv);
      // Eigen normalized(): z = squaredNorm; if (z > 0) v / sqrt(z)
      // squaredNorm: x*x + (y*y + z*z); contracted: fma(x, x, fma(y, y, z*z))
      const float sq = nvcc_model() ? std::fmaf(ray.x, ray.x, std::fmaf(ray.y, ray.y, ray.z * ray.z))
                                    : sum3(ray.x * ray.x, ray.y * ray.y, ray.z * ray.z);
      V3 dc = ray;
      if (sq > 0.0f) {
        const float n = std::sqrt(sq);
        dc = V3{ray.x / n, ray.y / n, ray.z / n};
      }
      const V3 dl = dev_rotate(T_L_C, dc);
      float t;
      int steps = 0;
      if (sphere_cast(o, origin, dl, trunc, &t, &steps))
        o.synth[(size_t)r * cols + c] = t * dc.z;
      else
        o.synth[(size_t)r * cols + c] = -1.0f;
      steps_total += steps;
      steps_max = std::max<int64_t>(steps_max, steps);
    }
  o.last_trace_steps = steps_total;
  o.last_trace_max_steps = steps_max;
}

// ------------------------------------------------------------------------------------------------
// a5/a8: feature integration.  projective_appearance_integrator.cu:72-169,267-353,
// projective_integrator_impl.cuh:156-214, interpolation_2d_impl.h:157-204
// ------------------------------------------------------------------------------------------------
void integrate_features(Oracle& o, const uint16_t* img, int rows, int cols, const uint8_t* mask, const Pose& T_L_C,
                        const Cam& cam) {
  o.cnt.feature_frames++;
  o.last_feat_list.clear();
  o.pixel_seen.clear();
  o.last_distinct_pixels = 0;
  o.last_n_upd = 0;
  const int C = o.C;
  const float trunc = o.p.appearance_truncation_distance_vox * o.voxel_size;
  std::vector<I3> cand =
      blocks_in_view_planes(o, o.planes_cache, T_L_C, cam, o.p.max_integration_distance_m + trunc);
  // reduceBlocksToThoseInTruncationBand
  std::vector<I3> band;
  for (const I3& b : cand) {
    auto it = o.tsdf.find(b);
    if (it == o.tsdf.end()) continue;
    o.cnt.feature_candidate_blocks++;
    bool in_band = false;
    for (int l = 0; l < 512 && !in_band; ++l)
      if (it->second.w[l] > 0.0f && std::fabs(it->second.d[l]) < trunc) in_band = true;
    if (in_band) band.push_back(b);
  }
  o.last_feat_list = band;
  if (band.empty()) return;
  for (const I3& b : band) {
    if (!o.feat.count(b)) {
      o.feat[b].assign((size_t)512 * (C + 1), 0);
      o.cnt.feature_blocks_allocated++;
    }
  }
  render_synthetic_depth(o, cam, T_L_C, trunc);
  const int sub = rows / o.synth_rows;  // projective_integrator_impl.cuh:424-425
  const Pose T_C_L = inverse(T_L_C);
  const float alpha = o.p.appearance_measurement_weight;
  std::vector<std::vector<uint16_t>*> fblk_ptrs(band.size());
  for (size_t i = 0; i < band.size(); ++i) fblk_ptrs[i] = &o.feat[band[i]];
  int64_t n_upd = 0;
  const bool fused_half = fused_half_on();
  if (g_hooks.feat) {
    float Tm[16];
    pose_to_row_major(T_C_L, Tm);
    const float cm[4] = {cam.fu, cam.fv, cam.cu, cam.cv};
    const size_t per = (size_t)512 * (C + 1);
    std::vector<int> idx(band.size() * 3);
    std::vector<uint16_t> vox(band.size() * per);
    for (size_t i = 0; i < band.size(); ++i) {
      idx[3 * i] = band[i].x;
      idx[3 * i + 1] = band[i].y;
      idx[3 * i + 2] = band[i].z;
      std::memcpy(vox.data() + i * per, fblk_ptrs[i]->data(), per * 2);
    }
    if (g_hooks.feat(Tm, cm, cols, rows, img, mask, o.synth.data(), sub, o.block_size, o.p.max_integration_distance_m,
                     trunc, o.p.max_weight, alpha, idx.data(), (int)band.size(), vox.data()))
      std::fprintf(stderr, "orc: feature hook failed\n");
    for (size_t i = 0; i < band.size(); ++i) std::memcpy(fblk_ptrs[i]->data(), vox.data() + i * per, per * 2);
  } else
#pragma omp parallel reduction(+ : n_upd)
  {
  std::vector<uint16_t> meas((size_t)C);
  std::unordered_set<uint64_t> seen_local;
#pragma omp for schedule(dynamic, 2)
  for (size_t bi = 0; bi < band.size(); ++bi) {
    const I3 b = band[bi];
    std::vector<uint16_t>& blk = *fblk_ptrs[bi];
    for (int x = 0; x < 8; ++x)
      for (int y = 0; y < 8; ++y)
        for (int z = 0; z < 8; ++z) {
          float u, v, vd;
          if (!project_voxel(o, b, I3{x, y, z}, cam, T_C_L, &u, &v, &vd)) continue;
          // synthetic depth, bilinear, no validity check
          const float ud = u / (float)sub, vdp = v / (float)sub;
          float surface;
          {
            const float uc = ud - 0.5f, vc = vdp - 0.5f;
            const int lx = (int)std::floor(uc), ly = (int)std::floor(vc);
            if (lx < 0 || ly < 0 || (lx + 1) > (o.synth_cols - 1) || (ly + 1) > (o.synth_rows - 1)) continue;
            const float f00 = o.synth[(size_t)ly * o.synth_cols + lx];
            const float f01 = o.synth[(size_t)(ly + 1) * o.synth_cols + lx];
            const float f10 = o.synth[(size_t)ly * o.synth_cols + lx + 1];
            const float f11 = o.synth[(size_t)(ly + 1) * o.synth_cols + lx + 1];
            surface = interp_float(uc - (float)lx, vc - (float)ly, f00, f01, f10, f11);
          }
          if (std::fabs(surface - vd) > trunc) continue;
          // feature image, bilinear in fp16
          const float uc = u - 0.5f, vc = v - 0.5f;
          const int lx = (int)std::floor(uc), ly = (int)std::floor(vc);
          if (lx < 0 || ly < 0 || (lx + 1) > (cols - 1) || (ly + 1) > (rows - 1)) continue;
          // mask lookup happens after the interpolation succeeded: isMasked(int(v), int(u))
          const bool active = (mask == nullptr) || mask[(size_t)((int)v) * cols + (int)u];
          if (!active) continue;
          const uint16_t hx = f2h(uc - (float)lx), hy = f2h(vc - (float)ly);
          const uint16_t* p00 = img + ((size_t)ly * cols + lx) * C;
          const uint16_t* p01 = img + ((size_t)(ly + 1) * cols + lx) * C;
          const uint16_t* p10 = img + ((size_t)ly * cols + lx + 1) * C;
          const uint16_t* p11 = img + ((size_t)(ly + 1) * cols + lx + 1) * C;
          for (int c = 0; c < C; ++c) meas[c] = interp_half(hx, hy, p00[c], p01[c], p10[c], p11[c]);
          seen_local.insert(((uint64_t)ly << 32) | (uint32_t)lx);
          seen_local.insert(((uint64_t)(ly + 1) << 32) | (uint32_t)lx);
          seen_local.insert(((uint64_t)ly << 32) | (uint32_t)(lx + 1));
          seen_local.insert(((uint64_t)(ly + 1) << 32) | (uint32_t)(lx + 1));
          uint16_t* vox = blk.data() + (size_t)vlin(x, y, z) * (C + 1);
          const float w_cur = h2f(vox[C]);
          if (w_cur == 0.0f) {
            for (int c = 0; c < C; ++c) vox[c] = meas[c];
          } else {
            float w1 = 1.0f - alpha, w2 = alpha;
            const float tot = w1 + w2;
            w1 /= tot;
            w2 /= tot;
            const uint16_t h1 = f2h(w1), h2 = f2h(w2);
            for (int c = 0; c < C; ++c) {
              if (!fused_half)
                vox[c] = hadd(hmul(vox[c], h1), hmul(meas[c], h2));
              else
                vox[c] = hfma(vox[c], h1, hmul(meas[c], h2));
            }
          }
          vox[C] = f2h(std::fmin(alpha + w_cur, o.p.max_weight));
          n_upd++;
        }
  }
#pragma omp critical
  o.pixel_seen.insert(seen_local.begin(), seen_local.end());
  }  // omp parallel
  o.cnt.feature_voxels_updated += n_upd;
  o.last_n_upd += n_upd;
  o.cnt.feature_band_blocks += (int64_t)band.size();
  o.last_distinct_pixels = (int64_t)o.pixel_seen.size();
  for (const I3& b : band) {  // mapper.cpp:462
    o.mesh_dirty.insert(b);
    o.cmesh_dirty.insert(b);
  }
}

// ------------------------------------------------------------------------------------------------
// N1: colour integration.  The same ProjectiveAppearanceIntegrator template instantiated for ColorLayer
// (projective_appearance_integrator.cu:72-169; kernel projective_integrator_impl.cuh:156-214), with
//   * interpolatePixels(Color) interpolation_2d_impl.h:50-60: per channel the fp32 bilinear formula on the
//     uint8 values, then static_cast<uint8_t>(std::round(.));
//   * weightedSum(uint8_t, float, uint8_t, float) projective_appearance_integrator.cu:277-284, called with
//     __float2half(weight) (:301-302), i.e. both blend weights rounded to binary16 and widened back;
//   * ColorVoxel::weight is a float, but the first-observation test is `__half2float(voxel_ptr->weight) == 0`
//     (:328), i.e. the float weight rounded to binary16.
// ------------------------------------------------------------------------------------------------
inline uint8_t interp_color_channel(float x, float y, uint8_t c00, uint8_t c01, uint8_t c10, uint8_t c11) {
  const float v = interp_float(x, y, (float)c00, (float)c01, (float)c10, (float)c11);
  return (uint8_t)std::round(v);
}
inline uint8_t blend_color_channel(uint8_t a, float wa, uint8_t b, float wb) {
  if (nvcc_model()) return (uint8_t)std::round(std::fmaf(wa, (float)a, wb * (float)b));   // FMUL + FFMA
  return (uint8_t)std::round((float)a * wa + (float)b * wb);
}

void integrate_color(Oracle& o, const uint8_t* img, int rows, int cols, const uint8_t* mask, const Pose& T_L_C,
                     const Cam& cam) {
  o.cnt.color_frames++;
  o.last_color_list.clear();
  const float trunc = o.p.appearance_truncation_distance_vox * o.voxel_size;
  std::vector<I3> cand =
      blocks_in_view_planes(o, o.planes_cache, T_L_C, cam, o.p.max_integration_distance_m + trunc);
  std::vector<I3> band;
  for (const I3& b : cand) {
    auto it = o.tsdf.find(b);
    if (it == o.tsdf.end()) continue;
    bool in_band = false;
    for (int l = 0; l < 512 && !in_band; ++l)
      if (it->second.w[l] > 0.0f && std::fabs(it->second.d[l]) < trunc) in_band = true;
    if (in_band) band.push_back(b);
  }
  o.last_color_list = band;
  if (band.empty()) return;
  for (const I3& b : band) {
    if (!o.color.count(b)) {
      o.color[b];
      o.cnt.color_blocks_allocated++;
    }
  }
  render_synthetic_depth(o, cam, T_L_C, trunc);
  const int sub = rows / o.synth_rows;
  const Pose T_C_L = inverse(T_L_C);
  const float alpha = o.p.appearance_measurement_weight;
  int64_t n_upd = 0;
  if (g_hooks.color) {
    float Tm[16];
    pose_to_row_major(T_C_L, Tm);
    const float cm[4] = {cam.fu, cam.fv, cam.cu, cam.cv};
    std::vector<int> idx(band.size() * 3);
    std::vector<uint8_t> rgb(band.size() * 1536);
    std::vector<float> wgt(band.size() * 512);
    for (size_t i = 0; i < band.size(); ++i) {
      idx[3 * i] = band[i].x;
      idx[3 * i + 1] = band[i].y;
      idx[3 * i + 2] = band[i].z;
      const ColorBlock& cb = o.color[band[i]];
      std::memcpy(rgb.data() + i * 1536, cb.c, 1536);
      std::memcpy(wgt.data() + i * 512, cb.w, 2048);
    }
    if (g_hooks.color(Tm, cm, cols, rows, img, mask, o.synth.data(), sub, o.block_size,
                      o.p.max_integration_distance_m, trunc, o.p.max_weight, alpha, idx.data(), (int)band.size(),
                      rgb.data(), wgt.data()))
      std::fprintf(stderr, "orc: colour hook failed\n");
    for (size_t i = 0; i < band.size(); ++i) {
      ColorBlock& cb = o.color[band[i]];
      std::memcpy(cb.c, rgb.data() + i * 1536, 1536);
      std::memcpy(cb.w, wgt.data() + i * 512, 2048);
    }
  } else
  for (const I3& b : band) {
    ColorBlock& blk = o.color[b];
    for (int x = 0; x < 8; ++x)
      for (int y = 0; y < 8; ++y)
        for (int z = 0; z < 8; ++z) {
          float u, v, vd;
          if (!project_voxel(o, b, I3{x, y, z}, cam, T_C_L, &u, &v, &vd)) continue;
          const float ud = u / (float)sub, vdp = v / (float)sub;
          float surface;
          {
            const float uc = ud - 0.5f, vc = vdp - 0.5f;
            const int lx = (int)std::floor(uc), ly = (int)std::floor(vc);
            if (lx < 0 || ly < 0 || (lx + 1) > (o.synth_cols - 1) || (ly + 1) > (o.synth_rows - 1)) continue;
            const float f00 = o.synth[(size_t)ly * o.synth_cols + lx];
            const float f01 = o.synth[(size_t)(ly + 1) * o.synth_cols + lx];
            const float f10 = o.synth[(size_t)ly * o.synth_cols + lx + 1];
            const float f11 = o.synth[(size_t)(ly + 1) * o.synth_cols + lx + 1];
            surface = interp_float(uc - (float)lx, vc - (float)ly, f00, f01, f10, f11);
          }
          if (std::fabs(surface - vd) > trunc) continue;
          const float uc = u - 0.5f, vc = v - 0.5f;
          const int lx = (int)std::floor(uc), ly = (int)std::floor(vc);
          if (lx < 0 || ly < 0 || (lx + 1) > (cols - 1) || (ly + 1) > (rows - 1)) continue;
          const bool active = (mask == nullptr) || mask[(size_t)((int)v) * cols + (int)u];
          if (!active) continue;
          const float ox = uc - (float)lx, oy = vc - (float)ly;
          const uint8_t* p00 = img + ((size_t)ly * cols + lx) * 3;
          const uint8_t* p01 = img + ((size_t)(ly + 1) * cols + lx) * 3;
          const uint8_t* p10 = img + ((size_t)ly * cols + lx + 1) * 3;
          const uint8_t* p11 = img + ((size_t)(ly + 1) * cols + lx + 1) * 3;
          uint8_t meas[3];
          for (int c = 0; c < 3; ++c) meas[c] = interp_color_channel(ox, oy, p00[c], p01[c], p10[c], p11[c]);
          const int l = vlin(x, y, z);
          const float w_cur = blk.w[l];
          if (h2f(f2h(w_cur)) == 0.0f) {
            for (int c = 0; c < 3; ++c) blk.c[l][c] = meas[c];
          } else {
            float w1 = 1.0f - alpha, w2 = alpha;
            const float tot = w1 + w2;
            w1 /= tot;
            w2 /= tot;
            const float h1 = h2f(f2h(w1)), h2 = h2f(f2h(w2));
            for (int c = 0; c < 3; ++c) blk.c[l][c] = blend_color_channel(blk.c[l][c], h1, meas[c], h2);
          }
          blk.w[l] = std::fmin(alpha + w_cur, o.p.max_weight);
          n_upd++;
        }
  }
  o.cnt.color_voxels_updated += n_upd;
  o.cnt.color_band_blocks += (int64_t)band.size();
  for (const I3& b : band) {  // mapper.cpp:448
    o.mesh_dirty.insert(b);
    o.cmesh_dirty.insert(b);
  }
}

// ------------------------------------------------------------------------------------------------
// a9: decay.  NB/src/integrators/tsdf_decay_integrator.cu:58-112, decayer_impl.cuh:83-274,
// mapper.cpp:466-495,761-849
// ------------------------------------------------------------------------------------------------
void decay(Oracle& o) {
  for (auto& kv : o.tsdf) {
    o.mesh_dirty.insert(kv.first);
    o.cmesh_dirty.insert(kv.first);
  }
  const float thr = o.p.tsdf_decayed_weight_threshold;
  const float free_d = o.p.tsdf_decayed_free_distance_vox * o.voxel_size;
  std::vector<I3> removed;
  for (auto& kv : o.tsdf) {
    TsdfBlock& b = kv.second;
    bool all = true;
    for (int l = 0; l < 512; ++l) {
      float w = b.w[l];
      if (!(w < (thr - 1e-6f))) {
        w *= o.p.tsdf_decay_factor;
        w = std::fmax(w, thr);
        b.w[l] = w;
        if (o.p.tsdf_set_free_distance_on_decayed && (b.w[l] < (thr + 1e-6f))) b.d[l] = free_d;
      }
      if (!(b.w[l] < (thr + 1e-6f))) all = false;
    }
    if (all && o.p.deallocate_decayed_blocks) removed.push_back(kv.first);
  }
  for (const I3& b : removed) {
    o.tsdf.erase(b);
    o.feat.erase(b);
    o.color.erase(b);
    o.mesh.erase(b);
    o.cmesh.erase(b);
    o.mesh_dirty.erase(b);
    o.cmesh_dirty.erase(b);
    o.cnt.blocks_deallocated++;
  }
}

// ------------------------------------------------------------------------------------------------
// a10/a11: feature mesh.  NB/src/mesh/mesh_integrator.cu:64-103,308-822, marching_cubes_impl.h:6-62,
// cuda/marching_cubes_impl.cuh:10-72, mesh_integrator_appearance.cu:97-146,290-340
//
// Canonical order (the reference's is non-deterministic: atomicAdd slot reservation, hash-map block
// order): voxels in ascending memory order vx*64+vy*8+vz, triangles in table order; the weld keeps
// the first vertex (in that order) of every key and orders the survivors by ascending key; blocks are
// serialised in ascending (x, y, z).
// ------------------------------------------------------------------------------------------------
inline uint64_t weld_key(const V3& v) {  // Index3DHash(Index3D(v * 1000)), hash.h:32-40
  const int ix = (int)(v.x * 1000), iy = (int)(v.y * 1000), iz = (int)(v.z * 1000);
  return (uint64_t)(int64_t)ix + (uint64_t)(int64_t)iy * 17191ull + (uint64_t)(int64_t)iz * (17191ull * 17191ull);
}
inline V3 interp_vertex(const V3& a, const V3& b, float s1, float s2) {  // marching_cubes_impl.h:28-41
  const float diff = s1 - s2;
  if (std::fabs(diff) >= 1e-4f) {
    const float t = s1 / diff;
    return V3{dev_fma(t, b.x - a.x, a.x), dev_fma(t, b.y - a.y, a.y), dev_fma(t, b.z - a.z, a.z)};
  }
  return V3{0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z)};
}

void mesh_block(Oracle& o, const I3& bi, MeshBlock* out) {
  out->verts.clear();
  out->tris.clear();
  out->feats.clear();
  const TsdfBlock* nb[8];
  for (int j = 0; j < 8; ++j) {
    const I3 d{(j & 4) >> 2, (j & 2) >> 1, j & 1};  // directionFromNeighborIndex
    auto it = o.tsdf.find(I3{bi.x + d.x, bi.y + d.y, bi.z + d.z});
    nb[j] = it == o.tsdf.end() ? nullptr : &it->second;
  }
  const float vs = o.voxel_size;
  const V3 origin{o.block_size * (float)bi.x, o.block_size * (float)bi.y, o.block_size * (float)bi.z};
  const float min_w = o.p.mesh_min_weight;
  std::vector<V3> verts;
  for (int x = 0; x < 8; ++x)
    for (int y = 0; y < 8; ++y)
      for (int z = 0; z < 8; ++z) {
        float sdf[8];
        V3 pos[8];
        bool skip = false;
        for (int i = 0; i < 8 && !skip; ++i) {
          int c[3] = {x + kMcCornerOffsets[i][0], y + kMcCornerOffsets[i][1], z + kMcCornerOffsets[i][2]};
          int off[3] = {0, 0, 0};
          for (int j = 0; j < 3; ++j)
            if (c[j] >= 8) {
              c[j] -= 8;
              off[j] = 1;
            }
          const TsdfBlock* blk = nb[(off[0] << 2) | (off[1] << 1) | off[2]];
          if (blk == nullptr) {
            skip = true;
            break;
          }
          const int l = vlin(c[0], c[1], c[2]);
          if (blk->w[l] < min_w) {
            skip = true;
            break;
          }
          sdf[i] = blk->d[l];
          // block_position + voxel_size * (corner + 0.5 + 8*block_offset)   (mesh_integrator.cu:421-424)
          pos[i].x = dev_fma(vs, ((float)c[0] + 0.5f) + (float)(8 * off[0]), origin.x);
          pos[i].y = dev_fma(vs, ((float)c[1] + 0.5f) + (float)(8 * off[1]), origin.y);
          pos[i].z = dev_fma(vs, ((float)c[2] + 0.5f) + (float)(8 * off[2]), origin.z);
        }
        if (skip) continue;
        int cfg = 0;
        for (int i = 0; i < 8; ++i)
          if (sdf[i] < 0) cfg |= 1 << i;
        const int nv = kMcNumVerts[cfg];
        if (nv == 0) continue;
        V3 edge[12];
        for (int e = 0; e < 12; ++e) {
          const int a = kMcEdgePairs[e][0], b2 = kMcEdgePairs[e][1];
          if ((sdf[a] < 0 && sdf[b2] >= 0) || (sdf[a] >= 0 && sdf[b2] < 0)) edge[e] = interp_vertex(pos[a], pos[b2], sdf[a], sdf[b2]);
        }
        const int8_t* row = kMcTriTable[cfg];
        for (int t = 0; t < nv; t += 3) {  // calculateVertices writes col+2, col+1, col
          verts.push_back(edge[row[t + 2]]);
          verts.push_back(edge[row[t + 1]]);
          verts.push_back(edge[row[t]]);
        }
      }
  const int n = (int)verts.size();
  if (n == 0) return;
  out->tris.resize(n);
  if (o.p.mesh_weld_vertices && n < 128 * 20) {  // weldVerticesCubKernel<128,20>: skipped when n >= 2560
    std::vector<std::pair<uint64_t, int>> keyed(n);
    for (int i = 0; i < n; ++i) keyed[i] = {weld_key(verts[i]), i};
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a.first < b.first; });
    int unique = 0;
    for (int i = 0; i < n; ++i) {
      if (i == 0 || keyed[i].first != keyed[i - 1].first) {
        out->verts.push_back(verts[keyed[i].second]);
        unique++;
      }
      out->tris[keyed[i].second] = unique - 1;
    }
  } else {
    out->verts = verts;
    for (int i = 0; i < n; ++i) out->tris[i] = i;
  }
}

void paint_block(Oracle& o, const I3& bi, MeshBlock* mb) {
  const int C = o.C;
  mb->feats.assign(mb->verts.size() * (size_t)C, 0);
  auto it = o.feat.find(bi);
  if (it == o.feat.end()) return;  // updateAppearanceBlocksConstant with the default (zero) feature
  const V3 origin{o.block_size * (float)bi.x, o.block_size * (float)bi.y, o.block_size * (float)bi.z};
  // voxel_size = mesh_layer->block_size() / kVoxelsPerSide   (mesh_integrator_appearance.cu:268-269)
  const float vs = o.block_size / 8;
  for (size_t i = 0; i < mb->verts.size(); ++i) {
    const V3& v = mb->verts[i];
    // p_L_V - block_size * float(idx): mul.f32 + sub.f32 in the PTX, one FFMA in the SASS
    const float bs = o.block_size;
    int ix = (int)((nvcc_model() ? std::fmaf(-bs, (float)bi.x, v.x) : v.x - origin.x) / vs);
    int iy = (int)((nvcc_model() ? std::fmaf(-bs, (float)bi.y, v.y) : v.y - origin.y) / vs);
    int iz = (int)((nvcc_model() ? std::fmaf(-bs, (float)bi.z, v.z) : v.z - origin.z) / vs);
    ix = std::max(std::min(ix, 7), 0);
    iy = std::max(std::min(iy, 7), 0);
    iz = std::max(std::min(iz, 7), 0);
    const uint16_t* vox = it->second.data() + (size_t)vlin(ix, iy, iz) * (C + 1);
    std::memcpy(mb->feats.data() + i * (size_t)C, vox, (size_t)C * 2);
  }
}

// AppearanceGetter<ColorVoxel> (mesh_integrator_appearance.cu:43-54): closest voxel's colour, Gray without a
// colour block
void paint_block_color(Oracle& o, const I3& bi, MeshBlock* mb) {
  mb->colors.assign(mb->verts.size() * 3, 127);
  auto it = o.color.find(bi);
  if (it == o.color.end()) return;
  const V3 origin{o.block_size * (float)bi.x, o.block_size * (float)bi.y, o.block_size * (float)bi.z};
  const float vs = o.block_size / 8;
  for (size_t i = 0; i < mb->verts.size(); ++i) {
    const V3& v = mb->verts[i];
    // p_L_V - block_size * float(idx): mul.f32 + sub.f32 in the PTX, one FFMA in the SASS
    const float bs = o.block_size;
    int ix = (int)((nvcc_model() ? std::fmaf(-bs, (float)bi.x, v.x) : v.x - origin.x) / vs);
    int iy = (int)((nvcc_model() ? std::fmaf(-bs, (float)bi.y, v.y) : v.y - origin.y) / vs);
    int iz = (int)((nvcc_model() ? std::fmaf(-bs, (float)bi.z, v.z) : v.z - origin.z) / vs);
    ix = std::max(std::min(ix, 7), 0);
    iy = std::max(std::min(iy, 7), 0);
    iz = std::max(std::min(iz, 7), 0);
    std::memcpy(mb->colors.data() + i * 3, it->second.c[vlin(ix, iy, iz)], 3);
  }
}

// Mapper::updateColorMesh (mapper.cpp:616-619): the same updateMeshTemplate over the colour mesh layer and
// the colour "to update" set.
void update_color_mesh(Oracle& o) {
  std::vector<I3> blocks;
  for (const I3& b : o.cmesh_dirty)
    if (o.tsdf.count(b)) blocks.push_back(b);
  const float cutoff = o.p.mesh_cutoff_distance_vox * o.voxel_size;
  for (const I3& b : blocks) {
    auto mit = o.cmesh.find(b);
    if (mit != o.cmesh.end()) {
      mit->second.verts.clear();
      mit->second.tris.clear();
      mit->second.colors.clear();
    }
    const TsdfBlock& tb = o.tsdf[b];
    bool meshable = false;
    for (int l = 0; l < 512 && !meshable; ++l)
      if (std::fabs(tb.d[l]) <= cutoff && tb.w[l] >= o.p.mesh_min_weight) meshable = true;
    if (!meshable) continue;
    MeshBlock mb;
    mesh_block(o, b, &mb);
    if (mb.tris.empty()) continue;
    o.cmesh[b] = mb;
  }
  for (const I3& b : blocks) {
    auto mit = o.cmesh.find(b);
    if (mit == o.cmesh.end()) continue;
    paint_block_color(o, b, &mit->second);
  }
  o.cmesh_dirty.clear();
}

void update_feature_mesh(Oracle& o) {
  // getIndicesInLayer: only blocks still allocated in the TSDF layer
  std::vector<I3> blocks;
  for (const I3& b : o.mesh_dirty)
    if (o.tsdf.count(b)) blocks.push_back(b);
  const float cutoff = o.p.mesh_cutoff_distance_vox * o.voxel_size;
  for (const I3& b : blocks) {
    auto mit = o.mesh.find(b);
    if (mit != o.mesh.end()) {  // "clear all blocks if they exist"
      mit->second.verts.clear();
      mit->second.tris.clear();
      mit->second.feats.clear();
    }
    const TsdfBlock& tb = o.tsdf[b];
    bool meshable = false;  // isBlockMeshableKernel
    for (int l = 0; l < 512 && !meshable; ++l)
      if (std::fabs(tb.d[l]) <= cutoff && tb.w[l] >= o.p.mesh_min_weight) meshable = true;
    if (!meshable) continue;
    MeshBlock mb;
    mesh_block(o, b, &mb);
    o.cnt.mesh_blocks_remeshed++;
    if (mb.tris.empty()) continue;  // num_vertices == 0: no mesh block is allocated
    o.mesh[b] = mb;
  }
  // updateAppearance over the requested blocks that have a mesh block
  for (const I3& b : blocks) {
    auto mit = o.mesh.find(b);
    if (mit == o.mesh.end()) continue;
    paint_block(o, b, &mit->second);
  }
  o.mesh_dirty.clear();
  int64_t nv = 0;
  for (auto& kv : o.mesh) nv += (int64_t)kv.second.verts.size();
  o.cnt.mesh_vertices = nv;
}

}  // namespace

// ================================================================================================
// C interface (ctypes)
// ================================================================================================
extern "C" {

void orc_default_params(nvbx_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->max_integration_distance_m = 7.0f;
  p->truncation_distance_vox = 4.0f;
  p->weighting_mode = NVBX_WEIGHT_INVERSE_SQUARE;
  p->max_weight = 5.0f;
  p->invalid_depth_decay_factor = -1.0f;
  p->appearance_measurement_weight = 0.8f;
  p->appearance_truncation_distance_vox = 4.0f;
  p->sphere_tracing_subsampling = 4;
  p->sphere_tracing_max_ray_length_m = 7.0f;
  p->sphere_tracing_max_steps = 100;
  p->sphere_tracing_surface_epsilon_vox = 0.1f;
  p->tsdf_decay_factor = 0.95f;
  p->tsdf_decayed_weight_threshold = 1e-3f;
  p->tsdf_set_free_distance_on_decayed = 0;
  p->tsdf_decayed_free_distance_vox = 4.0f;
  p->deallocate_decayed_blocks = 1;
  p->raycast_subsampling_factor = 4;
  p->workspace_bounds_type = NVBX_WORKSPACE_UNBOUNDED;
  p->workspace_min[0] = 0.0f;
  p->workspace_min[1] = 2.0f;
  p->workspace_min[2] = 0.0f;
  p->workspace_max[0] = 0.0f;
  p->workspace_max[1] = 2.0f;
  p->workspace_max[2] = 1.0f;
  p->cache_last_viewpoint = 1;
  p->mesh_min_weight = 1e-4f;
  p->mesh_weld_vertices = 1;
  p->mesh_cutoff_distance_vox = 5.0f;
  p->num_preallocated_blocks = 2048;
  p->expansion_factor = 2.0f;
  p->strict_blend = 0;
}

void* orc_create(float voxel_size, int C, const nvbx_params* p) {
  Oracle* o = new Oracle();
  o->voxel_size = voxel_size;
  o->block_size = voxel_size * 8;  // voxelSizeToBlockSize
  o->C = C;
  o->p = *p;
  std::memset(&o->cnt, 0, sizeof(o->cnt));
  return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }
void orc_set_fused_half(int v) { g_fused_half_override = v < 0 ? -1 : (v != 0); }
void orc_set_fp_model(int m) { g_fp_model = m; }
void orc_set_ref_hooks(orc_hook_raycast_t raycast, orc_hook_tsdf_t tsdf, orc_hook_trace_t trace, orc_hook_feat_t feat,
                       orc_hook_color_t color) {
  g_hooks.raycast = raycast;
  g_hooks.tsdf = tsdf;
  g_hooks.trace = trace;
  g_hooks.feat = feat;
  g_hooks.color = color;
}
int orc_get_fp_model(void) { return g_fp_model; }
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

static Cam make_cam(float fx, float fy, float cx, float cy, int H, int W) { return Cam{fx, fy, cx, cy, W, H}; }

void orc_integrate_depth(void* h, const float* depth, int H, int W, const uint8_t* mask, const float* T, float fx,
                         float fy, float cx, float cy) {
  integrate_depth(*(Oracle*)h, depth, H, W, mask, pose_from_row_major(T), make_cam(fx, fy, cx, cy, H, W));
}
void orc_integrate_features(void* h, const uint16_t* feat, int H, int W, const uint8_t* mask, const float* T,
                            float fx, float fy, float cx, float cy) {
  integrate_features(*(Oracle*)h, feat, H, W, mask, pose_from_row_major(T), make_cam(fx, fy, cx, cy, H, W));
}
void orc_integrate_color(void* h, const uint8_t* rgb, int H, int W, const uint8_t* mask, const float* T, float fx,
                         float fy, float cx, float cy) {
  integrate_color(*(Oracle*)h, rgb, H, W, mask, pose_from_row_major(T), make_cam(fx, fy, cx, cy, H, W));
}
void orc_decay(void* h) { decay(*(Oracle*)h); }
void orc_clear(void* h) {  // py_mapper.cu:286-306 -- layers only; caches and tracker survive
  Oracle& o = *(Oracle*)h;
  o.tsdf.clear();
  o.feat.clear();
  o.color.clear();
  o.mesh.clear();
  o.cmesh.clear();
}
void orc_update_feature_mesh(void* h) { update_feature_mesh(*(Oracle*)h); }
void orc_update_color_mesh(void* h) { update_color_mesh(*(Oracle*)h); }
void orc_color_mesh_sizes(void* h, int64_t* n_verts, int64_t* n_tri_idx) {
  Oracle& o = *(Oracle*)h;
  int64_t nv = 0, nt = 0;
  for (auto& kv : o.cmesh) {
    nv += (int64_t)kv.second.verts.size();
    nt += (int64_t)kv.second.tris.size();
  }
  *n_verts = nv;
  *n_tri_idx = nt;
}
void orc_color_mesh_copy(void* h, float* verts, uint8_t* colors, int32_t* tris) {
  Oracle& o = *(Oracle*)h;
  int64_t vo = 0, to = 0;
  for (auto& kv : o.cmesh) {
    const MeshBlock& mb = kv.second;
    for (size_t i = 0; i < mb.verts.size(); ++i) {
      verts[(vo + (int64_t)i) * 3 + 0] = mb.verts[i].x;
      verts[(vo + (int64_t)i) * 3 + 1] = mb.verts[i].y;
      verts[(vo + (int64_t)i) * 3 + 2] = mb.verts[i].z;
    }
    if (colors && !mb.colors.empty()) std::memcpy(colors + vo * 3, mb.colors.data(), mb.colors.size());
    for (size_t i = 0; i < mb.tris.size(); ++i) tris[to + (int64_t)i] = mb.tris[i] + (int32_t)vo;
    vo += (int64_t)mb.verts.size();
    to += (int64_t)mb.tris.size();
  }
}
// colour blocks in block-index order: out_rgb uint8[n*512*3], out_w float[n*512]
void orc_get_all_color_blocks(void* h, uint8_t* out_rgb, float* out_w) {
  Oracle& o = *(Oracle*)h;
  for (auto& kv : o.color) {
    std::memcpy(out_rgb, kv.second.c, 512 * 3);
    std::memcpy(out_w, kv.second.w, 512 * 4);
    out_rgb += 512 * 3;
    out_w += 512;
  }
}

void orc_mesh_sizes(void* h, int64_t* n_verts, int64_t* n_tri_idx) {
  Oracle& o = *(Oracle*)h;
  int64_t nv = 0, nt = 0;
  for (auto& kv : o.mesh) {
    nv += (int64_t)kv.second.verts.size();
    nt += (int64_t)kv.second.tris.size();
  }
  *n_verts = nv;
  *n_tri_idx = nt;
}
// blocks in ascending (x,y,z); triangles made global by adding the block's vertex offset (py_mesh.cpp:33-52)
void orc_mesh_copy(void* h, float* verts, uint16_t* feats, int32_t* tris, int32_t* vertex_block_xyz) {
  Oracle& o = *(Oracle*)h;
  int64_t vo = 0, to = 0;
  for (auto& kv : o.mesh) {
    const MeshBlock& mb = kv.second;
    for (size_t i = 0; i < mb.verts.size(); ++i) {
      verts[(vo + (int64_t)i) * 3 + 0] = mb.verts[i].x;
      verts[(vo + (int64_t)i) * 3 + 1] = mb.verts[i].y;
      verts[(vo + (int64_t)i) * 3 + 2] = mb.verts[i].z;
      if (vertex_block_xyz) {
        vertex_block_xyz[(vo + (int64_t)i) * 3 + 0] = kv.first.x;
        vertex_block_xyz[(vo + (int64_t)i) * 3 + 1] = kv.first.y;
        vertex_block_xyz[(vo + (int64_t)i) * 3 + 2] = kv.first.z;
      }
    }
    if (feats && !mb.feats.empty()) std::memcpy(feats + vo * o.C, mb.feats.data(), mb.feats.size() * 2);
    for (size_t i = 0; i < mb.tris.size(); ++i) tris[to + (int64_t)i] = mb.tris[i] + (int32_t)vo;
    vo += (int64_t)mb.verts.size();
    to += (int64_t)mb.tris.size();
  }
}

int64_t orc_num_blocks(void* h, int layer) {
  Oracle& o = *(Oracle*)h;
  if (layer == NVBX_LAYER_COLOR) return (int64_t)o.color.size();
  return layer == NVBX_LAYER_TSDF ? (int64_t)o.tsdf.size() : (int64_t)o.feat.size();
}
void orc_block_indices(void* h, int layer, int32_t* out) {
  Oracle& o = *(Oracle*)h;
  int64_t i = 0;
  if (layer == NVBX_LAYER_TSDF) {
    for (auto& kv : o.tsdf) {
      out[i * 3] = kv.first.x;
      out[i * 3 + 1] = kv.first.y;
      out[i * 3 + 2] = kv.first.z;
      ++i;
    }
  } else if (layer == NVBX_LAYER_COLOR) {
    for (auto& kv : o.color) {
      out[i * 3] = kv.first.x;
      out[i * 3 + 1] = kv.first.y;
      out[i * 3 + 2] = kv.first.z;
      ++i;
    }
  } else {
    for (auto& kv : o.feat) {
      out[i * 3] = kv.first.x;
      out[i * 3 + 1] = kv.first.y;
      out[i * 3 + 2] = kv.first.z;
      ++i;
    }
  }
}
// TSDF: out float[512*2] interleaved (distance, weight) in voxels[x][y][z] order; feature: uint16[512*(C+1)]
int orc_get_block(void* h, int layer, int x, int y, int z, void* out) {
  Oracle& o = *(Oracle*)h;
  const I3 b{x, y, z};
  if (layer == NVBX_LAYER_TSDF) {
    auto it = o.tsdf.find(b);
    if (it == o.tsdf.end()) return 0;
    float* f = (float*)out;
    for (int l = 0; l < 512; ++l) {
      f[2 * l] = it->second.d[l];
      f[2 * l + 1] = it->second.w[l];
    }
    return 1;
  }
  auto it = o.feat.find(b);
  if (it == o.feat.end()) return 0;
  std::memcpy(out, it->second.data(), it->second.size() * 2);
  return 1;
}
// bulk export in block-index order (for fast parity checks)
void orc_get_all_blocks(void* h, int layer, void* out) {
  Oracle& o = *(Oracle*)h;
  if (layer == NVBX_LAYER_TSDF) {
    float* f = (float*)out;
    for (auto& kv : o.tsdf)
      for (int l = 0; l < 512; ++l) {
        *f++ = kv.second.d[l];
        *f++ = kv.second.w[l];
      }
  } else {
    uint16_t* f = (uint16_t*)out;
    for (auto& kv : o.feat) {
      std::memcpy(f, kv.second.data(), kv.second.size() * 2);
      f += kv.second.size();
    }
  }
}
void orc_set_tsdf_block(void* h, int x, int y, int z, const float* in) {
  Oracle& o = *(Oracle*)h;
  TsdfBlock& b = o.tsdf[I3{x, y, z}];
  for (int l = 0; l < 512; ++l) {
    b.d[l] = in[2 * l];
    b.w[l] = in[2 * l + 1];
  }
}
// feature block in the oracle's layout ([8][8][8][C+1] halves); allocates the block when absent (test helper)
void orc_set_feature_block(void* h, int x, int y, int z, const uint16_t* in) {
  Oracle& o = *(Oracle*)h;
  std::vector<uint16_t>& blk = o.feat[I3{x, y, z}];
  blk.assign(in, in + (size_t)512 * (o.C + 1));
}
void orc_mark_all_dirty(void* h) {
  Oracle& o = *(Oracle*)h;
  for (auto& kv : o.tsdf) {
    o.mesh_dirty.insert(kv.first);
    o.cmesh_dirty.insert(kv.first);
  }
}
// queryTSDFKernel / queryFeatureKernel, NT/cpp/src/sdf_query.cu:206-270
void orc_query_tsdf(void* h, const float* xyz, int64_t n, float* out) {
  Oracle& o = *(Oracle*)h;
  for (int64_t i = 0; i < n; ++i) {
    I3 b, v;
    dev_block_and_voxel_from_position(o.block_size, V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, &b, &v);
    auto it = o.tsdf.find(b);
    if (it == o.tsdf.end()) continue;
    const int l = vlin(v.x, v.y, v.z);
    out[2 * i] = it->second.d[l];
    out[2 * i + 1] = it->second.w[l];
  }
}
void orc_query_features(void* h, const float* xyz, int64_t n, uint16_t* out) {
  Oracle& o = *(Oracle*)h;
  for (int64_t i = 0; i < n; ++i) {
    I3 b, v;
    dev_block_and_voxel_from_position(o.block_size, V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, &b, &v);
    auto it = o.feat.find(b);
    if (it == o.feat.end()) continue;
    const uint16_t* vox = it->second.data() + (size_t)vlin(v.x, v.y, v.z) * (o.C + 1);
    // weight goes through float: output = half(float(weight)) -- identity
    std::memcpy(out + i * (o.C + 1), vox, (size_t)(o.C + 1) * 2);
  }
}
void orc_get_counters(void* h, nvbx_counters* out) { *out = ((Oracle*)h)->cnt; }
void orc_reset_counters(void* h) { std::memset(&((Oracle*)h)->cnt, 0, sizeof(nvbx_counters)); }
int64_t orc_last_distinct_pixels(void* h) { return ((Oracle*)h)->last_distinct_pixels; }
int64_t orc_last_feature_voxels(void* h) { return ((Oracle*)h)->last_n_upd; }
int64_t orc_last_trace_steps(void* h) { return ((Oracle*)h)->last_trace_steps; }
int64_t orc_last_trace_max_steps(void* h) { return ((Oracle*)h)->last_trace_max_steps; }
int64_t orc_last_block_list(void* h, int which, int32_t* out, int64_t cap) {
  Oracle& o = *(Oracle*)h;
  const std::vector<I3>& l = which == 0 ? o.last_tsdf_list : (which == 1 ? o.last_feat_list : o.last_color_list);
  if (out)
    for (int64_t i = 0; i < (int64_t)l.size() && i < cap; ++i) {
      out[3 * i] = l[i].x;
      out[3 * i + 1] = l[i].y;
      out[3 * i + 2] = l[i].z;
    }
  return (int64_t)l.size();
}
void orc_last_synthetic_depth(void* h, float* out, int* rows, int* cols) {
  Oracle& o = *(Oracle*)h;
  *rows = o.synth_rows;
  *cols = o.synth_cols;
  if (out) std::memcpy(out, o.synth.data(), o.synth.size() * 4);
}

// ---- unit hooks used to pin the restatement against the reference's own unit tests ----------------
// ---- N4: the extractor's up-sampling (third-party algorithm: PyTorch, pinned by mindmap's environment; torch
// 2.11.0+cu128 in this image).  Restates aten/src/ATen/native/cuda/UpSampleBilinear2d.cu
// (upsample_bilinear2d_out_frame / upsample_bilinear2d_nhwc_out_frame) + UpSample.cuh:114-130
// (area_pixel_compute_source_index) with the multiply-add contraction of the sm_100 build (fmaf below = one
// rounding), then feature_extraction.py:190-210 (HWC, zero-pad to C) and nvblox_mapping_helpers.py:256
// (.to(float16)).  low: [lh][lw][lc] fp32 (values already exact in the source dtype); mode: 0 NCHW-kernel order,
// 1 NHWC-f32-kernel order, 2 NCHW order with the bf16 store rounding.  out: [H][W][C] binary16.
static inline float bf16_round(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return v;  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  std::memcpy(&v, &u, 4);
  return v;
}
void orc_upsample_bilinear(const float* low, int lh, int lw, int lc, int C, int H, int W, int mode, uint16_t* out) {
  const float rh = (float)lh / (float)H, rw = (float)lw / (float)W;
  struct Axis {
    int i0, i1;
    float l0, l1;
  };
  auto axis = [](float scale, int dst, int n) {
    Axis a;
    float s = std::fmaf((float)dst + 0.5f, scale, -0.5f);
    s = (s >= 0.0f) ? s : 0.0f;
    a.i0 = (int)s;
    a.i1 = a.i0 + ((a.i0 < n - 1) ? 1 : 0);
    a.l1 = s - (float)a.i0;
    a.l0 = 1.0f - a.l1;
    return a;
  };
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y) {
    const Axis ya = axis(rh, y, lh);
    for (int x = 0; x < W; ++x) {
      const Axis xa = axis(rw, x, lw);
      const float* pa = low + ((size_t)ya.i0 * lw + xa.i0) * lc;
      const float* pb = low + ((size_t)ya.i0 * lw + xa.i1) * lc;
      const float* pc = low + ((size_t)ya.i1 * lw + xa.i0) * lc;
      const float* pd = low + ((size_t)ya.i1 * lw + xa.i1) * lc;
      uint16_t* o = out + ((size_t)y * W + x) * C;
      for (int c = 0; c < C; ++c) {
        if (c >= lc) {
          o[c] = 0;
          continue;
        }
        const float top = mode == 1 ? std::fmaf(xa.l1, pb[c], xa.l0 * pa[c]) : std::fmaf(xa.l0, pa[c], xa.l1 * pb[c]);
        const float bot = std::fmaf(xa.l0, pc[c], xa.l1 * pd[c]);
        float v = std::fmaf(ya.l0, top, ya.l1 * bot);
        if (mode == 2) v = bf16_round(v);
        o[c] = f2h(v);
      }
    }
  }
}

uint16_t orc_f2h(float f) { return f2h(f); }
float orc_h2f(uint16_t h) { return h2f(h); }
uint16_t orc_hadd(uint16_t a, uint16_t b) { return hadd(a, b); }
uint16_t orc_hsub(uint16_t a, uint16_t b) { return hsub(a, b); }
uint16_t orc_hmul(uint16_t a, uint16_t b) { return hmul(a, b); }
uint16_t orc_interp_half(float x, float y, uint16_t f00, uint16_t f01, uint16_t f10, uint16_t f11) {
  return interp_half(f2h(x), f2h(y), f00, f01, f10, f11);
}
float orc_interp_float(float x, float y, float f00, float f01, float f10, float f11) {
  return interp_float(x, y, f00, f01, f10, f11);
}
float orc_weighting(int mode, float measured, float voxel_depth, float trunc) {
  return weighting(mode, measured, voxel_depth, trunc);
}
// RayCaster::getAllIndices (ray_caster_impl.h:69-75)
int orc_raycast(const float* origin, const float* dest, int32_t* out, int cap) {
  RayCaster rc(V3{origin[0], origin[1], origin[2]}, V3{dest[0], dest[1], dest[2]});
  int n = 0, idx[3];
  while (rc.next(idx)) {
    if (n < cap) {
      out[3 * n] = idx[0];
      out[3 * n + 1] = idx[1];
      out[3 * n + 2] = idx[2];
    }
    ++n;
  }
  return n;
}
void orc_voxel_center(float block_size, const int32_t* b, const int32_t* v, float* out) {
  const V3 p = voxel_center(block_size, I3{b[0], b[1], b[2]}, I3{v[0], v[1], v[2]});
  out[0] = p.x;
  out[1] = p.y;
  out[2] = p.z;
}
void orc_block_and_voxel(float block_size, const float* p, int32_t* b, int32_t* v) {
  I3 bb, vv;
  block_and_voxel_from_position(block_size, V3{p[0], p[1], p[2]}, &bb, &vv);
  b[0] = bb.x;
  b[1] = bb.y;
  b[2] = bb.z;
  v[0] = vv.x;
  v[1] = vv.y;
  v[2] = vv.z;
}
int orc_poses_close(const float* A, const float* B, float tol_m, float tol_deg) {
  return poses_close(pose_from_row_major(A), pose_from_row_major(B), tol_m, tol_deg) ? 1 : 0;
}
int orc_project(float fx, float fy, float cx, float cy, int H, int W, const float* p, float* uv) {
  float u, v;
  if (!project(make_cam(fx, fy, cx, cy, H, W), V3{p[0], p[1], p[2]}, &u, &v)) return 0;
  uv[0] = u;
  uv[1] = v;
  return 1;
}
uint64_t orc_weld_key(const float* v) { return weld_key(V3{v[0], v[1], v[2]}); }


// ---- function-level entry points in the DEVICE model, mirrored one for one by the probes of
// oracle/ref_snippets/ref_kernels.cu (ref_fn_*): tests/test_ref_vectors.py compares them bit for bit. ----
void orc_fn_interp_half(int n, const float* xy, const uint16_t* f, uint16_t* out) {
  for (int i = 0; i < n; ++i)
    out[i] = interp_half(f2h(xy[2 * i]), f2h(xy[2 * i + 1]), f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}
void orc_fn_interp_float(int n, const float* xy, const float* f, float* out) {
  for (int i = 0; i < n; ++i)
    out[i] = interp_float(xy[2 * i], xy[2 * i + 1], f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}
// blendTwoArrays<FeatureArray> (projective_appearance_integrator.cu:286-305) on n arrays of C halves
void orc_fn_blend(int n, int C, const uint16_t* a, const uint16_t* b, const float* w, uint16_t* out) {
  for (int i = 0; i < n; ++i) {
    float w1 = w[2 * i], w2 = w[2 * i + 1];
    const float tot = w1 + w2;
    w1 /= tot;
    w2 /= tot;
    const uint16_t h1 = f2h(w1), h2 = f2h(w2);
    for (int c = 0; c < C; ++c) {
      const size_t k = (size_t)i * C + c;
      out[k] = fused_half_on() ? hfma(a[k], h1, hmul(b[k], h2)) : hadd(hmul(a[k], h1), hmul(b[k], h2));
    }
  }
}
void orc_fn_interp_vertex(int n, const float* v1, const float* v2, const float* sdf, float* out) {
  for (int i = 0; i < n; ++i) {
    const V3 r = interp_vertex(V3{v1[3 * i], v1[3 * i + 1], v1[3 * i + 2]}, V3{v2[3 * i], v2[3 * i + 1], v2[3 * i + 2]},
                               sdf[2 * i], sdf[2 * i + 1]);
    out[3 * i] = r.x;
    out[3 * i + 1] = r.y;
    out[3 * i + 2] = r.z;
  }
}
void orc_fn_block_voxel(int n, float block_size, const float* p, int32_t* out) {
  for (int i = 0; i < n; ++i) {
    I3 b, v;
    dev_block_and_voxel_from_position(block_size, V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}, &b, &v);
    out[6 * i] = b.x;
    out[6 * i + 1] = b.y;
    out[6 * i + 2] = b.z;
    out[6 * i + 3] = v.x;
    out[6 * i + 4] = v.y;
    out[6 * i + 5] = v.z;
  }
}
// projectThreadVoxel: out = (u, v, depth, p_C.xyz), ok = its return value
void orc_fn_project(int n, const float* T_C_L_rm, float fx, float fy, float cx, float cy, int W, int H, float block_size,
                    float max_depth, const int32_t* bv, float* out, int32_t* ok) {
  const Pose T = pose_from_row_major(T_C_L_rm);
  Cam cam{fx, fy, cx, cy, W, H};
  for (int i = 0; i < n; ++i) {
    const V3 pl = dev_voxel_center(block_size, I3{bv[6 * i], bv[6 * i + 1], bv[6 * i + 2]},
                                   I3{bv[6 * i + 3], bv[6 * i + 4], bv[6 * i + 5]});
    const V3 pc = dev_xform(T, pl);
    float u = 0.0f, v = 0.0f;
    bool good = dev_project(cam, pc, &u, &v);
    if (!good && pc.z >= 1e-6f) {  // the reference leaves the out-of-image coordinates in u_px
      u = nvcc_model() ? std::fmaf(pc.x / pc.z, cam.fu, cam.cu) : (pc.x / pc.z) * cam.fu + cam.cu;
      v = nvcc_model() ? std::fmaf(pc.y / pc.z, cam.fv, cam.cv) : (pc.y / pc.z) * cam.fv + cam.cv;
    }
    float depth = 0.0f;
    if (good) {
      depth = pc.z;
      if (max_depth > 0.0f && depth > max_depth) good = false;
    }
    out[6 * i] = u;
    out[6 * i + 1] = v;
    out[6 * i + 2] = depth;
    out[6 * i + 3] = pc.x;
    out[6 * i + 4] = pc.y;
    out[6 * i + 5] = pc.z;
    ok[i] = good ? 1 : 0;
  }
}
// UpdateTsdfVoxelFunctor (projective_tsdf_integrator.cu:25-99) on explicit tuples
void orc_fn_tsdf_functor(int n, float trunc, float max_weight, float invalid_decay, int mode, const float* in,
                         const uint8_t* active, float* voxels, uint8_t* updated) {
  for (int i = 0; i < n; ++i) {
    const float meas = in[2 * i], vd = in[2 * i + 1];
    float& d = voxels[2 * i];
    float& w = voxels[2 * i + 1];
    updated[i] = 0;
    if (meas <= 0.0f) {
      if (invalid_decay >= 0.0f) w *= invalid_decay;
      continue;
    }
    const float sdf = meas - vd;
    if (sdf < -trunc) continue;
    if (!active[i] && sdf < trunc) continue;
    const float w_m = weighting(mode, meas, vd, trunc);
    float fused = (nvcc_model() ? std::fmaf(d, w, sdf * w_m) : (sdf * w_m + d * w)) / (w_m + w);
    if (fused > 0.0f)
      fused = std::fmin(trunc, fused);
    else
      fused = std::fmax(-trunc, fused);
    const float w_new = std::fmin(w_m + w, max_weight);
    d = fused;
    w = w_new;
    updated[i] = 1;
  }
}
}  // extern "C"
