// nvbx_math.cuh -- geometry shared by host orchestration and device kernels.
//
// Floating-point contract (DESIGN.md "Numerics"): every fp32 expression below is evaluated as
// individually rounded IEEE operations in the order Eigen 3.4 evaluates the reference's expression
// (3-term reductions are a0 + (a1 + a2); Isometry * v is t + R.row.v).  The library is compiled with
// -fmad=false (device) and -ffp-contract=off (host) so the compiler never fuses a mul+add; divisions
// and square roots are IEEE (nvcc defaults -prec-div=true -prec-sqrt=true).  That makes host, device
// and the CPU oracle agree bit for bit, which is what gives bit-exact block sets.
#pragma once
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#define NVBX_HD __host__ __device__ __forceinline__

namespace nvbx {

struct V3 {
  float x, y, z;
};
struct I3 {
  int x, y, z;
};
// Eigen::Isometry3f: p' = t + R p.   (NB/include/nvblox/core/types.h:145)
struct Pose {
  float R[3][3];
  float t[3];
};
// Pinhole camera, NB/include/nvblox/sensors/camera.h
struct Cam {
  float fu, fv, cu, cv;
  int width, height;
};

NVBX_HD float sum3(float a, float b, float c) { return a + (b + c); }

NVBX_HD V3 xform(const Pose& T, const V3& v) {
  V3 o;
  o.x = T.t[0] + sum3(T.R[0][0] * v.x, T.R[0][1] * v.y, T.R[0][2] * v.z);
  o.y = T.t[1] + sum3(T.R[1][0] * v.x, T.R[1][1] * v.y, T.R[1][2] * v.z);
  o.z = T.t[2] + sum3(T.R[2][0] * v.x, T.R[2][1] * v.y, T.R[2][2] * v.z);
  return o;
}
NVBX_HD V3 rotate(const Pose& T, const V3& v) {
  V3 o;
  o.x = sum3(T.R[0][0] * v.x, T.R[0][1] * v.y, T.R[0][2] * v.z);
  o.y = sum3(T.R[1][0] * v.x, T.R[1][1] * v.y, T.R[1][2] * v.z);
  o.z = sum3(T.R[2][0] * v.x, T.R[2][1] * v.y, T.R[2][2] * v.z);
  return o;
}
// Isometry inverse: R' = R^T, t' = (-R^T) t
inline Pose inverse(const Pose& T) {
  Pose o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.R[i][j] = T.R[j][i];
  for (int i = 0; i < 3; ++i)
    o.t[i] = sum3((-o.R[i][0]) * T.t[0], (-o.R[i][1]) * T.t[1], (-o.R[i][2]) * T.t[2]);
  return o;
}
inline Pose pose_from_row_major(const float* m) {
  Pose p;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.R[i][j] = m[i * 4 + j];
    p.t[i] = m[i * 4 + 3];
  }
  return p;
}

// ---- indexing: NB/include/nvblox/core/internal/impl/indexing_impl.h:22-81 ------------------------------
NVBX_HD I3 block_index_from_position(float block_size, const V3& p) {
  I3 r;
  r.x = (int)floorf(p.x / block_size);
  r.y = (int)floorf(p.y / block_size);
  r.z = (int)floorf(p.z / block_size);
  return r;
}
// voxel_size_inv is float(1.0 / double(voxel_size)) computed once on the host (indexing_impl.h:43)
NVBX_HD void block_and_voxel_from_position(float block_size, float voxel_size_inv, const V3& p, I3* b, I3* v) {
  *b = block_index_from_position(block_size, p);
  const float rx = (p.x - block_size * (float)b->x) * voxel_size_inv;
  const float ry = (p.y - block_size * (float)b->y) * voxel_size_inv;
  const float rz = (p.z - block_size * (float)b->z) * voxel_size_inv;
  v->x = min((int)rx, 7);
  v->y = min((int)ry, 7);
  v->z = min((int)rz, 7);
}
NVBX_HD V3 voxel_center(float block_size, const I3& b, int vx, int vy, int vz) {
  const float voxel_size = block_size * (1.0f / 8.0f);
  const float half_voxel = block_size * (0.5f / 8.0f);
  V3 p;
  p.x = (block_size * (float)b.x + voxel_size * (float)vx) + half_voxel;
  p.y = (block_size * (float)b.y + voxel_size * (float)vy) + half_voxel;
  p.z = (block_size * (float)b.z + voxel_size * (float)vz) + half_voxel;
  return p;
}

// ---- camera: NB/include/nvblox/sensors/internal/impl/camera_impl.h:20-91 -----------------------------------
NVBX_HD V3 ray_from_image_plane(const Cam& c, float u, float v) {
  V3 r;
  r.x = (u - c.cu) / c.fu;
  r.y = (v - c.cv) / c.fv;
  r.z = 1.0f;
  return r;
}
NVBX_HD bool project(const Cam& c, const V3& p, float* u, float* v) {
  if (!(p.z >= 1e-6f)) return false;
  float un = p.x / p.z;
  float vn = p.y / p.z;
  un = un * c.fu + c.cu;
  vn = vn * c.fv + c.cv;
  if (un > (float)c.width || vn > (float)c.height || un < 0 || vn < 0) return false;
  *u = un;
  *v = vn;
  return true;
}

// interpolatePixels<float>, NB/include/nvblox/interpolation/internal/impl/interpolation_2d_impl.h:33-48
NVBX_HD float interp_float(float x, float y, float f00, float f01, float f10, float f11) {
  const float dx = f10 - f00;
  return ((f00 + x * dx) + y * (f01 - f00)) + (x * y) * ((f11 - f01) - dx);
}

// ---- weighting: NB/include/nvblox/integrators/internal/impl/weighting_function_impl.h:29-117 ------------------
NVBX_HD float weight_dropoff(float measured, float voxel_depth, float trunc) {
  if (trunc <= 1e-2f) return 0.0f;
  if (voxel_depth > measured) {
    const float behind = voxel_depth - measured;
    if (behind > trunc) return 0.0f;
    return (trunc - behind) / trunc;
  }
  return 1.0f;
}
NVBX_HD float weight_inverse_square(float measured, float voxel_depth, float trunc) {
  if (voxel_depth <= 1e-2f) return 1.0f;
  if (voxel_depth - measured >= trunc) return 0.0f;
  return 1.0f / (voxel_depth * voxel_depth);
}
NVBX_HD float weighting(int mode, float measured, float voxel_depth, float trunc) {
  switch (mode) {
    case 0:
      return 1.0f;
    case 1:
      return 1.0f * weight_dropoff(measured, voxel_depth, trunc);
    case 2:
      return weight_inverse_square(measured, voxel_depth, trunc);
    case 3:
      return weight_inverse_square(measured, voxel_depth, trunc) * weight_dropoff(measured, voxel_depth, trunc);
    case 4: {
      const float d = measured - voxel_depth;
      return weight_inverse_square(measured, voxel_depth, trunc) * ((fabsf(d) >= trunc) ? 0.1f : 1.0f);
    }
    case 5:
      return voxel_depth > 1.0f ? 1.0f / voxel_depth : 1.0f;
  }
  return 0.0f;
}

// ---- DEVICE-code forms (nvcc-contracted, see the header) ---------------------------------------------------
// R.row(i) . v: Eigen's a0 + (a1 + a2) becomes fma(v.x, R0, fma(v.y, R1, v.z * R2))
NVBX_HD float dev_dot3(const float* r, const V3& v) { return fmaf(v.x, r[0], fmaf(v.y, r[1], v.z * r[2])); }
NVBX_HD V3 dev_rotate(const Pose& T, const V3& v) {
  V3 o;
  o.x = dev_dot3(T.R[0], v);
  o.y = dev_dot3(T.R[1], v);
  o.z = dev_dot3(T.R[2], v);
  return o;
}
// Transform * Vector3f: the translation is added last, unfused
NVBX_HD V3 dev_xform(const Pose& T, const V3& v) {
  V3 o;
  o.x = T.t[0] + dev_dot3(T.R[0], v);
  o.y = T.t[1] + dev_dot3(T.R[1], v);
  o.z = T.t[2] + dev_dot3(T.R[2], v);
  return o;
}
// getCenterPositionFromBlockIndexAndVoxelIndex inside projectThreadVoxel: fma(bs, 1/16, fma(bs, b, (bs/8) * v))
NVBX_HD V3 dev_voxel_center(float block_size, const I3& b, int vx, int vy, int vz) {
  const float voxel_size = block_size * (1.0f / 8.0f);
  V3 p;
  p.x = fmaf(block_size, 0.5f / 8.0f, fmaf(block_size, (float)b.x, voxel_size * (float)vx));
  p.y = fmaf(block_size, 0.5f / 8.0f, fmaf(block_size, (float)b.y, voxel_size * (float)vy));
  p.z = fmaf(block_size, 0.5f / 8.0f, fmaf(block_size, (float)b.z, voxel_size * (float)vz));
  return p;
}
// Camera::project inside a kernel: one FFMA per axis for the intrinsics
NVBX_HD bool dev_project(const Cam& c, const V3& p, float* u, float* v) {
  if (!(p.z >= 1e-6f)) return false;
  const float un = fmaf(p.x / p.z, c.fu, c.cu);
  const float vn = fmaf(p.y / p.z, c.fv, c.cv);
  if (un > (float)c.width || vn > (float)c.height || un < 0 || vn < 0) return false;
  *u = un;
  *v = vn;
  return true;
}
// interpolatePixels<float> inside the appearance kernel: three FFMAs
NVBX_HD float dev_interp_float(float x, float y, float f00, float f01, float f10, float f11) {
  const float dx = f10 - f00;
  return fmaf(x * y, (f11 - f01) - dx, fmaf(y, f01 - f00, fmaf(x, dx, f00)));
}
// getBlockAndVoxelIndexFromPositionInLayer inside a kernel: p - bs*b is one FFMA
NVBX_HD void dev_block_and_voxel_from_position(float block_size, float voxel_size_inv, const V3& p, I3* b, I3* v) {
  *b = block_index_from_position(block_size, p);
  v->x = min((int)(fmaf(-(float)b->x, block_size, p.x) * voxel_size_inv), 7);
  v->y = min((int)(fmaf(-(float)b->y, block_size, p.y) * voxel_size_inv), 7);
  v->z = min((int)(fmaf(-(float)b->z, block_size, p.z) * voxel_size_inv), 7);
}

// ---- AABB helpers (host): camera.cpp:51-103,153-166, workspace_bounds.cpp:20-61 ----------------------
struct Aabb {
  float mn[3], mx[3];
  bool empty() const { return mn[0] > mx[0] || mn[1] > mx[1] || mn[2] > mx[2]; }
};
inline Aabb view_aabb(const Cam& c, const Pose& T_L_C, float min_depth, float max_depth) {
  const V3 rays[4] = {ray_from_image_plane(c, (float)c.width, (float)c.height),
                      ray_from_image_plane(c, (float)c.width, 0.0f), ray_from_image_plane(c, 0.0f, 0.0f),
                      ray_from_image_plane(c, 0.0f, (float)c.height)};
  Aabb a;
  for (int i = 0; i < 3; ++i) {
    a.mn[i] = 3.402823466e+38f;
    a.mx[i] = -3.402823466e+38f;
  }
  for (int k = 0; k < 8; ++k) {
    const float d = k < 4 ? min_depth : max_depth;
    const V3& r = rays[k & 3];
    V3 pc;
    pc.x = d * r.x;
    pc.y = d * r.y;
    pc.z = d * r.z;
    const V3 pl = xform(T_L_C, pc);
    const float v[3] = {pl.x, pl.y, pl.z};
    for (int i = 0; i < 3; ++i) {
      a.mn[i] = fminf(a.mn[i], v[i]);
      a.mx[i] = fmaxf(a.mx[i], v[i]);
    }
  }
  return a;
}

// ---- viewpoint cache equality (host): camera.cpp:31-49, transforms.cpp:20-37 ----------------------------
inline bool poses_close(const Pose& A, const Pose& B, float tol_m, float tol_deg) {
  const Pose Ai = inverse(A);
  float R[3][3], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      R[i][j] = sum3(Ai.R[i][0] * B.R[0][j], Ai.R[i][1] * B.R[1][j], Ai.R[i][2] * B.R[2][j]);
    t[i] = Ai.t[i] + sum3(Ai.R[i][0] * B.t[0], Ai.R[i][1] * B.t[1], Ai.R[i][2] * B.t[2]);
  }
  const float n = sqrtf(sum3(t[0] * t[0], t[1] * t[1], t[2] * t[2]));
  if (n > tol_m) return false;
  float qw, qx, qy, qz;
  float tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0f) {
    tr = sqrtf(tr + 1.0f);
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
    tr = sqrtf(R[i][i] - R[j][j] - R[k][k] + 1.0f);
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
  const float vn = sqrtf(sum3(qx * qx, qy * qy, qz * qz));
  const float angle = 2.0f * atan2f(vn, fabsf(qw));
  const float deg = (float)((double)(angle * 180.0f) / 3.14159265358979323846);
  return !(fabsf(deg) > tol_deg);
}
inline bool cameras_equal(const Cam& a, const Cam& b, const Pose& Ta, const Pose& Tb) {
  const bool ext = poses_close(Ta, Tb, 0.001f, 0.1f);
  bool in = true;
  in &= fabsf(a.fu - b.fu) <= 0.1;
  in &= fabsf(a.fv - b.fv) <= 0.1;
  in &= fabsf(a.cu - b.cu) <= 0.1;
  in &= fabsf(a.cv - b.cv) <= 0.1;
  in &= a.width == b.width;
  in &= a.height == b.height;
  return ext && in;
}

}  // namespace nvbx
