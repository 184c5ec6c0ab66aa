// N4: the step BEFORE the path, fused into it.
//
// mindmap never hands nvblox the vision backbone's native output.  FeatureExtractor.compute()
// (mindmap/image_processing/feature_extraction.py:170-211) takes the [1, c, h, w] patch features of the
// backbone (RADIO: 32 x 32 x 768 for a 512^2 image), up-samples them to the camera resolution with
// torch.nn.functional.interpolate(mode='bilinear', align_corners=False) (:110-129), re-arranges to HWC,
// zero-pads to nvblox's channel count (:198-210) and converts to fp16
// (mindmap/mapping/helpers/nvblox_mapping_helpers.py:256).  The result, a 384 MiB frame, is what
// add_feature_frame() then gathers 4 bilinear neighbours per voxel from.
//
// Here the gather reads the LOW-RES map (1.5-3 MB, L2 resident) and evaluates the up-sampling for exactly
// the <= 4 x N_upd high-res pixels the frame touches; the 384 MiB frame and the ~3 GB of HBM traffic that
// interpolate + cat + .half() spend producing it never exist.
//
// Arithmetic contract: the value of a high-res pixel is what PyTorch's CUDA kernel produces, bit for
// bit.  The algorithm lives in a third-party dependency of the reference (PyTorch,
// aten/src/ATen/native/cuda/UpSampleBilinear2d.cu + UpSample.cuh:114-130), so it is restated from the
// published source AND from the sm_100 SASS of the torch 2.11 build in this image (cuobjdump of
// upsample_bilinear2d_out_frame / upsample_bilinear2d_nhwc_out_frame), which fixes where nvcc
// contracted multiply-adds:
//
//   s       = fma(dst + 0.5, scale, -0.5), clamped below at 0        (scale = float(in) / float(out))
//   i0      = trunc(s);  i1 = i0 + (i0 < in - 1);  l1 = s - i0;  l0 = 1 - l1
//   top     = fma(w0, a, w1 * b)          NCHW kernels (f32, f16, bf16) and the NHWC f16 / bf16 kernels
//           = fma(w1, b, w0 * a)          NHWC f32 kernel (what RADIO's permuted [1,c,h,w] view selects)
//   bottom  = fma(w0, c, w1 * d)
//   value  = fma(h0, top, h1 * bottom)    -> rounded to the tensor dtype, then .to(float16)
//
// The library is compiled -fmad=false, so every fma / mul below is explicit.
#pragma once

#include <cuda_bf16.h>

#include "nvbx_kernels.cuh"

namespace nvbx {

// MODE: 0 = NCHW-kernel order, 1 = NHWC-f32-kernel order, 2 = NCHW order with a bf16 intermediate rounding
struct UpFrame {
  const float* low;  // [lh][lw][C] fp32, zero-padded to the map's C channels (16-byte aligned)
  int lh, lw;
  float rh, rw;      // float(lh) / float(H), float(lw) / float(W)
};

struct UpAxis {
  int i0, i1;
  float l0, l1;
};

// area_pixel_compute_source_index (UpSample.cuh:114-130, align_corners = false, cubic = false) followed by the
// index / lambda computation of upsample_bilinear2d_out_frame.
__device__ __forceinline__ UpAxis up_axis(float scale, int dst, int in_size) {
  UpAxis a;
  float s = __fmaf_rn((float)dst + 0.5f, scale, -0.5f);
  s = (s >= 0.0f) ? s : 0.0f;
  a.i0 = (int)s;
  a.i1 = a.i0 + ((a.i0 < in_size - 1) ? 1 : 0);
  a.l1 = s - (float)a.i0;
  a.l0 = 1.0f - a.l1;
  return a;
}

template <int MODE>
__device__ __forceinline__ float up_top(float w0, float w1, float a, float b) {
  return MODE == 1 ? __fmaf_rn(w1, b, __fmul_rn(w0, a)) : __fmaf_rn(w0, a, __fmul_rn(w1, b));
}
__device__ __forceinline__ float up_bottom(float w0, float w1, float c, float d) {
  return __fmaf_rn(w0, c, __fmul_rn(w1, d));
}
template <int MODE>
__device__ __forceinline__ unsigned up_pack(float v0, float v1) {
  if (MODE == 2) {  // scalar_t = BFloat16: the kernel's store rounds to bf16, .to(float16) rounds again
    v0 = __bfloat162float(__float2bfloat16_rn(v0));
    v1 = __bfloat162float(__float2bfloat16_rn(v1));
  }
  const __half2 h = __floats2half2_rn(v0, v1);
  return *reinterpret_cast<const unsigned*>(&h);
}

struct F8 {
  float4 lo, hi;
};
__device__ __forceinline__ F8 ld8(const float* p) {
  F8 r;
  r.lo = __ldg(reinterpret_cast<const float4*>(p));
  r.hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return r;
}

// 8 channels of ONE high-res pixel whose column / row parameters are (x, y); `base` already points at the lane's
// first channel of low-res pixel (0, 0).
template <int MODE>
__device__ __forceinline__ uint4 up_pixel8(const float* base, int lw, int C, const UpAxis& x, const UpAxis& y) {
  const F8 a = ld8(base + ((size_t)y.i0 * lw + x.i0) * C);
  const F8 b = ld8(base + ((size_t)y.i0 * lw + x.i1) * C);
  const F8 c = ld8(base + ((size_t)y.i1 * lw + x.i0) * C);
  const F8 d = ld8(base + ((size_t)y.i1 * lw + x.i1) * C);
  const float* pa = reinterpret_cast<const float*>(&a);
  const float* pb = reinterpret_cast<const float*>(&b);
  const float* pc = reinterpret_cast<const float*>(&c);
  const float* pd = reinterpret_cast<const float*>(&d);
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    v[k] = __fmaf_rn(y.l0, up_top<MODE>(x.l0, x.l1, pa[k], pb[k]),
                     __fmul_rn(y.l1, up_bottom(x.l0, x.l1, pc[k], pd[k])));
  return make_uint4(up_pack<MODE>(v[0], v[1]), up_pack<MODE>(v[2], v[3]), up_pack<MODE>(v[4], v[5]),
                    up_pack<MODE>(v[6], v[7]));
}

// The four high-res neighbours (px,py) (px+1,py) (px,py+1) (px+1,py+1) of one work item, 8 channels each.
// With 16x up-sampling 88 % of the 2x2 patches sit inside ONE low-res cell: the four corners are loaded once
// and the row sums `top` / `bottom` are shared between the two rows of the patch.
template <int MODE>
__device__ __forceinline__ void up_patch8(const float* base, int lw, int C, const UpAxis& x0, const UpAxis& x1,
                                          const UpAxis& y0, const UpAxis& y1, uint4& r00, uint4& r10, uint4& r01,
                                          uint4& r11) {
  if (x0.i0 == x1.i0 && y0.i0 == y1.i0) {  // warp-uniform
    const F8 a = ld8(base + ((size_t)y0.i0 * lw + x0.i0) * C);
    const F8 b = ld8(base + ((size_t)y0.i0 * lw + x0.i1) * C);
    const F8 c = ld8(base + ((size_t)y0.i1 * lw + x0.i0) * C);
    const F8 d = ld8(base + ((size_t)y0.i1 * lw + x0.i1) * C);
    const float* pa = reinterpret_cast<const float*>(&a);
    const float* pb = reinterpret_cast<const float*>(&b);
    const float* pc = reinterpret_cast<const float*>(&c);
    const float* pd = reinterpret_cast<const float*>(&d);
    unsigned o00[4], o10[4], o01[4], o11[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v00[2], v10[2], v01[2], v11[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ch = 2 * k + j;
        const float t0 = up_top<MODE>(x0.l0, x0.l1, pa[ch], pb[ch]);
        const float b0 = up_bottom(x0.l0, x0.l1, pc[ch], pd[ch]);
        const float t1 = up_top<MODE>(x1.l0, x1.l1, pa[ch], pb[ch]);
        const float b1 = up_bottom(x1.l0, x1.l1, pc[ch], pd[ch]);
        v00[j] = __fmaf_rn(y0.l0, t0, __fmul_rn(y0.l1, b0));
        v10[j] = __fmaf_rn(y0.l0, t1, __fmul_rn(y0.l1, b1));
        v01[j] = __fmaf_rn(y1.l0, t0, __fmul_rn(y1.l1, b0));
        v11[j] = __fmaf_rn(y1.l0, t1, __fmul_rn(y1.l1, b1));
      }
      o00[k] = up_pack<MODE>(v00[0], v00[1]);
      o10[k] = up_pack<MODE>(v10[0], v10[1]);
      o01[k] = up_pack<MODE>(v01[0], v01[1]);
      o11[k] = up_pack<MODE>(v11[0], v11[1]);
    }
    r00 = make_uint4(o00[0], o00[1], o00[2], o00[3]);
    r10 = make_uint4(o10[0], o10[1], o10[2], o10[3]);
    r01 = make_uint4(o01[0], o01[1], o01[2], o01[3]);
    r11 = make_uint4(o11[0], o11[1], o11[2], o11[3]);
  } else {
    r00 = up_pixel8<MODE>(base, lw, C, x0, y0);
    r10 = up_pixel8<MODE>(base, lw, C, x1, y0);
    r01 = up_pixel8<MODE>(base, lw, C, x0, y1);
    r11 = up_pixel8<MODE>(base, lw, C, x1, y1);
  }
}

// Same work decomposition as k_feature_gather (warp = 512-byte chunk of one item's channel vector, persistent
// grid); the four 128-bit image loads are replaced by up_patch8 over the L2-resident low-res map.
// Algorithmic bytes: 2(C+1) x N_upd written (+ 2C x N_upd when blending) + 4 C lh lw read once.
template <int CH, int MODE, int CTAS>
__global__ void __launch_bounds__(256, CTAS) k_feature_gather_up(MapDev m, const FeatItem* __restrict__ items,
                                                                 FeatFrame f, UpFrame uf, int last_chunk) {
  pdl_prologue();
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

  for (long long q = warp; q < n_units; q += warps_total) {
    const int item = (int)(q / ch_per_item);
    const int cvec = (int)(q - (long long)item * ch_per_item) * 32 + lane;
    FeatItem it;
    *reinterpret_cast<uint4*>(&it) = __ldg(reinterpret_cast<const uint4*>(items + item));
    if (CH == 0 && cvec >= nvec) continue;
    const int py = it.pix / f.cols, px = it.pix - py * f.cols;
    const UpAxis x0 = up_axis(uf.rw, px, uf.lw), x1 = up_axis(uf.rw, px + 1, uf.lw);
    const UpAxis y0 = up_axis(uf.rh, py, uf.lh), y1 = up_axis(uf.rh, py + 1, uf.lh);
    uint4 a00, a10, a01, a11;
    up_patch8<MODE>(uf.low + (size_t)cvec * 8, uf.lw, C, x0, x1, y0, y1, a00, a10, a01, a11);

    const int fslot = it.row >> 9, vox = it.row & 511;
    uint4* dst = reinterpret_cast<uint4*>(feat_block(m, fslot)) + (size_t)vox * row_vecs;
    const bool blend = (!it.first) && f.read_old;
    uint4 old = make_uint4(0, 0, 0, 0);
    if (blend) old = dst[cvec];
    const __half hx = __ushort_as_half(it.hx), hy = __ushort_as_half(it.hy);
    uint4 o = interp_vec(__half2half2(hx), __half2half2(hy), __half2half2(__hmul_rn(hx, hy)), a00, a01, a10, a11);
    if (blend) o = blend_vec(old, o, w1, w2);
    dst[cvec] = o;
  }
  if (last_chunk && blockIdx.x == 0 && threadIdx.x == 0) {
    m.ctrl->band_count[m.fp] = 0;
    m.ctrl->newfeat_count[m.fp] = 0;
  }
}

// The whole [H, W, C] fp16 frame the chained path would have produced (parity checks, visualisation; not on
// the per-frame path).  Thread = 8 channels of one high-res pixel.
template <int MODE>
__global__ void __launch_bounds__(256) k_upsample_materialise(UpFrame uf, int H, int W, int C,
                                                              __half* __restrict__ out) {
  pdl_prologue();
  const int nvec = C >> 3;
  const long long total = (long long)H * W * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cvec = (int)(i % nvec);
    const long long pix = i / nvec;
    const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
    const UpAxis xa = up_axis(uf.rw, x, uf.lw), ya = up_axis(uf.rh, y, uf.lh);
    reinterpret_cast<uint4*>(out)[i] = up_pixel8<MODE>(uf.low + (size_t)cvec * 8, uf.lw, C, xa, ya);
  }
}

// Stage the backbone output as [lh][lw][C] fp32: dtype conversion (exact), CHW -> HWC, zero-padding of the
// channels the backbone does not produce (feature_extraction.py:198-210).  <= 3 MB; no dependency on the map,
// so it overlaps the depth kernels of the same frame.
__global__ void __launch_bounds__(256) k_lowres_stage(const void* __restrict__ src, int dtype, int chw, int lh, int lw,
                                                      int lc, int C, float* __restrict__ dst) {
  pdl_prologue();
  const long long total = (long long)lh * lw * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;  // y * lw + x
    float v = 0.0f;
    if (c < lc) {
      const long long j = chw ? ((long long)c * lh * lw + p) : (p * lc + c);
      if (dtype == 0)
        v = reinterpret_cast<const float*>(src)[j];
      else if (dtype == 1)
        v = __half2float(reinterpret_cast<const __half*>(src)[j]);
      else
        v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[j]);
    }
    dst[i] = v;
  }
}

}  // namespace nvbx
