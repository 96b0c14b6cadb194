// glue.cu -- the HBM-bound kernels around the tensor-core convolutions.
//
// Everything between two convolutions of the reference graph is one pass here: instance norm
// (tfa.layers.InstanceNormalization, module.py:212...), its activation (module.py:213,285...),
// the residual add (module.py:217), and the padding the next convolution wants (tf.pad REFLECT,
// module.py:210,214,230,262; Keras 'same' zero padding; the 2x2 phase split a stride-2 layer
// reads) are fused into a single read of the raw conv output and a single write of the next
// frame.  All accesses are 128-bit (8 bf16 channels per thread), NHWC, coalesced along C.
#include "glue.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <math.h>

namespace sggan {

// ------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void ld_bf16x8(const sg_bf16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void st_bf16x8(sg_bf16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float act_fwd(float v, int act, float a) {
  if (act == SG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == SG_ACT_LRELU) return v > 0.f ? v : a * v;
  if (act == SG_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float act_grad(float zpre, int act, float a) {
  if (act == SG_ACT_RELU) return zpre > 0.f ? 1.f : 0.f;
  if (act == SG_ACT_LRELU) return zpre > 0.f ? 1.f : a;
  return 1.f;
}
// Rows (or columns) that logical index i occupies in a reflect-padded plane: itself, plus its
// mirror images in the border of width p.
__device__ __forceinline__ int reflect_set(int i, int n, int p, int* out) {
  int k = 0;
  out[k++] = i;
  if (p > 0) {
    if (i >= 1 && i <= p) out[k++] = -i;
    if (i >= n - 1 - p && i <= n - 2) out[k++] = 2 * (n - 1) - i;
  }
  return k;
}
__device__ __forceinline__ void write_frame8(sg_bf16* dst, const FrameMap& m, int b, int i, int j, int c0,
                                             const float* v) {
  sg_bf16* base = dst + int64_t(b) * m.frame_pix * m.C + c0;
  if (m.kind == 0 && m.reflect > 0) {
    int rr[3], cc[3];
    const int nr = reflect_set(i, m.H, m.reflect, rr), nc = reflect_set(j, m.W, m.reflect, cc);
    for (int a = 0; a < nr; ++a)
      for (int q = 0; q < nc; ++q) st_bf16x8(base + frame_pixel(m, rr[a], cc[q]) * m.C, v);
  } else {
    st_bf16x8(base + frame_pixel(m, i, j) * m.C, v);
  }
}
// acc[0..8) += gradient source at logical (i, j), channels c0..c0+7 (folding reflected borders back)
__device__ __forceinline__ void grad_read8(const GradSrc& g, int b, int i, int j, int H, int W, int C, int c0,
                                           float* acc) {
  if (g.ptr == nullptr) return;
  int rr[3], cc[3];
  int nr = 1, nc = 1;
  rr[0] = i;
  cc[0] = j;
  // only pixels within `fold` of the image edge receive folded-back border gradient
  if (g.fold > 0 && (i <= g.fold || i >= H - 1 - g.fold || j <= g.fold || j >= W - 1 - g.fold)) {
    nr = reflect_set(i, H, g.fold, rr);
    nc = reflect_set(j, W, g.fold, cc);
  }
  for (int a = 0; a < nr; ++a)
    for (int q = 0; q < nc; ++q) {
      const int64_t idx = ((int64_t(b) * g.Hs + (rr[a] + g.oy)) * g.Ws + (cc[q] + g.ox)) * C + c0;
      if (g.f32) {
        const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.ptr) + idx);
        const float4 u0 = __ldg(s), u1 = __ldg(s + 1);
        acc[0] += u0.x; acc[1] += u0.y; acc[2] += u0.z; acc[3] += u0.w;
        acc[4] += u1.x; acc[5] += u1.y; acc[6] += u1.z; acc[7] += u1.w;
      } else {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const sg_bf16*>(g.ptr) + idx));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t = __bfloat1622float2(h[k]);
          acc[2 * k] += t.x;
          acc[2 * k + 1] += t.y;
        }
      }
    }
}

constexpr int kGlueThreads = 256;
static inline int pix_per_block_for(int HW, int B) {
  // about 3 blocks per SM over the whole grid (fat blocks amortise the per-block coefficient set-up),
  // at least 64 pixels per block
  const int want_blocks_per_image = (148 * 3 + B - 1) / B;
  int ppb = (HW + want_blocks_per_image - 1) / want_blocks_per_image;
  ppb = (ppb + 63) / 64 * 64;
  return ppb < 64 ? 64 : ppb;
}

// ------------------------------------------------------------------------------------------ prep
// Four pixels (12 fp32 = three 128-bit loads) per thread; the 8-channel bf16 frame pixels are written through write_frame8
// (reflected copies included), 16 bytes each -- consecutive pixels of a row are consecutive in the frame, so a warp writes
// 2 KB runs.  W % 4 == 0 and a 16-byte aligned source (checked by the launcher), else one pixel per thread.
__global__ void __launch_bounds__(256) prep_image3_kernel(const float* __restrict__ src, int B, int H, int W, sg_bf16* dst,
                                                          FrameMap m, int dst_b0, int vec) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t tot = int64_t(B) * H * W;
  if (vec) {
    const int64_t p0 = idx * 4;
    if (p0 >= tot) return;
    const int b = int(p0 / (int64_t(H) * W));
    const int r = int(p0 - int64_t(b) * H * W);
    const int i = r / W, j = r - i * W;
    const float4* s4 = reinterpret_cast<const float4*>(src + p0 * 3);
    const float4 a = __ldg(s4), c = __ldg(s4 + 1), d = __ldg(s4 + 2);
    const float f[12] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v[8] = {f[3 * q], f[3 * q + 1], f[3 * q + 2], 0.f, 0.f, 0.f, 0.f, 0.f};
      write_frame8(dst, m, dst_b0 + b, i, j + q, 0, v);
    }
    return;
  }
  if (idx >= tot) return;
  const int b = int(idx / (int64_t(H) * W));
  const int r = int(idx - int64_t(b) * H * W);
  const int i = r / W, j = r - i * W;
  const float* s = src + idx * 3;
  float v[8] = {s[0], s[1], s[2], 0.f, 0.f, 0.f, 0.f, 0.f};
  write_frame8(dst, m, dst_b0 + b, i, j, 0, v);
}
void launch_prep_image3(const float* src, int B, int H, int W, sg_bf16* dst, const FrameMap& dmap, int dst_b0,
                        cudaStream_t st) {
  const int64_t tot = int64_t(B) * H * W;
  const int vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int64_t nthr = vec ? tot / 4 : tot;
  prep_image3_kernel<<<unsigned((nthr + 255) / 256), 256, 0, st>>>(src, B, H, W, dst, dmap, dst_b0, vec);
}

__global__ void f32_to_frame_kernel(const float* __restrict__ src, int B, int H, int W, int C, sg_bf16* dst,
                                    FrameMap m) {
  const int C8 = C >> 3;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t tot = int64_t(B) * H * W * C8;
  if (idx >= tot) return;
  const int cg = int(idx % C8);
  const int64_t pix = idx / C8;
  const int b = int(pix / (int64_t(H) * W));
  const int r = int(pix - int64_t(b) * H * W);
  const int i = r / W, j = r - i * W;
  const float* s = src + pix * C + cg * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = s[e];
  write_frame8(dst, m, b, i, j, cg * 8, v);
}
void launch_f32_to_frame(const float* src, int B, int H, int W, int C, sg_bf16* dst, const FrameMap& dmap,
                         cudaStream_t st) {
  const int64_t tot = int64_t(B) * H * W * (C / 8);
  f32_to_frame_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(src, B, H, W, C, dst, dmap);
}

__global__ void bf16_to_f32_kernel(const sg_bf16* __restrict__ s, float* d, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s)[i]);
}
void launch_bf16_to_f32(const sg_bf16* src, float* dst, int64_t n, cudaStream_t st) {
  bf16_to_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(src, dst, n);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ s, sg_bf16* d, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<__nv_bfloat16*>(d)[i] = __float2bfloat16_rn(s[i]);
}
void launch_f32_to_bf16(const float* src, sg_bf16* dst, int64_t n, cudaStream_t st) {
  f32_to_bf16_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(src, dst, n);
}
__global__ void lrelu_kernel(const float* __restrict__ x, float* y, int64_t n, float leak) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(x[i], leak * x[i]);
}
void launch_lrelu(const float* x, float* y, int64_t n, float leak, cudaStream_t st) {
  lrelu_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(x, y, n, leak);
}

// ------------------------------------------------------------------------------------------ fp32 / tf32 tier
// Helpers of the fp32-accurate operator tier (ops.conv2d / deconv2d / instance_norm with precision="tf32"): activations
// stay fp32 in HBM, the convolution's operands are rounded to tf32 (cvt.rna: nearest, what cuDNN's TF32 mode feeds the
// tensor cores) when the frame / the weight slabs are built, accumulation is fp32.  Parity-first, not tuned.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// passes == 3 ("3xTF32"): the frame holds three channel groups of Kp channels each, [hi(x) | lo(x) | hi(x)] with
// hi = tf32(x), lo = tf32(x - hi); the weight slabs hold [hi(w) | hi(w) | lo(w)], so that one convolution over 3 Kp input
// channels accumulates hi*hi + lo*hi + hi*lo in fp32: the 2^-11 operand rounding of a single tf32 pass (~3e-4 relative
// on a K = 2304 dot product) drops to the ~2^-22 of the neglected lo*lo term.
__device__ __forceinline__ float tf32_part(float x, int group) {
  const float hi = round_tf32(x);
  return group == 1 ? round_tf32(x - hi) : hi;
}
__global__ void f32_to_frame_f32_kernel(const float* __restrict__ src, int B, int H, int W, int Cs, float* dst, FrameMap m, int Kp) {
  const int C4 = m.C >> 2;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(B) * H * W * C4) return;
  const int cg = int(idx % C4);
  const int64_t pix = idx / C4;
  const int b = int(pix / (int64_t(H) * W));
  const int r = int(pix - int64_t(b) * H * W);
  const int i = r / W, j = r - i * W;
  float4 v;
  float* vv = &v.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cg * 4 + e, grp = c / Kp, cc = c - grp * Kp;
    vv[e] = cc < Cs ? tf32_part(src[pix * Cs + cc], grp) : 0.f;
  }
  float* base = dst + int64_t(b) * m.frame_pix * m.C + cg * 4;
  if (m.kind == 0 && m.reflect > 0) {
    int rr[3], cc[3];
    const int nr = reflect_set(i, m.H, m.reflect, rr), nc = reflect_set(j, m.W, m.reflect, cc);
    for (int a = 0; a < nr; ++a)
      for (int q = 0; q < nc; ++q) *reinterpret_cast<float4*>(base + frame_pixel(m, rr[a], cc[q]) * m.C) = v;
  } else {
    *reinterpret_cast<float4*>(base + frame_pixel(m, i, j) * m.C) = v;
  }
}
void launch_f32_to_frame_f32(const float* src, int B, int H, int W, int Cs, float* dst, const FrameMap& dmap, int Kp, cudaStream_t st) {
  const int64_t tot = int64_t(B) * H * W * (dmap.C / 4);
  f32_to_frame_f32_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(src, B, H, W, Cs, dst, dmap, Kp);
}
// instance-norm statistics of a plain fp32 [B][HW][C] tensor in double precision: stats[b][c] = (sum, sum of squares)
__global__ void __launch_bounds__(256) in_stats_f32_kernel(const float* __restrict__ x, int HW, int C, double* stats, int ppb) {
  __shared__ double sh[8][32][2];
  const int b = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const int pix0 = blockIdx.z * ppb, pix1 = min(HW, pix0 + ppb);
  double s1 = 0.0, s2 = 0.0;
  if (c < C)
    for (int pix = pix0 + ty; pix < pix1; pix += 8) {
      const double v = x[(int64_t(b) * HW + pix) * C + c];
      s1 += v;
      s2 += v * v;
    }
  sh[ty][threadIdx.x][0] = s1;
  sh[ty][threadIdx.x][1] = s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int q = 1; q < 8; ++q) { s1 += sh[q][threadIdx.x][0]; s2 += sh[q][threadIdx.x][1]; }
    atomicAdd(stats + (int64_t(b) * C + c) * 2, s1);
    atomicAdd(stats + (int64_t(b) * C + c) * 2 + 1, s2);
  }
}
__global__ void in_apply_f32_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ res, float* y, int HW, int C,
                                    int64_t n, float eps, int act, float alpha) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = int(i % C);
  const int b = int(i / (int64_t(HW) * C));
  const double mu = stats[(int64_t(b) * C + c) * 2] / HW;
  const double var = fmax(stats[(int64_t(b) * C + c) * 2 + 1] / HW - mu * mu, 0.0);
  const float inv = float(1.0 / sqrt(var + double(eps))) * (gamma ? gamma[c] : 1.f);
  float z = (x[i] - float(mu)) * inv + (beta ? beta[c] : 0.f);
  z = act_fwd(z, act, alpha);
  if (res != nullptr) z += res[i];
  y[i] = z;
}
void launch_instance_norm_f32(const float* x, const float* gamma, const float* beta, const float* res, float* y, int B, int HW,
                              int C, float eps, int act, float alpha, double* stats, cudaStream_t st) {
  cudaMemsetAsync(stats, 0, size_t(B) * C * 2 * sizeof(double), st);
  int splits = (148 * 4) / (B * ((C + 31) / 32));
  splits = splits < 1 ? 1 : (splits > 64 ? 64 : splits);
  const int ppb = (HW + splits - 1) / splits;
  in_stats_f32_kernel<<<dim3((C + 31) / 32, B, (HW + ppb - 1) / ppb), dim3(32, 8), 0, st>>>(x, HW, C, stats, ppb);
  const int64_t n = int64_t(B) * HW * C;
  in_apply_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(x, stats, gamma, beta, res, y, HW, C, n, eps, act, alpha);
}

// ------------------------------------------------------------------------------------------ IN fwd
__global__ void __launch_bounds__(kGlueThreads) in_stats_kernel(const sg_bf16* __restrict__ y, int HW, int C,
                                                                float* stats, int ppb) {
  extern __shared__ float sred[];
  const int b = blockIdx.y;
  const int C8 = C >> 3;
  const int cg = threadIdx.x % C8, lp = threadIdx.x / C8, ppi = kGlueThreads / C8;
  for (int t = threadIdx.x; t < 2 * C; t += kGlueThreads) sred[t] = 0.f;
  __syncthreads();
  float a1[8] = {0}, a2[8] = {0};
  const int pix0 = blockIdx.x * ppb, pix1 = min(HW, pix0 + ppb);
  for (int pix = pix0 + lp; pix < pix1; pix += ppi) {
    float v[8];
    ld_bf16x8(y + (int64_t(b) * HW + pix) * C + cg * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a1[e] += v[e];
      a2[e] += v[e] * v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&sred[(cg * 8 + e) * 2], a1[e]);
    atomicAdd(&sred[(cg * 8 + e) * 2 + 1], a2[e]);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * C; t += kGlueThreads) atomicAdd(stats + int64_t(b) * C * 2 + t, sred[t]);
}
void launch_in_stats(const sg_bf16* y, int B, int HW, int C, float* stats, cudaStream_t st) {
  const int ppb = pix_per_block_for(HW, B);
  dim3 grid((HW + ppb - 1) / ppb, B);
  in_stats_kernel<<<grid, kGlueThreads, 2 * C * sizeof(float), st>>>(y, HW, C, stats, ppb);
}

// (instance-norm forward / backward and the residual-gradient gather live in glue_rows.cu)
__global__ void in_param_grad_kernel(const float* __restrict__ sums, int nb, int C, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float g = 0.f, be = 0.f;
  for (int b = 0; b < nb; ++b) {
    be += sums[(int64_t(b) * C + c) * 2];
    g += sums[(int64_t(b) * C + c) * 2 + 1];
  }
  dgamma[c] = g;
  dbeta[c] = be;
}
void launch_in_param_grad(const float* sums, int nb, int C, float* dgamma, float* dbeta, cudaStream_t st) {
  in_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, nb, C, dgamma, dbeta);
}

// ------------------------------------------------------------------------------------------ gather
__device__ bool grid_sum_ordered(const float* part_sm, int K, const OrderedSum& red, float* out, float* tmp_sm);  // below

__global__ void __launch_bounds__(kGlueThreads) act_bwd_kernel(const ActBwdParams p, int ppb) {
  extern __shared__ float sred[];
  const int b = blockIdx.y;
  const int ba = b < p.nb_act ? b : b - p.act_wrap;
  const int C8 = p.C >> 3;
  const int cg = threadIdx.x % C8, lp = threadIdx.x / C8, ppi = kGlueThreads / C8;
  const int c0 = cg * 8, HW = p.H * p.W;
  for (int t = threadIdx.x; t < p.C; t += kGlueThreads) sred[t] = 0.f;
  __syncthreads();
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int pix0 = blockIdx.x * ppb, pix1 = min(HW, pix0 + ppb);
  // common case (no folded border, bf16 gradient, plain destination frame): two pixels per iteration with all four
  // 16-byte loads issued before any arithmetic -- the generic helpers below cost twice the time on this pass
  const bool simple = p.g.ptr != nullptr && p.g.fold == 0 && !p.g.f32 && !(p.dmap.kind == 0 && p.dmap.reflect > 0);
  if (simple) {
    const sg_bf16* zb = p.Z + int64_t(ba) * p.zmap.frame_pix * p.zmap.C + c0;
    const sg_bf16* gb = reinterpret_cast<const sg_bf16*>(p.g.ptr) + int64_t(b) * p.g.Hs * p.g.Ws * p.C + c0;
    sg_bf16* db = p.dst + int64_t(b) * p.dmap.frame_pix * p.dmap.C + c0;
    for (int pix = pix0 + lp; pix < pix1; pix += 2 * ppi) {
      const int pixB = pix + ppi;
      const bool hasB = pixB < pix1;
      const int iA = pix / p.W, jA = pix - iA * p.W;
      const int iB = hasB ? pixB / p.W : iA, jB = hasB ? pixB - iB * p.W : jA;
      const uint4 zA = __ldg(reinterpret_cast<const uint4*>(zb + frame_pixel(p.zmap, iA, jA) * p.zmap.C));
      const uint4 gA = __ldg(reinterpret_cast<const uint4*>(gb + (int64_t(iA + p.g.oy) * p.g.Ws + (jA + p.g.ox)) * p.C));
      const uint4 zB = __ldg(reinterpret_cast<const uint4*>(zb + frame_pixel(p.zmap, iB, jB) * p.zmap.C));
      const uint4 gB = __ldg(reinterpret_cast<const uint4*>(gb + (int64_t(iB + p.g.oy) * p.g.Ws + (jB + p.g.ox)) * p.C));
      float z[8], d[8];
      bf16x8_to_float(zA, z);
      bf16x8_to_float(gA, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        d[e] = z[e] > 0.f ? d[e] : p.alpha * d[e];
        acc[e] += d[e];
      }
      st_bf16x8(db + frame_pixel(p.dmap, iA, jA) * p.dmap.C, d);
      if (hasB) {
        bf16x8_to_float(zB, z);
        bf16x8_to_float(gB, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          d[e] = z[e] > 0.f ? d[e] : p.alpha * d[e];
          acc[e] += d[e];
        }
        st_bf16x8(db + frame_pixel(p.dmap, iB, jB) * p.dmap.C, d);
      }
    }
  } else
  for (int pix = pix0 + lp; pix < pix1; pix += ppi) {
    const int i = pix / p.W, j = pix - i * p.W;
    float z[8], d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ld_bf16x8(p.Z + (int64_t(ba) * p.zmap.frame_pix + frame_pixel(p.zmap, i, j)) * p.zmap.C + c0, z);
    grad_read8(p.g, b, i, j, p.H, p.W, p.C, c0, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d[e] = z[e] > 0.f ? d[e] : p.alpha * d[e];
      acc[e] += d[e];
    }
    write_frame8(p.dst, p.dmap, b, i, j, c0, d);
  }
  if (p.dbias == nullptr) return;
  // bias gradient: the ppi threads that share a channel group park their sums, thread t < C adds them in order; blocks of
  // images >= nb_bias (the generator's virtual images) deposit zeros so that every block takes part in the ordered sum
  const bool counts = b < p.nb_bias;
  float* slab = sred + p.C;  // [ppi][C]
#pragma unroll
  for (int e = 0; e < 8; ++e) slab[lp * p.C + c0 + e] = counts ? acc[e] : 0.f;
  __syncthreads();
  for (int t = threadIdx.x; t < p.C; t += kGlueThreads) {
    float s = 0.f;
    for (int q = 0; q < ppi; ++q) s += slab[q * p.C + t];
    sred[t] = s;
  }
  __syncthreads();
  if (p.red.scratch != nullptr) {
    grid_sum_ordered(sred, p.C, p.red, p.dbias, slab);
  } else if (counts) {
    for (int t = threadIdx.x; t < p.C; t += kGlueThreads) atomicAdd(p.dbias + t, sred[t]);
  }
}
void launch_act_bwd(const ActBwdParams& p, cudaStream_t st) {
  const int HW = p.H * p.W, ppb = pix_per_block_for(HW, p.B);
  dim3 grid((HW + ppb - 1) / ppb, p.B);
  // shared memory: [C] block sums + [ppi][C] = kGlueThreads * 8 per-thread sums (also >= kGlueThreads floats of scratch)
  act_bwd_kernel<<<grid, kGlueThreads, (p.C + kGlueThreads * 8) * sizeof(float), st>>>(p, ppb);
}
int act_bwd_blocks(int B, int H, int W) {
  const int HW = H * W, ppb = pix_per_block_for(HW, B);
  return ((HW + ppb - 1) / ppb) * B;
}

// ------------------------------------------------------------------------------------------ losses
__device__ __forceinline__ float softplus_negabs(float x) { return log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__device__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// See OrderedSum (glue.h).  part_sm[0..K) = this block's partial values in shared memory (all threads may read them after the
// caller's __syncthreads); the last block accumulates the grid totals into out[0..K).  tmp_sm: blockDim.x floats of
// shared memory.  Called by all threads of every block.
__device__ bool grid_sum_ordered(const float* part_sm, int K, const OrderedSum& red, float* out, float* tmp_sm) {
  const int nblk = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  __shared__ unsigned int s_last;
  for (int k = threadIdx.x; k < K; k += blockDim.x) __stcg(red.scratch + size_t(bid) * K + k, part_sm[k]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicInc(red.ticket, unsigned(nblk - 1)) == unsigned(nblk - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  // thread (g, k): deposits g, g + groups, ... of value k, added in that order; then the groups in order
  for (int k0 = 0; k0 < K; k0 += blockDim.x) {
    const int kn = min(K - k0, int(blockDim.x));
    const int groups = blockDim.x / kn;
    const int g = threadIdx.x / kn, k = threadIdx.x - g * kn;
    float acc = 0.f;
    if (g < groups)
      for (int b = g; b < nblk; b += groups) acc += __ldcg(red.scratch + size_t(b) * K + k0 + k);
    __syncthreads();
    if (g < groups) tmp_sm[g * kn + k] = acc;
    __syncthreads();
    if (int(threadIdx.x) < kn) {
      float t = 0.f;
      for (int q = 0; q < groups; ++q) t += tmp_sm[q * kn + threadIdx.x];
      out[k0 + threadIdx.x] += t;
    }
  }
  __syncthreads();
  return true;  // uniform over the block: this is the block that holds the totals
}

// The logits are KB-sized (B x 5 x 13 at 256x512): two small multi-block launches.
__global__ void disc_logits_kernel(const DiscLossParams p) {
  const int Ho = max(p.Hd, p.hm), Wo = max(p.Wd, p.wm);
  const int npos = Ho * Wo;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * p.B * npos) return;
  const int b = idx / npos, r = idx - b * npos, I = r / Wo, J = r - I * Wo;
  const int ih = p.Hd == 1 ? 0 : I, jh = p.Wd == 1 ? 0 : J, im = p.hm == 1 ? 0 : I, jm = p.wm == 1 ? 0 : J;
  const float* h = p.h4 + ((int64_t(b) * p.Hd + ih) * p.Wd + jh) * p.Cs;
  const float* mk = p.mask + ((int64_t(b % p.B) * p.hm + im) * p.wm + jm) * p.Cs;
  float s = 0.f;
  for (int c = 0; c < p.Cs; ++c) s += h[c] * mk[c];
  p.logits[idx] = s;
}
__global__ void __launch_bounds__(256) disc_loss_grad_kernel(const DiscLossParams p) {
  __shared__ float sh[32];
  __shared__ float tmp[256];
  extern __shared__ float sbias[];  // [groups][Cs] partial bias gradients, [2 + Cs] block partials, [2 + Cs] grid totals
  const int Ho = max(p.Hd, p.hm), Wo = max(p.Wd, p.wm);
  const int B2 = 2 * p.B, npos = Ho * Wo;
  const float N = float(p.B) * npos;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  // losses
  float lg = 0.f, ld = 0.f;
  for (int idx = gtid; idx < B2 * npos; idx += gstride) {
    const int b = idx / npos;
    const float x = p.logits[idx];
    if (p.lsgan) {
      if (b < p.B) ld += (x - 1.f) * (x - 1.f);
      else { ld += x * x; lg += (x - 1.f) * (x - 1.f); }
    } else {
      const float sp = softplus_negabs(x), mx = fmaxf(x, 0.f);
      if (b < p.B) ld += mx - x + sp;           // label 1
      else { ld += mx + sp; lg += mx - x + sp; }  // fake: label 0 for D, label 1 for G
    }
  }
  lg = block_sum(lg, sh);
  ld = block_sum(ld, sh);
  // d logits -> d h4 (3B virtual images: real-D, fake-D, fake-G) and the bias gradient (first 2B).  Thread (g, c) keeps its
  // channel and walks the positions g, g + G * gridDim.x, ..., so its bias partial needs no atomics.
  const int G = blockDim.x / p.Cs;  // position groups per block (7 at Cs = 34)
  const int g = threadIdx.x / p.Cs, c = threadIdx.x - g * p.Cs;
  const int npos_h = 3 * p.B * p.Hd * p.Wd;
  float bsum = 0.f;
  if (g < G)
    for (int pos = blockIdx.x * G + g; pos < npos_h; pos += gridDim.x * G) {
      int r = pos;
      const int jh = r % p.Wd;
      r /= p.Wd;
      const int ih = r % p.Hd, v = r / p.Hd;
      const int bs = v < B2 ? v : v - p.B;
      const float label = (v < p.B || v >= B2) ? 1.f : 0.f;
      const float sc = (v < B2 ? p.disc_scale : 1.f) / N;
      const int I0 = p.Hd == 1 ? 0 : ih, I1 = p.Hd == 1 ? Ho : ih + 1;
      const int J0 = p.Wd == 1 ? 0 : jh, J1 = p.Wd == 1 ? Wo : jh + 1;
      float gv = 0.f;
      for (int I = I0; I < I1; ++I)
        for (int J = J0; J < J1; ++J) {
          const int im = p.hm == 1 ? 0 : I, jm = p.wm == 1 ? 0 : J;
          const float x = p.logits[(bs * Ho + I) * Wo + J];
          const float dl = (p.lsgan ? 2.f * (x - label) : (sigmoidf(x) - label)) * sc;
          gv += dl * p.mask[((int64_t(bs % p.B) * p.hm + im) * p.wm + jm) * p.Cs + c];
        }
      reinterpret_cast<__nv_bfloat16*>(p.dst)[(int64_t(v) * p.dmap.frame_pix + frame_pixel(p.dmap, ih, jh)) * p.dmap.C + c] =
          __float2bfloat16_rn(gv);
      if (v < B2) bsum += gv;
    }
  if (g < G) sbias[g * p.Cs + c] = bsum;
  __syncthreads();
  float bc = 0.f;
  if (int(threadIdx.x) < p.Cs)
    for (int q = 0; q < G; ++q) bc += sbias[q * p.Cs + threadIdx.x];
  __syncthreads();
  float* part = sbias + G * p.Cs;  // [2 + Cs]
  if (int(threadIdx.x) < p.Cs) part[2 + threadIdx.x] = bc;
  if (threadIdx.x == 0) { part[0] = lg / N; part[1] = ld * p.disc_scale / N; }
  __syncthreads();
  if (p.red.scratch != nullptr) {
    // totals: loss[0], loss[1], dbias[0..Cs) -- two destinations, so the last block adds into a staging row first
    float* tot = part + 2 + p.Cs;
    for (int t = threadIdx.x; t < 2 + p.Cs; t += blockDim.x) tot[t] = 0.f;
    __syncthreads();
    if (grid_sum_ordered(part, 2 + p.Cs, p.red, tot, tmp)) {
      if (threadIdx.x < 2) p.loss[threadIdx.x] += tot[threadIdx.x];
      if (int(threadIdx.x) < p.Cs) p.dbias[threadIdx.x] += tot[2 + threadIdx.x];
    }
  } else {
    if (threadIdx.x < 2) atomicAdd(p.loss + threadIdx.x, part[threadIdx.x]);
    if (int(threadIdx.x) < p.Cs && part[2 + threadIdx.x] != 0.f) atomicAdd(p.dbias + threadIdx.x, part[2 + threadIdx.x]);
  }
}
void launch_disc_loss(const DiscLossParams& p, cudaStream_t st) {
  const int Ho = p.Hd > p.hm ? p.Hd : p.hm, Wo = p.Wd > p.wm ? p.Wd : p.wm;
  const int nlog = 2 * p.B * Ho * Wo;
  disc_logits_kernel<<<(nlog + 127) / 128, 128, 0, st>>>(p);
  const int G = 256 / p.Cs;
  const int npos_h = 3 * p.B * p.Hd * p.Wd;
  int blocks = (npos_h + G - 1) / G;
  if (blocks > 148) blocks = 148;
  if (blocks < 1) blocks = 1;
  disc_loss_grad_kernel<<<blocks, 256, (G * p.Cs + 2 * (2 + p.Cs)) * sizeof(float), st>>>(p);
}
int disc_loss_blocks(int B, int Hd, int Wd, int Cs) {
  const int G = 256 / Cs;
  int blocks = (3 * B * Hd * Wd + G - 1) / G;
  return blocks > 148 ? 148 : (blocks < 1 ? 1 : blocks);
}

__global__ void mask_reduce_kernel(const float* __restrict__ x, const float* __restrict__ mask, int B, int Hd, int Wd,
                                   int hm, int wm, int Cs, float* out) {
  const int Ho = max(Hd, hm), Wo = max(Wd, wm);
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(B) * Ho * Wo) return;
  const int b = int(idx / (Ho * Wo)), r = int(idx - int64_t(b) * Ho * Wo), I = r / Wo, J = r - I * Wo;
  const int ih = Hd == 1 ? 0 : I, jh = Wd == 1 ? 0 : J, im = hm == 1 ? 0 : I, jm = wm == 1 ? 0 : J;
  const float* h = x + ((int64_t(b) * Hd + ih) * Wd + jh) * Cs;
  const float* mk = mask + ((int64_t(b) * hm + im) * wm + jm) * Cs;
  float s = 0.f;
  for (int c = 0; c < Cs; ++c) s += h[c] * mk[c];
  out[idx] = s;
}
void launch_mask_reduce(const float* x, const float* mask, int B, int Hd, int Wd, int hm, int wm, int Cs, float* out,
                        cudaStream_t st) {
  const int Ho = Hd > hm ? Hd : hm, Wo = Wd > wm ? Wd : wm;
  const int64_t tot = int64_t(B) * Ho * Wo;
  mask_reduce_kernel<<<unsigned((tot + 127) / 128), 128, 0, st>>>(x, mask, B, Hd, Wd, hm, wm, Cs, out);
}

// L1 sign gradient + discriminator / gradient-loss gradients, times tanh', into the 8-channel bf16 seed frame of the output
// convolution; sum |diff| and the bias gradient on the way.  VEC: four pixels per thread, every stream read with three
// 128-bit loads (W % 4 == 0, 16-byte aligned sources); else one pixel per thread.
template <bool VEC>
__global__ void __launch_bounds__(256) fake_grad_kernel(const FakeGradParams p) {
  __shared__ float sh[32];
  constexpr int PX = VEC ? 4 : 1;
  const int64_t tot = int64_t(p.B) * p.H * p.W;
  const float inv_n = 1.f / (float(tot) * 3.f);
  const float lw = p.l1_weight * inv_n;
  float l1 = 0.f, db[3] = {0.f, 0.f, 0.f};
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t * PX < tot; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t idx = t * PX;
    const int b = int(idx / (int64_t(p.H) * p.W));
    const int r = int(idx - int64_t(b) * p.H * p.W);
    const int i = r / p.W, j = r - i * p.W;
    float f[3 * PX], tg[3 * PX], g[3 * PX];
    if (VEC) {
      const float4* f4 = reinterpret_cast<const float4*>(p.fake + idx * 3);
      const float4* t4 = reinterpret_cast<const float4*>(p.target + idx * 3);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 a = __ldg(f4 + q), c = __ldg(t4 + q);
        f[4 * q] = a.x; f[4 * q + 1] = a.y; f[4 * q + 2] = a.z; f[4 * q + 3] = a.w;
        tg[4 * q] = c.x; tg[4 * q + 1] = c.y; tg[4 * q + 2] = c.z; tg[4 * q + 3] = c.w;
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) g[e] = 0.f;
      if (p.dD != nullptr) {
        const float4* d4 = reinterpret_cast<const float4*>(p.dD + idx * 3);
#pragma unroll
        for (int q = 0; q < 3; ++q) { const float4 a = __ldg(d4 + q); g[4 * q] += a.x; g[4 * q + 1] += a.y; g[4 * q + 2] += a.z; g[4 * q + 3] += a.w; }
      }
      if (p.dG != nullptr) {
        const float4* d4 = reinterpret_cast<const float4*>(p.dG + idx * 3);
#pragma unroll
        for (int q = 0; q < 3; ++q) { const float4 a = __ldg(d4 + q); g[4 * q] += a.x; g[4 * q + 1] += a.y; g[4 * q + 2] += a.z; g[4 * q + 3] += a.w; }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        f[c] = p.fake[idx * 3 + c];
        tg[c] = p.target[idx * 3 + c];
        g[c] = (p.dD != nullptr ? p.dD[idx * 3 + c] : 0.f) + (p.dG != nullptr ? p.dG[idx * 3 + c] : 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int e = 3 * q + c;
        const float diff = f[e] - tg[e];
        l1 += fabsf(diff);
        // the same order of operations as before: L1 sign term first, then the two gradient streams
        float gg = lw * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) + g[e];
        gg *= (1.f - f[e] * f[e]);
        v[c] = gg;
        db[c] += gg;
      }
      write_frame8(p.dst, p.dmap, b, i, j + q, 0, v);
    }
  }
  __shared__ float part[4], gsum[4], tmp[256];
  l1 = block_sum(l1, sh);
  if (threadIdx.x == 0) part[0] = l1;
  for (int c = 0; c < 3; ++c) {
    const float s = block_sum(db[c], sh);
    if (threadIdx.x == 0) part[1 + c] = s;
  }
  if (threadIdx.x < 4) gsum[threadIdx.x] = 0.f;
  __syncthreads();
  if (p.red.scratch != nullptr) {
    if (grid_sum_ordered(part, 4, p.red, gsum, tmp)) {
      if (threadIdx.x == 0) p.loss[2] += gsum[0];
      if (threadIdx.x < 3) p.dbias[threadIdx.x] += gsum[1 + threadIdx.x];
    }
  } else if (threadIdx.x == 0) {
    atomicAdd(p.loss + 2, part[0]);
    for (int c = 0; c < 3; ++c) atomicAdd(p.dbias + c, part[1 + c]);
  }
}
void launch_fake_grad(const FakeGradParams& p, cudaStream_t st) {
  const int64_t tot = int64_t(p.B) * p.H * p.W;
  const uintptr_t al = reinterpret_cast<uintptr_t>(p.fake) | reinterpret_cast<uintptr_t>(p.target) |
                       reinterpret_cast<uintptr_t>(p.dD) | reinterpret_cast<uintptr_t>(p.dG);
  const bool vec = (p.W % 4 == 0) && (al & 15) == 0;
  const int64_t nthr = vec ? tot / 4 : tot;
  int blocks = int((nthr + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (vec) fake_grad_kernel<true><<<blocks, 256, 0, st>>>(p);
  else fake_grad_kernel<false><<<blocks, 256, 0, st>>>(p);
}
size_t ordered_sum_scratch_floats(int B, int H, int W, int Cs, int C_h0, int H_h0, int W_h0, int nb_h0) {
  size_t n = size_t(148 * 8) * 4;                                            // fake_grad: at most 148 * 8 blocks x 4 values
  n = std::max(n, size_t((W + 63) / 64) * ((H + 15) / 16) * B);             // gradloss: one value per tile
  n = std::max(n, size_t(148) * (2 + Cs));                                   // disc_loss_grad
  n = std::max(n, size_t(act_bwd_blocks(nb_h0, H_h0, W_h0)) * C_h0);         // act_bwd of D's first layer
  return n + 64;
}

__global__ void finalize_losses_kernel(const float* loss, float l1_weight, float n_l1, float lg_weight, float* out) {
  out[0] = loss[0] + l1_weight * loss[2] / n_l1 + lg_weight * loss[3];
  out[1] = loss[1];
}
void launch_finalize_losses(const float* loss, float l1_weight, float n_l1, float lg_weight, float* out,
                            cudaStream_t st) {
  finalize_losses_kernel<<<1, 1, 0, st>>>(loss, l1_weight, n_l1, lg_weight, out);
}

// ---- SG-GAN criteria ------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

__global__ void seg_edge_weight_kernel(const float* __restrict__ seg, int B, int H, int W, float* w) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(B) * H * W) return;
  const int b = int(idx / (int64_t(H) * W));
  const int r = int(idx - int64_t(b) * H * W), i = r / W, j = r - i * W;
  const float* s = seg + int64_t(b) * H * W * 3;
  float acc = 0.f;
  const int jl = reflect_idx(j - 1, W), jr = reflect_idx(j + 1, W), iu = reflect_idx(i - 1, H), id = reflect_idx(i + 1, H);
  for (int c = 0; c < 3; ++c) {
    acc += fabsf(s[(int64_t(i) * W + jr) * 3 + c] - s[(int64_t(i) * W + jl) * 3 + c]);
    acc += fabsf(s[(int64_t(id) * W + j) * 3 + c] - s[(int64_t(iu) * W + j) * 3 + c]);
  }
  w[idx] = acc > 0.f ? 1.f : 0.f;
}
void launch_seg_edge_weight(const float* seg, int B, int H, int W, float* weight, cudaStream_t st) {
  const int64_t tot = int64_t(B) * H * W;
  seg_edge_weight_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(seg, B, H, W, weight);
}

// ---- gradient-sensitive loss (module.py:325-351) as ONE shared-memory-tiled pass ------------------------------------
// loss = mean_{b,i,j} w[i,j] * mean_{c, d in {x,y}} | |S_d in|[i,j,c] - |S_d tgt|[i,j,c] |, S = 3x3 Sobel with SAME zero
// padding; d loss / d in[u,v,c] = sum over the 3x3 positions (i,j) whose window covers (u,v) of
// w * sgn(e_d) * sgn(S_d in) * K_d[u-i+1][v-j+1].  A block owns a 16 x 64 tile of pixels:
//   phase 1  `in` and `tgt` with a 2-pixel halo -> shared memory (coalesced rows, zeros outside the image);
//   phase 2  both Sobel responses at every position of the tile + 1-pixel halo, out of shared memory: the loss terms of
//            the tile's own positions and the six coefficients  w * sgn(e_d) * sgn(S_d in)  per position -> shared memory;
//   phase 3  every pixel of the tile gathers its 9 x 2 x 3 coefficient taps and writes d_in.
// Each input byte is read from HBM once (+ halo re-reads out of L2): (2 x 12 + 4) B read, 12 B written per pixel; the
// naive form evaluated ~160 Sobel windows per pixel from global memory (233 us at 8 x 256 x 512; this one: profiles/).
constexpr int kGlTh = 16, kGlTw = 64;
constexpr int kGlIw = kGlTw + 4, kGlIh = kGlTh + 4;  // image tile with halo 2
constexpr int kGlCw = kGlTw + 2, kGlCh = kGlTh + 2;  // coefficient tile with halo 1
__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ void sobel_smem(const float* t, int r, int q, int c, float* gx, float* gy) {
  // t: [kGlIh][kGlIw][3], (r, q) = position in halo-2 coordinates of the window CENTRE
  const float* p0 = t + ((r - 1) * kGlIw + (q - 1)) * 3 + c;
  const float* p1 = p0 + kGlIw * 3;
  const float* p2 = p1 + kGlIw * 3;
  const float a00 = p0[0], a01 = p0[3], a02 = p0[6], a10 = p1[0], a12 = p1[6], a20 = p2[0], a21 = p2[3], a22 = p2[6];
  *gx = (a02 - a00) + 2.f * (a12 - a10) + (a22 - a20);
  *gy = (a20 - a00) + 2.f * (a21 - a01) + (a22 - a02);
}
__global__ void __launch_bounds__(256) gradloss_kernel(const float* __restrict__ in, const float* __restrict__ tgt,
                                                       const float* __restrict__ wgt, int B, int H, int W,
                                                       float scale, float* loss_slot, float* d_in, OrderedSum red) {
  extern __shared__ float gl_smem[];
  float* s_in = gl_smem;                          // [kGlIh][kGlIw][3]
  float* s_tg = s_in + kGlIh * kGlIw * 3;
  float* s_cf = s_tg + kGlIh * kGlIw * 3;         // [kGlCh][kGlCw][c][d]
  __shared__ float sh[32];
  const int b = blockIdx.z, i0 = blockIdx.y * kGlTh, j0 = blockIdx.x * kGlTw;
  const float* ib = in + int64_t(b) * H * W * 3;
  const float* tb = tgt + int64_t(b) * H * W * 3;
  const float* wb = wgt + int64_t(b) * H * W;
  const float inv = 1.f / (float(int64_t(B) * H * W) * 6.f);
  // ---- phase 1
  for (int e = threadIdx.x; e < kGlIh * kGlIw * 3; e += 256) {
    const int r = e / (kGlIw * 3), x = e - r * (kGlIw * 3);
    const int i = i0 - 2 + r, j = j0 - 2 + x / 3;
    const bool ok = i >= 0 && i < H && j >= 0 && j < W;
    const int64_t g = (int64_t(i) * W + (j0 - 2)) * 3 + x;
    s_in[e] = ok ? __ldg(ib + g) : 0.f;
    s_tg[e] = ok ? __ldg(tb + g) : 0.f;
  }
  __syncthreads();
  // ---- phase 2
  float lsum = 0.f;
  for (int e = threadIdx.x; e < kGlCh * kGlCw; e += 256) {
    const int r = e / kGlCw, q = e - r * kGlCw;
    const int i = i0 - 1 + r, j = j0 - 1 + q;
    const bool ok = i >= 0 && i < H && j >= 0 && j < W;
    const float w = ok ? __ldg(wb + int64_t(i) * W + j) : 0.f;
    const bool own = ok && r >= 1 && r <= kGlTh && q >= 1 && q <= kGlTw;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gxi, gyi, gxt, gyt;
      sobel_smem(s_in, r + 1, q + 1, c, &gxi, &gyi);
      sobel_smem(s_tg, r + 1, q + 1, c, &gxt, &gyt);
      const float ex = fabsf(gxi) - fabsf(gxt), ey = fabsf(gyi) - fabsf(gyt);
      if (own) lsum += w * (fabsf(ex) + fabsf(ey));
      s_cf[e * 6 + c * 2] = w * sgnf(ex) * sgnf(gxi);
      s_cf[e * 6 + c * 2 + 1] = w * sgnf(ey) * sgnf(gyi);
    }
  }
  __syncthreads();
  // ---- phase 3
  if (d_in != nullptr) {
    const float KX[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
    const float KY[3][3] = {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}};
    for (int e = threadIdx.x; e < kGlTh * kGlTw * 3; e += 256) {
      const int r = e / (kGlTw * 3), x = e - r * (kGlTw * 3), q = x / 3, c = x - q * 3;
      const int u = i0 + r, v = j0 + q;
      if (u >= H || v >= W) continue;
      float g = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          // window centre (i, j) = (u - (a - 1), v - (t - 1)) -> coefficient-tile coordinates (r + 1 - (a - 1), q + 1 - (t - 1))
          const float* cf = s_cf + ((r + 2 - a) * kGlCw + (q + 2 - t)) * 6 + c * 2;
          g += cf[0] * KX[a][t] + cf[1] * KY[a][t];
        }
      d_in[(int64_t(b) * H * W + int64_t(u) * W + j0) * 3 + x] = g * inv * scale;
    }
  }
  lsum = block_sum(lsum, sh);
  if (loss_slot == nullptr) return;
  if (red.scratch != nullptr) {
    __shared__ float part[1], tmp[256];
    if (threadIdx.x == 0) part[0] = lsum * inv;
    __syncthreads();
    grid_sum_ordered(part, 1, red, loss_slot, tmp);
  } else if (threadIdx.x == 0) {
    atomicAdd(loss_slot, lsum * inv);
  }
}
void launch_gradloss(const float* in, const float* target, const float* weight, int B, int H, int W, float scale,
                     float* loss_slot, float* d_in, cudaStream_t st, OrderedSum red) {
  dim3 grid((W + kGlTw - 1) / kGlTw, (H + kGlTh - 1) / kGlTh, B);
  constexpr size_t smem = size_t(2 * kGlIh * kGlIw * 3 + kGlCh * kGlCw * 6) * sizeof(float);  // 61 KB: three blocks per SM
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gradloss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    attr_set = true;
  }
  gradloss_kernel<<<grid, 256, smem, st>>>(in, target, weight, B, H, W, scale, loss_slot, d_in, red);
}

// module.tf_deriv (module.py:325-334) as a standalone operator: Sobel x / y per channel, SAME zero padding or VALID;
// out[b, i, j, c * 2 + d].  Any channel count; one thread per output (pixel, channel), rows read coalesced.
__global__ void __launch_bounds__(256) sobel_deriv_kernel(const float* __restrict__ x, int B, int H, int W, int C, int valid,
                                                          float* __restrict__ out) {
  const int Ho = valid ? H - 2 : H, Wo = valid ? W - 2 : W;
  const int64_t tot = int64_t(B) * Ho * Wo * C;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < tot; idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const int64_t pix = idx / C;
    const int j = int(pix % Wo), i = int((pix / Wo) % Ho), b = int(pix / (int64_t(Wo) * Ho));
    const float* xb = x + int64_t(b) * H * W * C + c;
    float v[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int ii = i + a - (valid ? 0 : 1), jj = j + q - (valid ? 0 : 1);
        v[a][q] = (ii >= 0 && ii < H && jj >= 0 && jj < W) ? __ldg(xb + (int64_t(ii) * W + jj) * C) : 0.f;
      }
    out[idx * 2] = (v[0][2] - v[0][0]) + 2.f * (v[1][2] - v[1][0]) + (v[2][2] - v[2][0]);
    out[idx * 2 + 1] = (v[2][0] - v[0][0]) + 2.f * (v[2][1] - v[0][1]) + (v[2][2] - v[0][2]);
  }
}
void launch_sobel_deriv(const float* x, int B, int H, int W, int C, int valid, float* out, cudaStream_t st) {
  const int64_t tot = int64_t(B) * (valid ? H - 2 : H) * (valid ? W - 2 : W) * C;
  int blocks = int((tot + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  sobel_deriv_kernel<<<blocks, 256, 0, st>>>(x, B, H, W, C, valid, out);
}

__global__ void __launch_bounds__(256) criterion_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        int64_t n, int mode, float* out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float x = a[i], y = b[i];
    if (mode == 0) s += fabsf(x - y);
    else if (mode == 1) s += (x - y) * (x - y);
    else s += fmaxf(x, 0.f) - x * y + softplus_negabs(x);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s / float(n));
}
void launch_criterion(const float* a, const float* b, int64_t n, int mode, float* out, cudaStream_t st) {
  int blocks = int((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaMemsetAsync(out, 0, sizeof(float), st);
  criterion_kernel<<<blocks, 256, 0, st>>>(a, b, n, mode, out);
}

// ------------------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float alpha_t, float beta1, float beta2, float eps, float gscale,
                                                   const long long* __restrict__ step_dev, float lr) {
  if (step_dev != nullptr) {
    // the number of completed steps lives on the device so that a captured CUDA graph of the step stays valid from
    // one replay to the next: alpha_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), t = completed + 1 (Keras, A.8)
    __shared__ float s_alpha;
    if (threadIdx.x == 0) {
      const double t = double(*step_dev + 1);
      s_alpha = float(double(lr) * sqrt(1.0 - pow(double(beta2), t)) / (1.0 - pow(double(beta1), t)));
    }
    __syncthreads();
    alpha_t = s_alpha;
  }
  const int64_t n4 = n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i],
           vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = G[e] * gscale;
      M[e] += (ge - M[e]) * (1.f - beta1);
      V[e] += (ge * ge - V[e]) * (1.f - beta2);
      P[e] -= alpha_t * M[e] / (sqrtf(V[e]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float ge = g[i] * gscale;
      m[i] += (ge - m[i]) * (1.f - beta1);
      v[i] += (ge * ge - v[i]) * (1.f - beta2);
      p[i] -= alpha_t * m[i] / (sqrtf(v[i]) + eps);
    }
  }
}
void launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float alpha_t, float beta1, float beta2,
                 float eps, float gscale, cudaStream_t st, const long long* step_dev, float lr) {
  adam_kernel<<<148 * 8, 256, 0, st>>>(p, g, m, v, n, alpha_t, beta1, beta2, eps, gscale, step_dev, lr);
}
__global__ void bump_step_kernel(long long* step_dev) { *step_dev += 1; }
void launch_bump_step(long long* step_dev, cudaStream_t st) { bump_step_kernel<<<1, 1, 0, st>>>(step_dev); }

// ------------------------------------------------------------------------------------------ packing
// Source layouts (Keras): conv kernel [KH][KW][Cin][Cout]; transposed-conv kernel [KH][KW][Cout][Cin]
// where Cin is the layer INPUT channel count in both cases.  Destination: bf16 slabs [T][N][K].
//  mode 0 conv fwd      t=(kh,kw)  n=co  k=ci
//  mode 1 conv dgrad    t=(kh,kw)  n=ci  k=co
//  mode 2 deconv fwd    t=(kh,kw)  n=co  k=ci
//  mode 3 deconv dgrad  t=(kh,kw)  n=ci  k=co
//  mode 4 window fwd, stride 1 (generator c1):   t=kh        n=co  k=q*8+ci, kw=q
//  mode 5 window dgrad (generator output conv):  t=kh        n=ci  k=q*8+co, kw=KW-1-q
//  mode 6 window fwd, stride 2 (disc. h0):       t=kh*2+bp   n=co  k=q*8+ci, kw=2q+bp
//  mode 7 shift-sum fwd (generator output conv): t=kh        n=kw*4+co  k=ci
__device__ __forceinline__ float pack_value(const PackParams& p, int t, int n, int k) {
  float v = 0.f;
  const int Ci = p.Cin, Co = p.Cout;
  switch (p.mode) {
    case 0: if (n < Co && k < Ci) v = p.src[(int64_t(t) * Ci + k) * Co + n]; break;
    case 1: if (n < Ci && k < Co) v = p.src[(int64_t(t) * Ci + n) * Co + k]; break;
    case 2: if (n < Co && k < Ci) v = p.src[(int64_t(t) * Co + n) * Ci + k]; break;
    case 3: if (n < Ci && k < Co) v = p.src[(int64_t(t) * Co + k) * Ci + n]; break;
    case 4: { const int q = k >> 3, ci = k & 7;
      if (q < p.KW && ci < Ci && n < Co) v = p.src[((int64_t(t) * p.KW + q) * Ci + ci) * Co + n]; break; }
    case 5: { const int q = k >> 3, co = k & 7;
      if (q < p.KW && co < Co && n < Ci) v = p.src[((int64_t(t) * p.KW + (p.KW - 1 - q)) * Ci + n) * Co + co]; break; }
    case 7: { const int kw = n >> 2, co = n & 3;  // shift-sum output conv: t = kh, n = kw*4 + co, k = ci
      if (kw < p.KW && co < Co && k < Ci) v = p.src[((int64_t(t) * p.KW + kw) * Ci + k) * Co + co]; break; }
    case 6: { const int kh = t >> 1, bp = t & 1, q = k >> 3, ci = k & 7, kw = 2 * q + bp;
      if (kw < p.KW && ci < Ci && n < Co) v = p.src[((int64_t(kh) * p.KW + kw) * Ci + ci) * Co + n]; break; }
  }
  return v;
}
// one thread packs 8 consecutive k (K is a multiple of 64) and writes them with one 128-bit store
__device__ __forceinline__ void pack_one(const PackParams& p, int64_t idx8) {
  const int64_t tot8 = (int64_t(p.T) * p.N * p.K) >> 3;
  if (idx8 >= tot8) return;
  const int K8 = p.K >> 3;
  const int k0 = int(idx8 % K8) * 8;
  const int64_t tn = idx8 / K8;
  const int n = int(tn % p.N), t = int(tn / p.N);
  uint4 w;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
  for (int e = 0; e < 4; ++e)
    h[e] = __floats2bfloat162_rn(pack_value(p, t, n, k0 + 2 * e), pack_value(p, t, n, k0 + 2 * e + 1));
  reinterpret_cast<uint4*>(p.dst)[idx8] = w;
}
__global__ void pack_weights_kernel(const PackParams p) { pack_one(p, int64_t(blockIdx.x) * blockDim.x + threadIdx.x); }
// the same slabs as fp32 elements rounded to tf32 (dst is float [T][N][K]); K = groups x Kp, see f32_to_frame_f32_kernel
__global__ void pack_weights_f32_kernel(const PackParams p, float* dst, int Kp) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t tot = int64_t(p.T) * p.N * p.K;
  if (idx >= tot) return;
  const int k = int(idx % p.K), grp = k / Kp, kk = k - grp * Kp;
  const int64_t tn = idx / p.K;
  dst[idx] = tf32_part(pack_value(p, int(tn / p.N), int(tn % p.N), kk), grp == 2 ? 1 : 0);
}
void launch_pack_weights_f32(const PackParams& p, float* dst, int Kp, cudaStream_t st) {
  const int64_t tot = int64_t(p.T) * p.N * p.K;
  pack_weights_f32_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(p, dst, Kp);
}
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const PackParams* __restrict__ jobs,
                                                                 const int* __restrict__ starts, int njobs) {
  __shared__ PackParams sp;
  __shared__ int sj;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;  // last job whose first block is <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (starts[mid] <= int(blockIdx.x)) lo = mid; else hi = mid - 1;
    }
    sj = lo;
    sp = jobs[lo];
  }
  __syncthreads();
  pack_one(sp, int64_t(int(blockIdx.x) - starts[sj]) * 256 + threadIdx.x);  // 256 threads x 8 elements per block
}
void launch_pack_weights_batch(const PackParams* jobs, const int* starts, int njobs, int total_blocks, cudaStream_t st) {
  pack_weights_batch_kernel<<<total_blocks, 256, 0, st>>>(jobs, starts, njobs);
}
void launch_pack_weights(const PackParams& p, cudaStream_t st) {
  const int64_t tot8 = (int64_t(p.T) * p.N * p.K) >> 3;
  pack_weights_kernel<<<unsigned((tot8 + 255) / 256), 256, 0, st>>>(p);
}

// scratch [pair][128 rows][ncol] from the window wgrad launches -> Keras [KH][KW][Cin][Cout] (+=)
//  mode 0 (c1):  row = h*64 + q*8 + ci, col = co            kh = 2*pair+h, kw = q
//  mode 1 (out): row = h*64 + ci,       col = q*8 + co      kh = 2*pair+h, kw = KW-1-q
//  mode 2 (h0):  row = h*64 + q*8 + ci, col = co            group = 2*pair+h = kh*2+bp, kw = 2q+bp
__global__ void unpack_wgrad_kernel(const float* __restrict__ s, float* dW, int mode, int KH, int KW, int Cin, int Cout,
                                    int ncol) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int tot = KH * KW * Cin * Cout;
  if (idx >= tot) return;
  const int co = idx % Cout, ci = (idx / Cout) % Cin, kw = (idx / (Cout * Cin)) % KW, kh = idx / (Cout * Cin * KW);
  int pair, row, col;
  if (mode == 0) { pair = kh >> 1; row = (kh & 1) * 64 + kw * 8 + ci; col = co; }
  else if (mode == 1) { pair = kh >> 1; row = (kh & 1) * 64 + ci; col = (KW - 1 - kw) * 8 + co; }
  else { const int grp = kh * 2 + (kw & 1); pair = grp >> 1; row = (grp & 1) * 64 + (kw >> 1) * 8 + ci; col = co; }
  dW[idx] += s[(int64_t(pair) * 128 + row) * ncol + col];
}
void launch_unpack_wgrad(const float* scratch, float* dW, int mode, int KH, int KW, int Cin, int Cout, int ncol,
                         cudaStream_t st) {
  const int tot = KH * KW * Cin * Cout;
  unpack_wgrad_kernel<<<(tot + 255) / 256, 256, 0, st>>>(scratch, dW, mode, KH, KW, Cin, Cout, ncol);
}

}  // namespace sggan
