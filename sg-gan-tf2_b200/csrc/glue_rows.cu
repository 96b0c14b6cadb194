// glue_rows.cu -- the per-layer elementwise passes of the step, organised by image ROWS:
//
//   in_apply      instance norm + activation (+ residual) of a raw conv output -> next frame
//   in_bwd        its backward (reduce pass and apply pass)
//   grad_gather   residual-stream gradient accumulation
//
// A block owns a few whole image rows; a thread owns 8 channels (one 128-bit access) and walks the
// row's columns with a fixed stride.  Row / column bookkeeping (reflected border rows, phase planes,
// folded gradient rows) is resolved once per row into base pointers, so the per-pixel work is a handful
// of address adds around the 128-bit loads and stores -- these passes were instruction-bound, not
// HBM-bound, when every pixel recomputed its (i, j) with integer divisions and 64-bit frame arithmetic.
// All primary loads of a batch of U pixels are issued before any of them is used.
#include <cuda_bf16.h>

#include "glue.h"

namespace sggan {

constexpr int kRowThreads = 256;

__device__ __forceinline__ float rw_act_fwd(float v, int act, float a) {
  if (act == SG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SG_ACT_LRELU) return v > 0.f ? v : a * v;
  if (act == SG_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float rw_act_grad(float zpre, int act, float a) {
  if (act == SG_ACT_RELU) return zpre > 0.f ? 1.f : 0.f;
  if (act == SG_ACT_LRELU) return zpre > 0.f ? 1.f : a;
  return 1.f;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
  return u;
}
// mirror images of index i inside a reflect border of width p (excluding i itself)
__device__ __forceinline__ int mirrors(int i, int n, int p, int* out) {
  int k = 0;
  if (p > 0) {
    if (i >= 1 && i <= p) out[k++] = -i;
    if (i >= n - 1 - p && i <= n - 2) out[k++] = 2 * (n - 1) - i;
  }
  return k;
}

// ---- destination row of a frame -----------------------------------------------------------------------
struct DstRow {
  sg_bf16* base[3];  // kind 0: row i and its reflected copies, at logical column 0; kind 1: [0] even, [1] odd columns
  int n;
};
__device__ __forceinline__ void dst_row_init(DstRow& r, sg_bf16* dst, const FrameMap& m, int b, int i, int c0) {
  sg_bf16* img = dst + int64_t(b) * m.frame_pix * m.C + c0;
  if (m.kind == 0) {
    int mr[2];
    const int nm = mirrors(i, m.H, m.reflect, mr);
    r.n = 1 + nm;
    r.base[0] = img + (int64_t(i + m.pt) * m.P + m.pl) * m.C;
    for (int k = 0; k < nm; ++k) r.base[1 + k] = img + (int64_t(mr[k] + m.pt) * m.P + m.pl) * m.C;
  } else {
    r.n = 1;
    const int64_t row = int64_t((i >> 1) + m.pt) * m.P + m.pl;
    r.base[0] = img + (int64_t((i & 1) * 2) * m.plane_pix + row) * m.C;
    r.base[1] = img + (int64_t((i & 1) * 2 + 1) * m.plane_pix + row) * m.C;
  }
}
__device__ __forceinline__ void dst_store8(const DstRow& r, const FrameMap& m, int j, const uint4& w) {
  if (m.kind == 0) {
    for (int k = 0; k < r.n; ++k) *reinterpret_cast<uint4*>(r.base[k] + int64_t(j) * m.C) = w;
    if (m.reflect > 0 && (j <= m.reflect || j >= m.W - 1 - m.reflect)) {
      int mc[2];
      const int nc = mirrors(j, m.W, m.reflect, mc);
      for (int q = 0; q < nc; ++q)
        for (int k = 0; k < r.n; ++k) *reinterpret_cast<uint4*>(r.base[k] + int64_t(mc[q]) * m.C) = w;
    }
  } else {
    *reinterpret_cast<uint4*>(r.base[j & 1] + int64_t(j >> 1) * m.C) = w;
  }
}

// ---- source row of a gradient buffer (bf16; fp32 sources never reach these kernels) -------------------------
struct SrcRow {
  const sg_bf16* base[3];  // row i (+ folded border rows), at logical column 0
  int n;
};
__device__ __forceinline__ void src_row_init(SrcRow& r, const GradSrc& g, int b, int i, int H, int C, int c0) {
  if (g.ptr == nullptr) {
    r.n = 0;
    return;
  }
  const sg_bf16* img = reinterpret_cast<const sg_bf16*>(g.ptr) + int64_t(b) * g.Hs * g.Ws * C + c0;
  int mr[2];
  const int nm = mirrors(i, H, g.fold, mr);
  r.n = 1 + nm;
  r.base[0] = img + (int64_t(i + g.oy) * g.Ws + g.ox) * C;
  for (int k = 0; k < nm; ++k) r.base[1 + k] = img + (int64_t(mr[k] + g.oy) * g.Ws + g.ox) * C;
}
// everything except the primary (row 0, column j) read
__device__ __forceinline__ void src_extra8(const SrcRow& r, const GradSrc& g, int j, int W, int C, float* acc) {
  int mc[2];
  const int nc = (g.fold > 0 && (j <= g.fold || j >= W - 1 - g.fold)) ? mirrors(j, W, g.fold, mc) : 0;
  for (int k = 0; k < r.n; ++k)
    for (int q = (k == 0 ? 1 : 0); q <= nc; ++q) {
      const int col = q == 0 ? j : mc[q - 1];
      float t[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(r.base[k] + int64_t(col) * C)), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += t[e];
    }
}
__device__ __forceinline__ bool src_has_extra(const SrcRow& r, const GradSrc& g, int j, int W) {
  return r.n > 1 || (g.fold > 0 && r.n > 0 && (j <= g.fold || j >= W - 1 - g.fold));
}

static inline int rows_per_block(int H, int B) {
  // ~5 blocks per SM over the grid
  int rpb = (H * B + 148 * 5 - 1) / (148 * 5);
  return rpb < 1 ? 1 : rpb;
}

// =======================================================================================================
// forward: z = act((y - mean) * gamma * rstd + beta) (+ residual)  ->  next frame
__global__ void __launch_bounds__(kRowThreads) in_apply_rows_kernel(const InApplyParams p, const int rpb) {
  const int b = blockIdx.y;
  const int C8 = p.C >> 3;
  const int cg = threadIdx.x % C8, lj = threadIdx.x / C8, ppi = kRowThreads / C8;
  const int c0 = cg * 8;
  const float n = float(p.H * p.W);
  float mean[8], scale[8], beta[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    float mu = 0.f, rs = 1.f;
    if (p.stats != nullptr) {
      const float2 s = reinterpret_cast<const float2*>(p.stats)[int64_t(b) * p.C + c];
      mu = s.x / n;
      rs = rsqrtf(fmaxf(s.y / n - mu * mu, 0.f) + p.eps);
    }
    mean[e] = mu;
    scale[e] = (p.gamma ? p.gamma[c] : 1.f) * rs;
    beta[e] = p.beta ? p.beta[c] : 0.f;  // z = (y - mean)*scale + beta: exactly beta when H*W == 1
  }
  const int r0 = blockIdx.x * rpb, r1 = min(p.H, r0 + rpb);
  constexpr int U = 4;
  for (int i = r0; i < r1; ++i) {
    const sg_bf16* yrow = p.Y + (int64_t(b) * p.H + i) * p.W * p.C + c0;
    const sg_bf16* rrow =
        p.res ? p.res + (int64_t(b) * p.rmap.frame_pix + int64_t(i + p.rmap.pt) * p.rmap.P + p.rmap.pl) * p.rmap.C + c0
              : nullptr;  // residual frames are single-plane (generator blocks)
    DstRow dr;
    dst_row_init(dr, p.dst, p.dmap, b, i, c0);
    for (int j0 = lj; j0 < p.W; j0 += U * ppi) {
      uint4 raw[U], rr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * ppi;
        if (j < p.W) {
          raw[u] = __ldg(reinterpret_cast<const uint4*>(yrow + int64_t(j) * p.C));
          if (rrow) rr[u] = __ldg(reinterpret_cast<const uint4*>(rrow + int64_t(j) * p.rmap.C));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * ppi;
        if (j >= p.W) break;
        float y[8];
        unpack8(raw[u], y);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = rw_act_fwd(fmaf(y[e] - mean[e], scale[e], beta[e]), p.act, p.act_alpha);
        if (rrow) {
          float r[8];
          unpack8(rr[u], r);
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] += r[e];
        }
        dst_store8(dr, p.dmap, j, pack8(y));
      }
    }
  }
}
void launch_in_apply(const InApplyParams& p, cudaStream_t st) {
  const int rpb = rows_per_block(p.H, p.B);
  dim3 grid((p.H + rpb - 1) / rpb, p.B);
  in_apply_rows_kernel<<<grid, kRowThreads, 0, st>>>(p, rpb);
}

// =======================================================================================================
// backward.  reduce: sums[b][c] += (sum dzh, sum dzh * xhat);  apply: dy = scale*((dzh - m1) - xhat*m2) -> dY frame
template <bool kApply>
__global__ void __launch_bounds__(kRowThreads, 2) in_bwd_rows_kernel(const InBwdParams p, const int rpb) {
  extern __shared__ float sred[];
  const int b = blockIdx.y;
  const int ba = b < p.nb_act ? b : b - p.act_wrap;
  const int C8 = p.C >> 3;
  const int cg = threadIdx.x % C8, lj = threadIdx.x / C8, ppi = kRowThreads / C8;
  const int c0 = cg * 8;
  const float n = float(p.H * p.W);
  // zpre = (y - mean)*scale + beta;  xhat = (y - mean)*rstd.  The apply form keeps a single-pixel norm
  // (xhat == 0, dzh == m1) back-propagating exactly zero, as the reference does at 128x128 (Appendix B).
  float mean[8], rstd[8], scale[8], beta[8], a1[8], a2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    const float2 s = reinterpret_cast<const float2*>(p.stats)[int64_t(ba) * p.C + c];
    const float mu = s.x / n;
    mean[e] = mu;
    rstd[e] = rsqrtf(fmaxf(s.y / n - mu * mu, 0.f) + p.eps);
    scale[e] = p.gamma[c] * rstd[e];
    beta[e] = p.beta[c];
    if (kApply) {
      const float2 q = reinterpret_cast<const float2*>(p.sums)[int64_t(b) * p.C + c];
      a1[e] = q.x / n;
      a2[e] = q.y / n;
    } else {
      a1[e] = a2[e] = 0.f;
    }
  }
  if (!kApply) {
    for (int t = threadIdx.x; t < 2 * p.C; t += kRowThreads) sred[t] = 0.f;
    __syncthreads();
  }
  const int r0 = blockIdx.x * rpb, r1 = min(p.H, r0 + rpb);
  constexpr int U = 2;
  for (int i = r0; i < r1; ++i) {
    const sg_bf16* yrow = p.Y + (int64_t(ba) * p.H + i) * p.W * p.C + c0;
    SrcRow s1, s2;
    src_row_init(s1, p.g1, b, i, p.H, p.C, c0);
    src_row_init(s2, p.g2, b, i, p.H, p.C, c0);
    DstRow dr;
    if (kApply) dst_row_init(dr, p.dst, p.dmap, b, i, c0);
    for (int j0 = lj; j0 < p.W; j0 += U * ppi) {
      uint4 raw[U], g1[U], g2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * ppi;
        g1[u] = make_uint4(0, 0, 0, 0);
        g2[u] = make_uint4(0, 0, 0, 0);
        if (j < p.W) {
          raw[u] = __ldg(reinterpret_cast<const uint4*>(yrow + int64_t(j) * p.C));
          if (s1.n) g1[u] = __ldg(reinterpret_cast<const uint4*>(s1.base[0] + int64_t(j) * p.C));
          if (s2.n) g2[u] = __ldg(reinterpret_cast<const uint4*>(s2.base[0] + int64_t(j) * p.C));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * ppi;
        if (j >= p.W) break;
        float y[8], d[8], t[8];
        unpack8(raw[u], y);
        unpack8(g1[u], d);
        unpack8(g2[u], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
        if (src_has_extra(s1, p.g1, j, p.W)) src_extra8(s1, p.g1, j, p.W, p.C, d);
        if (src_has_extra(s2, p.g2, j, p.W)) src_extra8(s2, p.g2, j, p.W, p.C, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float yc = y[e] - mean[e];
          const float dz = d[e] * rw_act_grad(fmaf(yc, scale[e], beta[e]), p.act, p.act_alpha);
          const float xh = yc * rstd[e];
          if (kApply) {
            d[e] = scale[e] * ((dz - a1[e]) - xh * a2[e]);
          } else {
            a1[e] += dz;
            a2[e] += dz * xh;
          }
        }
        if (kApply) dst_store8(dr, p.dmap, j, pack8(d));
      }
    }
  }
  if (!kApply) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&sred[(c0 + e) * 2], a1[e]);
      atomicAdd(&sred[(c0 + e) * 2 + 1], a2[e]);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * p.C; t += kRowThreads) atomicAdd(p.sums + int64_t(b) * p.C * 2 + t, sred[t]);
  }
}
void launch_in_bwd_reduce(const InBwdParams& p, cudaStream_t st) {
  const int rpb = rows_per_block(p.H, p.B);
  dim3 grid((p.H + rpb - 1) / rpb, p.B);
  in_bwd_rows_kernel<false><<<grid, kRowThreads, 2 * p.C * sizeof(float), st>>>(p, rpb);
}
void launch_in_bwd_apply(const InBwdParams& p, cudaStream_t st) {
  const int rpb = rows_per_block(p.H, p.B);
  dim3 grid((p.H + rpb - 1) / rpb, p.B);
  in_bwd_rows_kernel<true><<<grid, kRowThreads, 0, st>>>(p, rpb);
}

// =======================================================================================================
// out[b,i,j,:] = g1 + g2   (plain bf16 [B][H][W][C]); either source may fold a reflected border back
__global__ void __launch_bounds__(kRowThreads) grad_gather_rows_kernel(const GradSrc g1, const GradSrc g2, int H, int W,
                                                                      int C, sg_bf16* out, int rpb) {
  const int b = blockIdx.y;
  const int C8 = C >> 3;
  const int cg = threadIdx.x % C8, lj = threadIdx.x / C8, ppi = kRowThreads / C8;
  const int c0 = cg * 8;
  const int r0 = blockIdx.x * rpb, r1 = min(H, r0 + rpb);
  constexpr int U = 4;
  for (int i = r0; i < r1; ++i) {
    SrcRow s1, s2;
    src_row_init(s1, g1, b, i, H, C, c0);
    src_row_init(s2, g2, b, i, H, C, c0);
    sg_bf16* orow = out + (int64_t(b) * H + i) * W * C + c0;
    for (int j0 = lj; j0 < W; j0 += U * ppi) {
      uint4 a[U], c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * ppi;
        a[u] = make_uint4(0, 0, 0, 0);
        c[u] = make_uint4(0, 0, 0, 0);
        if (j < W) {
          if (s1.n) a[u] = __ldg(reinterpret_cast<const uint4*>(s1.base[0] + int64_t(j) * C));
          if (s2.n) c[u] = __ldg(reinterpret_cast<const uint4*>(s2.base[0] + int64_t(j) * C));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * ppi;
        if (j >= W) break;
        float d[8], t[8];
        unpack8(a[u], d);
        unpack8(c[u], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
        if (src_has_extra(s1, g1, j, W)) src_extra8(s1, g1, j, W, C, d);
        if (src_has_extra(s2, g2, j, W)) src_extra8(s2, g2, j, W, C, d);
        *reinterpret_cast<uint4*>(orow + int64_t(j) * C) = pack8(d);
      }
    }
  }
}
void launch_grad_gather(const GradSrc& g1, const GradSrc& g2, int B, int H, int W, int C, sg_bf16* out,
                        cudaStream_t st) {
  const int rpb = rows_per_block(H, B);
  dim3 grid((H + rpb - 1) / rpb, B);
  grad_gather_rows_kernel<<<grid, kRowThreads, 0, st>>>(g1, g2, H, W, C, out, rpb);
}

}  // namespace sggan
