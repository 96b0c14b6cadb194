// glue_rows.cu -- the per-layer elementwise passes of the step as TMA-staged row streams:
//
//   in_apply      instance norm + activation (+ residual) of a raw conv output -> next frame
//   in_bwd        its backward (reduce pass and apply pass)
//   grad_gather   residual-stream gradient accumulation
//
// These passes move 70-110 MB each and were latency-bound (1.2-2.2 TB/s) as plain load/compute/store loops:
// with ~100 live registers per thread only 16 warps fit an SM and every warp idles for the HBM round trip
// between its batches.  Here one persistent block per SM streams its share of an image as 16 KB row chunks:
// a producer warp runs four chunks ahead with 1-D bulk copies (cp.async.bulk -> shared, mbarrier
// complete_tx), so ~190 KB per SM are in flight independent of register pressure, and 16 consumer warps
// compute out of shared memory (conflict-free 128-bit accesses, 8 channels per thread) and store 128-bit
// results straight into the destination frame.  Row bookkeeping (reflected border rows/columns, 2x2 phase
// planes, folded gradient borders) is resolved once per chunk into base pointers.
#include <cuda_bf16.h>

#include "glue.h"
#include "tc_common.cuh"

namespace sggan {

#ifndef SG_CONSUMERS
#define SG_CONSUMERS 512
#endif
#ifndef SG_STAGES
#define SG_STAGES 4
#endif
constexpr int kConsumers = SG_CONSUMERS;         // consumer threads (16 warps)
constexpr int kStreamThreads = kConsumers + 32;  // + 1 producer warp
constexpr int kStages = SG_STAGES;
constexpr int kPipeBytes = 196608;  // shared memory of the whole pipeline: kStages x streams x chunk
constexpr int kMaxStreams = 3;

enum { RS_APPLY = 0, RS_BWD_REDUCE = 1, RS_BWD_APPLY = 2, RS_GATHER = 3 };

struct StreamDesc {
  const sg_bf16* base;  // null = absent (reads as zeros)
  int64_t img_stride;   // elements between images
  int64_t row_stride;   // elements between rows
  int oy, ox;           // logical (0,0) sits at row oy, column ox
  int act_index;        // 1: indexed by the activation image (virtual-batch wrap), 0: by the gradient image
};

struct RowStreamParams {
  int B, H, W, C, CW;  // CW = pixels per chunk
  int ns;              // active streams (a prefix of s[])
  int chunk_bytes;     // bytes per stream and stage: the pipeline memory is split over kStages x ns chunks
  int nb_act, act_wrap;
  StreamDesc s[kMaxStreams];
  // statistics / affine
  const float* stats;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  float alpha;
  float* sums;   // in_bwd: [B][C][2]
  GradSrc g[2];  // fold information of the gradient streams (extras are read directly from global)
  sg_bf16* dst;  // frame (apply modes) or plain [B][H][W][C] (gather)
  FrameMap dmap;
};

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
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

// ---- folded-border extras of a gradient source (everything except the primary read) ---------------------------
struct SrcRow {
  const sg_bf16* base[3];
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
__device__ __forceinline__ bool src_has_extra(const SrcRow& r, const GradSrc& g, int j, int W) {
  return r.n > 1 || (g.fold > 0 && r.n > 0 && (j <= g.fold || j >= W - 1 - g.fold));
}
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

// =======================================================================================================
template <int MODE, int NS>
__global__ void __launch_bounds__(kStreamThreads, 1) row_stream_kernel(const RowStreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // pointer arithmetic keeps the shared address space
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ba = b < p.nb_act ? b : b - p.act_wrap;
  const int cpr = (p.W + p.CW - 1) / p.CW;  // chunks per row
  const int nchunks = p.H * cpr;
  // each block owns a contiguous range of chunks, so consecutive chunks mostly share their image row and the
  // per-row bookkeeping (border rows, base pointers) is amortised
  const int cper = (nchunks + gridDim.x - 1) / gridDim.x;
  const int cbeg = blockIdx.x * cper, cend = min(nchunks, cbeg + cper);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == kConsumers / 32) {
    // ------------------------------------------------------------ producer: bulk copies, kStages chunks ahead
    if (lane == 0) {
      int k = 0;
      int i = cbeg / cpr, jc = cbeg - i * cpr;
      for (int c = cbeg; c < cend; ++c, ++k) {
        const int s = k % kStages;
        mbar_wait(&empty_bar[s], ((k / kStages) & 1) ^ 1, 41);
        const int j0 = jc * p.CW;
        const int cw = min(p.CW, p.W - j0);
        const uint32_t bytes = uint32_t(cw) * p.C * 2;
        mbar_arrive_expect_tx(&full_bar[s], bytes * NS);
        for (int q = 0; q < NS; ++q) {
          const StreamDesc& d = p.s[q];
          const sg_bf16* src = d.base + int64_t(d.act_index ? ba : b) * d.img_stride + int64_t(i + d.oy) * d.row_stride +
                               int64_t(j0 + d.ox) * p.C;
          bulk_load_1d(smem + (s * NS + q) * p.chunk_bytes, src, bytes, &full_bar[s]);
        }
        if (++jc == cpr) { jc = 0; ++i; }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  // Thread t owns vector t (16 B = 8 channels) of every 512-vector slab of a chunk: its channel group is
  // fixed, its pixel advances by 512/C8 per slab, so all addressing is incremental.
  const int C8 = p.C >> 3;
  const int cg = threadIdx.x % C8;
  const int c0 = cg * 8;
  const int px0 = threadIdx.x / C8, pstep = kConsumers / C8;  // pstep is even (C8 <= 64)
  const float n = float(p.H * p.W);
  // activation as a slope for the non-positive side: relu 0, leaky alpha, identity 1 (tanh never reaches the glue)
  const float gneg = p.act == SG_ACT_RELU ? 0.f : (p.act == SG_ACT_LRELU ? p.alpha : 1.f);
  float mean[8], rstd[8], scale[8], beta[8], a1[8], a2[8];
  if (MODE != RS_GATHER) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      float mu = 0.f, rs = 1.f;
      if (p.stats != nullptr) {
        const float2 st = reinterpret_cast<const float2*>(p.stats)[int64_t(ba) * p.C + c];
        mu = st.x / n;
        rs = rsqrtf(fmaxf(st.y / n - mu * mu, 0.f) + p.eps);
      }
      mean[e] = mu;
      rstd[e] = rs;
      scale[e] = (p.gamma ? p.gamma[c] : 1.f) * rs;
      beta[e] = p.beta ? p.beta[c] : 0.f;  // z = (y - mean)*scale + beta: exactly beta when H*W == 1
      a1[e] = a2[e] = 0.f;
      if (MODE == RS_BWD_APPLY) {
        const float2 q = reinterpret_cast<const float2*>(p.sums)[int64_t(b) * p.C + c];
        a1[e] = q.x / n;  // mean of dzh
        a2[e] = q.y / n;  // mean of dzh * xhat
      }
    }
  }
  const uint32_t chunk_bytes = p.chunk_bytes, stage_bytes = NS * p.chunk_bytes;
  const int W = p.W, C = p.C, CW = p.CW;
  const int src_band = (MODE == RS_APPLY) ? 0 : max(p.g[0].ptr ? p.g[0].fold : 0, p.g[1].ptr ? p.g[1].fold : 0);
  const int dst_band = ((MODE == RS_APPLY || MODE == RS_BWD_APPLY) && p.dmap.kind == 0) ? p.dmap.reflect : 0;
  const int band = max(src_band, dst_band);
  const int dC = (MODE == RS_GATHER) ? C : p.dmap.C;
  const int dkind = (MODE == RS_GATHER) ? 0 : p.dmap.kind;
  const int pshift = 31 - __clz(pstep);  // pstep = 512 / C8 is a power of two
  const uint32_t sbase = smem_u32(smem) + threadIdx.x * 16;
  DstRow dr;
  SrcRow e1, e2;
  e1.n = e2.n = 0;
  dr.n = 1;

  // one pixel (8 channels) of this thread: `sa` = its vector in stream 0 of the stage, `dptr` = where the lean
  // (no border bookkeeping) result goes, `j` = image column
  auto pixel = [&](const bool lean, const uint32_t sa, sg_bf16* dptr, const int j) {
    uint4 r0 = lds128(sa), r1 = make_uint4(0, 0, 0, 0), r2 = r1;
    if (NS > 1) r1 = lds128(sa + chunk_bytes);
    if (NS > 2) r2 = lds128(sa + 2 * chunk_bytes);
    float y[8], d[8];
    if (MODE == RS_APPLY) {
      unpack8(r0, y);
      unpack8(r1, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = fmaf(y[e] - mean[e], scale[e], beta[e]);
        y[e] = (z > 0.f ? z : z * gneg) + d[e];
      }
      if (lean) *reinterpret_cast<uint4*>(dptr) = pack8(y);
      else dst_store8(dr, p.dmap, j, pack8(y));
    } else if (MODE == RS_GATHER) {
      unpack8(r0, d);
      unpack8(r1, y);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] += y[e];
      if (!lean) {
        if (src_has_extra(e1, p.g[0], j, W)) src_extra8(e1, p.g[0], j, W, C, d);
        if (src_has_extra(e2, p.g[1], j, W)) src_extra8(e2, p.g[1], j, W, C, d);
      }
      *reinterpret_cast<uint4*>(dptr) = pack8(d);
    } else {
      unpack8(r0, y);
      unpack8(r1, d);
      if (NS > 2) {
        float t[8];
        unpack8(r2, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
      }
      if (!lean) {
        if (src_has_extra(e1, p.g[0], j, W)) src_extra8(e1, p.g[0], j, W, C, d);
        if (src_has_extra(e2, p.g[1], j, W)) src_extra8(e2, p.g[1], j, W, C, d);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        // zpre = (y - mean)*scale + beta;  xhat = (y - mean)*rstd.  The apply form keeps a single-pixel norm
        // (xhat == 0, dzh == m1) back-propagating exactly zero, as the reference does at 128x128 (Appendix B).
        const float yc = y[e] - mean[e];
        const float dz = fmaf(yc, scale[e], beta[e]) > 0.f ? d[e] : d[e] * gneg;
        const float xh = yc * rstd[e];
        if (MODE == RS_BWD_APPLY) {
          d[e] = scale[e] * ((dz - a1[e]) - xh * a2[e]);
        } else {
          a1[e] += dz;
          a2[e] = fmaf(dz, xh, a2[e]);
        }
      }
      if (MODE == RS_BWD_APPLY) {
        if (lean) *reinterpret_cast<uint4*>(dptr) = pack8(d);
        else dst_store8(dr, p.dmap, j, pack8(d));
      }
    }
  };

  int k = 0;
  int cur_i = -1;
  int i = cbeg / cpr, jc = cbeg - i * cpr;
  for (int c = cbeg; c < cend; ++c, ++k, jc = (jc + 1 == cpr ? 0 : jc + 1), i += (jc == 0)) {
    const int s = k % kStages;
    const int j0 = jc * CW;
    const int cw = min(CW, W - j0);
    if (i != cur_i) {  // new image row: resolve border rows and base pointers once
      cur_i = i;
      if (MODE == RS_APPLY || MODE == RS_BWD_APPLY) dst_row_init(dr, p.dst, p.dmap, b, i, c0);
      if (MODE != RS_APPLY) {
        src_row_init(e1, p.g[0], b, i, p.H, C, c0);
        src_row_init(e2, p.g[1], b, i, p.H, C, c0);
      }
    }
    // pixels [lo, hi) of this chunk need no border bookkeeping
    int lo = 0, hi = cw;
    if (dr.n > 1 || e1.n > 1 || e2.n > 1) {
      hi = 0;
    } else if (band > 0) {
      lo = min(cw, max(0, band + 1 - j0));
      hi = max(lo, min(cw, W - 1 - band - j0));
    }
    // this thread visits pixels px0 + t * pstep, t in [0, nt); t in [tlo, thi) are lean
    const int nt = cw > px0 ? ((cw - px0 + pstep - 1) >> pshift) : 0;
    const int tlo = min(nt, lo > px0 ? ((lo - px0 + pstep - 1) >> pshift) : 0);
    const int thi = max(tlo, min(nt, hi > px0 ? ((hi - px0 + pstep - 1) >> pshift) : 0));
    // incremental destination pointer of this thread (element units)
    sg_bf16* dptr;
    int dstep;
    if (MODE == RS_GATHER) {
      dptr = p.dst + ((int64_t(b) * p.H + i) * W + j0 + px0) * C + c0;
      dstep = pstep * C;
    } else if (dkind == 0) {
      dptr = dr.base[0] + (j0 + px0) * dC;
      dstep = pstep * dC;
    } else {  // phase planes: the column parity of this thread is fixed because pstep is even
      dptr = dr.base[(j0 + px0) & 1] + ((j0 + px0) >> 1) * dC;
      dstep = (pstep >> 1) * dC;
    }
    mbar_wait(&full_bar[s], (k / kStages) & 1, 42);
    uint32_t sa = sbase + s * stage_bytes;
    int j = j0 + px0;
    int t = 0;
    for (; t < tlo; ++t, sa += kConsumers * 16, dptr += dstep, j += pstep) pixel(false, sa, dptr, j);
#pragma unroll 2
    for (; t < thi; ++t, sa += kConsumers * 16, dptr += dstep, j += pstep) pixel(true, sa, dptr, j);
    for (; t < nt; ++t, sa += kConsumers * 16, dptr += dstep, j += pstep) pixel(false, sa, dptr, j);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  if (MODE == RS_BWD_REDUCE) {
    // threads that share a channel group differ in px0 (kConsumers / C8 of them): stage their partials in the
    // (now idle) pipeline buffers as [px0][e][k][cg] (consecutive lanes -> consecutive words) and add them
    // in a fixed order
    float* red = reinterpret_cast<float*>(smem);
    named_bar_sync(2, kConsumers);  // every consumer is past its last chunk: the stage buffers are free
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[((px0 * 8 + e) * 2 + 0) * C8 + cg] = a1[e];
      red[((px0 * 8 + e) * 2 + 1) * C8 + cg] = a2[e];
    }
    named_bar_sync(2, kConsumers);
    for (int t = threadIdx.x; t < 2 * C; t += kConsumers) {
      const int ch = t >> 1, kk = t & 1;
      const int idx = (((ch & 7) * 2) + kk) * C8 + (ch >> 3);
      float acc = 0.f;
      for (int q = 0; q < pstep; ++q) acc += red[q * 16 * C8 + idx];
      atomicAdd(p.sums + int64_t(b) * C * 2 + t, acc);
    }
  }
}

template <int MODE, int NS>
static void launch_row_stream_ns(const RowStreamParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(row_stream_kernel<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    attr_set = true;
  }
  row_stream_kernel<MODE, NS><<<grid, kStreamThreads, smem, st>>>(p);
}

template <int MODE>
static void launch_row_stream(RowStreamParams& p, cudaStream_t st) {
  p.ns = 0;
  while (p.ns < kMaxStreams && p.s[p.ns].base != nullptr) ++p.ns;  // active streams are a prefix
  if (p.ns == 0) return;
  p.CW = kPipeBytes / (kStages * p.ns) / (p.C * 2);
  if (p.CW > p.W) p.CW = p.W;
  p.chunk_bytes = p.CW * p.C * 2;
  const size_t smem = size_t(kPipeBytes) + 128;
  // one persistent block per SM; blocks never span images (per-image statistics)
  int gx = 148 / p.B;
  if (gx < 1) gx = 1;
  const int nchunks = p.H * ((p.W + p.CW - 1) / p.CW);
  if (gx > nchunks) gx = nchunks;
  dim3 grid(gx, p.B);
  if (p.ns == 1) launch_row_stream_ns<MODE, 1>(p, grid, smem, st);
  else if (p.ns == 2) launch_row_stream_ns<MODE, 2>(p, grid, smem, st);
  else launch_row_stream_ns<MODE, 3>(p, grid, smem, st);
}

static StreamDesc plain_stream(const sg_bf16* base, int H, int W, int C, int act_index) {
  StreamDesc d;
  d.base = base; d.img_stride = int64_t(H) * W * C; d.row_stride = int64_t(W) * C; d.oy = 0; d.ox = 0; d.act_index = act_index;
  return d;
}
static StreamDesc grad_stream(const GradSrc& g, int C) {
  StreamDesc d;
  d.base = reinterpret_cast<const sg_bf16*>(g.ptr);
  d.img_stride = int64_t(g.Hs) * g.Ws * C; d.row_stride = int64_t(g.Ws) * C; d.oy = g.oy; d.ox = g.ox; d.act_index = 0;
  return d;
}
static StreamDesc null_stream() {
  StreamDesc d;
  d.base = nullptr; d.img_stride = d.row_stride = 0; d.oy = d.ox = d.act_index = 0;
  return d;
}

void launch_in_apply(const InApplyParams& a, cudaStream_t st) {
  RowStreamParams p = {};
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.nb_act = a.B; p.act_wrap = 0;
  p.s[0] = plain_stream(a.Y, a.H, a.W, a.C, 0);
  p.s[1] = null_stream();
  if (a.res != nullptr) {  // residual frames are single-plane (generator blocks)
    p.s[1].base = a.res; p.s[1].img_stride = a.rmap.frame_pix * a.rmap.C; p.s[1].row_stride = int64_t(a.rmap.P) * a.rmap.C;
    p.s[1].oy = a.rmap.pt; p.s[1].ox = a.rmap.pl;
  }
  p.s[2] = null_stream();
  p.stats = a.stats; p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.act = a.act; p.alpha = a.act_alpha;
  p.dst = a.dst; p.dmap = a.dmap;
  launch_row_stream<RS_APPLY>(p, st);
}

static void bwd_params(const InBwdParams& a, RowStreamParams& p) {
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.nb_act = a.nb_act; p.act_wrap = a.act_wrap;
  p.s[0] = plain_stream(a.Y, a.H, a.W, a.C, 1);
  p.s[1] = a.g1.ptr ? grad_stream(a.g1, a.C) : null_stream();
  p.s[2] = a.g2.ptr ? grad_stream(a.g2, a.C) : null_stream();
  p.g[0] = a.g1; p.g[1] = a.g2;
  p.stats = a.stats; p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.act = a.act; p.alpha = a.act_alpha;
  p.sums = a.sums; p.dst = a.dst; p.dmap = a.dmap;
}
void launch_in_bwd_reduce(const InBwdParams& a, cudaStream_t st) {
  RowStreamParams p = {};
  bwd_params(a, p);
  launch_row_stream<RS_BWD_REDUCE>(p, st);
}
void launch_in_bwd_apply(const InBwdParams& a, cudaStream_t st) {
  RowStreamParams p = {};
  bwd_params(a, p);
  launch_row_stream<RS_BWD_APPLY>(p, st);
}

void launch_grad_gather(const GradSrc& g1, const GradSrc& g2, int B, int H, int W, int C, sg_bf16* out, cudaStream_t st) {
  RowStreamParams p = {};
  p.B = B; p.H = H; p.W = W; p.C = C; p.nb_act = B; p.act_wrap = 0;
  p.s[0] = g1.ptr ? grad_stream(g1, C) : null_stream();
  p.s[1] = g2.ptr ? grad_stream(g2, C) : null_stream();
  p.s[2] = null_stream();
  p.g[0] = g1; p.g[1] = g2;
  p.dst = out;
  launch_row_stream<RS_GATHER>(p, st);
}

}  // namespace sggan
