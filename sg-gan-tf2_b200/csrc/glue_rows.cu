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
// planes, folded gradient borders) is resolved once per chunk by the producer thread and handed over as a
// descriptor in shared memory.  Fused into these passes: the fixed-order finalize of the convolution's per-tile
// statistics (apply prologue), dgamma / dbeta (backward apply), and the residual-stream gradient gather
// (backward reduce of the layer that reads it).
#include <cuda_bf16.h>

#include <algorithm>
#include <type_traits>

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
constexpr int kMaxC = 512;                                    // channels per pixel the passes support
constexpr int kCoefBytes = 6 * kMaxC * 4;                     // per-channel coefficients staged in shared memory
constexpr int kSumBytes = (kConsumers + kMaxC) * 8;           // fixed-order sums: [slices][C] float2 + [C] float2

enum { RS_APPLY = 0, RS_BWD_REDUCE = 1, RS_BWD_APPLY = 2, RS_GATHER = 3 };

#ifdef SG_ROWS_DEBUG
// tests/gpu/rows_probe.cu: per-block clock64 stamps [block][8]: 0 start, 1 after the dependency wait, 2 consumer
// coefficients ready, 3 first chunk arrived, 4 last chunk consumed, 5 end, 6 last load issued (producer)
// slots 8 / 9: %globaltimer (ns, comparable across SMs and launches) at block start / end; launch l of a probe
// sequence writes to region (l % 8) (the host advances g_rows_dbg_launch)
__device__ long long* g_rows_dbg = nullptr;
static int g_rows_dbg_launch = 0;
__device__ __forceinline__ long long rs_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RS_STAMP(slot, cond)                                                                          \
  do {                                                                                                \
    if (g_rows_dbg != nullptr && (cond)) {                                                            \
      long long* d_ = g_rows_dbg + ((p.dbg_launch & 7) * 256 + blockIdx.y * gridDim.x + blockIdx.x) * 16; \
      d_[slot] = clock64();                                                                           \
      if ((slot) == 0) d_[8] = rs_globaltimer();                                                      \
      if ((slot) == 5) d_[9] = rs_globaltimer();                                                      \
    }                                                                                                 \
  } while (0)
#else
#define RS_STAMP(slot, cond) do { } while (0)
#endif

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
  int evict_first;     // load the streams with an L2 evict-first hint
  int chunk_bytes;     // bytes per slot: the pipeline memory is split over stages x (ns + mslot) slots
  int stages;          // pipeline stages in use (<= kStages)
  int mslot;           // 1: every stage has one more slot, for the mirrored source row of a folded-border row
  int nb_act, act_wrap;
  StreamDesc s[kMaxStreams];
  // statistics / affine
  const float* stats;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  float alpha;
  float* sums;   // in_bwd apply: the reduced (sum dzh, sum dzh * xhat) per (image, channel) are published here, [B][C][2]
  float* sums_part;  // in_bwd: per-block partial sums [B][gridDim.x][C][2]: written by the reduce pass with plain
                     // stores, added in a fixed order by the apply pass (no atomics: the backward is bit-reproducible)
  const float* stats_part;  // in_apply: per-tile partials to finalize in the kernel (or null), see InApplyParams
  int stats_T;
  float* stats_out;
  sg_bf16* gather_dst;  // in_bwd reduce: also store g1 + g2 (folded) as plain [B][H][W][C] (or null)
  int sums_nblk;  // in_bwd apply: partials per image (= blocks per image of the reduce launch)
  int* sync_ctr;  // fused in_bwd: one arrival counter per image, zero before the launch
  int mrows, wm;  // rows 1..mrows and H-1-mrows..H-2 fold a mirrored row in (source) or write one (destination): their
                  // chunks count `wm` times in the split of an image's chunks over its blocks
#ifdef SG_ROWS_DEBUG
  int dbg_launch;
#endif
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
// mirror images of index i inside a reflect border of width p (excluding i itself): returns how many (0..2)
__device__ __forceinline__ int mirrors(int i, int n, int p, int& m0, int& m1) {
  int k = 0;
  m0 = m1 = 0;
  if (p > 0) {
    if (i >= 1 && i <= p) { m0 = -i; k = 1; }
    if (i >= n - 1 - p && i <= n - 2) {
      if (k == 0) m0 = 2 * (n - 1) - i; else m1 = 2 * (n - 1) - i;
      ++k;
    }
  }
  return k;
}

// ---- per-chunk bookkeeping, resolved ONCE by the producer thread and handed to the consumers through shared
// memory next to the data (the consumers used to redo it per thread: reflected rows, plane bases, 64-bit
// multiplies, local-memory arrays -- about as many instructions as the arithmetic itself) -----------------------
struct __align__(16) ChunkDesc {
  int i, j0, cw, lo;   // image row, first column, width; pixels [lo, hi) need no border bookkeeping
  int hi, dn, n1, n2;  // rows the result goes to (1 + reflected copies); rows of the two folded gradient sources
  long long dbase[3];  // element offsets (without the channel) of the destination rows at logical column 0;
                       // phase-split frames: [0] even-column plane, [1] odd-column plane
  long long s1[3];     // element offsets of the gradient-source rows (primary first, then its mirrored rows)
  long long s2[3];
  int lead[2];         // border pixels loaded in front of the chunk for gradient source 0 / 1 (first chunk of a row)
};
static_assert(sizeof(ChunkDesc) == 112 || sizeof(ChunkDesc) == 128, "ChunkDesc layout");

__device__ __forceinline__ void dst_store8(const ChunkDesc* cd, const FrameMap& m, sg_bf16* dst_c0, int j, const uint4& w) {
  if (m.kind == 0) {
    const int dn = cd->dn;
    for (int k = 0; k < dn; ++k) *reinterpret_cast<uint4*>(dst_c0 + cd->dbase[k] + int64_t(j) * m.C) = w;
    if (m.reflect > 0 && (j <= m.reflect || j >= m.W - 1 - m.reflect)) {
      int c0, c1;
      const int nc = mirrors(j, m.W, m.reflect, c0, c1);
      for (int q = 0; q < nc; ++q)
        for (int k = 0; k < dn; ++k) *reinterpret_cast<uint4*>(dst_c0 + cd->dbase[k] + int64_t(q ? c1 : c0) * m.C) = w;
    }
  } else {
    *reinterpret_cast<uint4*>(dst_c0 + cd->dbase[j & 1] + int64_t(j >> 1) * m.C) = w;
  }
}
// folded-border extras of a gradient source: everything except the primary read (row 0, column j)
__device__ __forceinline__ bool src_has_extra(int n, const GradSrc& g, int j, int W) {
  return n > 1 || (g.fold > 0 && n > 0 && (j <= g.fold || j >= W - 1 - g.fold));
}
__device__ __forceinline__ void src_extra8(const long long* rows, int n, const GradSrc& g, int j, int W, int C, int c0, float* acc,
                                           uint32_t sa_q, const bool has_m = false, const uint32_t sm_q = 0) {
  // sm_q (if has_m): shared-memory address of this pixel's vector in the staged FIRST MIRRORED ROW of the source
  // sa_q: shared-memory address of THIS pixel's vector in the staged row of the source.  The producer also stages the
  // reflected border columns of the row (ChunkDesc::lead) and the first mirrored row (the stage's extra slot), so
  // everything a border of width 1 folds in comes out of shared memory; deeper borders (the 7x7 output convolution)
  // fetch their second and third mirrored row from global memory here.
  int m0 = 0, m1 = 0;
  const int nc = (g.fold > 0 && (j <= g.fold || j >= W - 1 - g.fold)) ? mirrors(j, W, g.fold, m0, m1) : 0;
  const sg_bf16* src_c0 = reinterpret_cast<const sg_bf16*>(g.ptr) + c0;
  for (int k = 0; k < n; ++k)
    for (int q = (k == 0 ? 1 : 0); q <= nc; ++q) {
      const int col = q == 0 ? j : (q == 1 ? m0 : m1);
      float t[8];
      if (k == 0) unpack8(lds128(sa_q + uint32_t((col - j) * C * 2)), t);
      else if (k == 1 && has_m) unpack8(lds128(sm_q + uint32_t((col - j) * C * 2)), t);
      else unpack8(__ldg(reinterpret_cast<const uint4*>(src_c0 + rows[k] + int64_t(col) * C)), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += t[e];
    }
}

// Column sums of src[T][C] (float2 entries) in a fixed order by the kConsumers consumer threads: thread (slice, channel
// pair) adds rows slice, slice + slices, ... (128-bit loads), then the slices are added in order.  Split in two halves
// so that the first kPre rows per thread are ISSUED before the producer releases its bulk loads (see `go_bar`): a small
// load queued behind the ~28 MB the producers of all blocks request at kernel start waits 4-8 us for its turn.
constexpr int kPre = 12;
struct SumPre {
  float4 v[kPre];
};
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {  // volatile: stays in front of the go_bar arrive
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_nc_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_nc_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {  // L2-coherent: data written by other blocks of this launch
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
template <bool COHERENT = false>
__device__ __forceinline__ void fixed_order_sum_issue(const float2* __restrict__ src, int T, int C, int tid, SumPre& pre) {
  const int CP = C >> 1;
  const int slices = kConsumers / CP;
  const int sl = tid / CP, cp = tid - sl * CP;
  if (sl < slices) {
    const float4* s4 = reinterpret_cast<const float4*>(src) + cp;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int t = sl + u * slices;
      pre.v[u] = t < T ? (COHERENT ? ld_cg_f4(s4 + int64_t(t) * CP) : ld_nc_f4(s4 + int64_t(t) * CP))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
// Second half, step 1: add this thread's rows (the pre-loaded ones, then any further rows in batches of eight
// independent loads) and park the slice sum in shared memory.  Everything this thread loads has ARRIVED when it returns.
template <bool COHERENT = false>
__device__ __forceinline__ void fixed_order_sum_accumulate(const float2* __restrict__ src, int T, int C, float2* scratch,
                                                           int tid, const SumPre& pre) {
  const int CP = C >> 1;
  const int slices = kConsumers / CP;
  const int sl = tid / CP, cp = tid - sl * CP;
  float4* part4 = reinterpret_cast<float4*>(scratch);  // [slices][CP]
  if (sl < slices) {
    const float4* s4 = reinterpret_cast<const float4*>(src) + cp;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kPre; ++u) { acc.x += pre.v[u].x; acc.y += pre.v[u].y; acc.z += pre.v[u].z; acc.w += pre.v[u].w; }
    for (int t = sl + kPre * slices; t < T; t += 8 * slices) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (t + u * slices < T) ? (COHERENT ? ld_cg_f4(s4 + int64_t(t + u * slices) * CP)
                                                : ld_nc_f4(s4 + int64_t(t + u * slices) * CP))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    part4[sl * CP + cp] = acc;
  }
}
// Step 2: add the slices in order.  Result: fin[C] in shared memory (valid after the call for all consumers).
__device__ __forceinline__ const float2* fixed_order_sum_combine(int C, float2* scratch, int tid) {
  const int CP = C >> 1;
  const int slices = kConsumers / CP;
  float4* part4 = reinterpret_cast<float4*>(scratch);            // [slices][CP]
  float4* fin4 = reinterpret_cast<float4*>(scratch + kConsumers);  // [CP]
  named_bar_sync(3, kConsumers);
  if (tid < CP) {
    float4 acc = part4[tid];
    for (int q = 1; q < slices; ++q) {
      const float4 o = part4[q * CP + tid];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    fin4[tid] = acc;
  }
  named_bar_sync(3, kConsumers);
  return scratch + kConsumers;
}

// Split of an image's chunks over its gx blocks.  Chunks of rows that involve a mirrored row cost more (an extra global
// round trip per chunk for a folded source row, doubled stores for a reflected destination row); with a plain even
// split the first and the last block of every image ran 1.3-1.6x as long as the others and set the pass's duration
// (tests/gpu/rows_probe.cu "block x -> cycles").  cost(c) = chunks before c, mirrored-row chunks counted wm times;
// block x owns [first chunk with cost >= x T / gx, first chunk with cost >= (x + 1) T / gx).
__device__ __forceinline__ int chunk_cost_prefix(int c, int H, int cpr, int f, int wm) {
  const int i = c / cpr, r = c - i * cpr;
  const int nm = min(max(i - 1, 0), f) + min(max(i - (H - 1 - f), 0), f);  // mirrored rows above row i
  const bool mi = (i >= 1 && i <= f) || (i >= H - 1 - f && i <= H - 2);
  return cpr * (i + (wm - 1) * nm) + (mi ? wm : 1) * r;
}
__device__ __forceinline__ int chunk_boundary(int x, int gx, int H, int cpr, int f, int wm) {
  const int nchunks = H * cpr;
  if (x <= 0) return 0;
  if (x >= gx) return nchunks;
  if (f <= 0 || wm <= 1 || H < 2 * f + 3) return int((int64_t(x) * nchunks + gx - 1) / gx);  // even split
  const int T = chunk_cost_prefix(nchunks, H, cpr, f, wm);
  const int target = int((int64_t(x) * T + gx - 1) / gx);
  int lo = 0, hi = nchunks;  // smallest c with prefix(c) >= target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (chunk_cost_prefix(mid, H, cpr, f, wm) >= target) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// =======================================================================================================
template <int MODE, int NS>
__global__ void __launch_bounds__(kStreamThreads, 1) row_stream_kernel(const RowStreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // pointer arithmetic keeps the shared address space
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  __shared__ uint64_t go_bar;  // consumers -> producer: the prologue's small loads are issued, start the bulk stream
  __shared__ ChunkDesc descs[kStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ba = b < p.nb_act ? b : b - p.act_wrap;
  const int cpr = (p.W + p.CW - 1) / p.CW;  // chunks per row
  // each block owns a contiguous range of chunks, so consecutive chunks mostly share their image row
  const int cbeg = chunk_boundary(blockIdx.x, gridDim.x, p.H, cpr, p.mrows, p.wm);
  const int cend = chunk_boundary(blockIdx.x + 1, gridDim.x, p.H, cpr, p.mrows, p.wm);
  const int dkind = (MODE == RS_GATHER || MODE == RS_BWD_REDUCE) ? 0 : p.dmap.kind;
  constexpr int kGradStream0 = MODE == RS_GATHER ? 0 : 1;  // stream index of gradient source 0 (stream 0 is Y otherwise)
  const int nst = p.stages;                    // pipeline stages in use
  const int nslot = NS + p.mslot;              // slots per stage (the last one holds a mirrored source row)

  RS_STAMP(0, threadIdx.x == 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumers / 32);
    }
    mbar_init(&go_bar, kConsumers / 32);
    fence_barrier_init();
  }
  pdl_launch_dependents();
  pdl_wait();  // barrier set-up overlapped the previous kernel's tail; global memory is touched only below
  __syncthreads();
  RS_STAMP(1, threadIdx.x == 0);

  if (warp == kConsumers / 32) {
    // ------------------------------------------------------------ producer: bookkeeping + bulk copies, `nst` ahead
    // (one thread; spreading the per-chunk bookkeeping over the lanes of the warp was measured and is slower: the
    // passes are bound by the consumers' per-chunk latency, not by this thread)
    if (elect_one_sync()) {
      const int src_band = (MODE == RS_APPLY) ? 0 : max(p.g[0].ptr ? p.g[0].fold : 0, p.g[1].ptr ? p.g[1].fold : 0);
      const int dst_band = ((MODE == RS_APPLY || MODE == RS_BWD_APPLY) && dkind == 0) ? p.dmap.reflect : 0;
      const int band = max(src_band, dst_band);
      const uint64_t pol = l2_policy_evict_first();  // the streams are read once per pass
      mbar_wait(&go_bar, 0, 40);
      ChunkDesc row;  // row-level part, recomputed when the image row changes
      int cur_i = -1;
      int k = 0;
      int i = cbeg / cpr, jc = cbeg - i * cpr;
      for (int c = cbeg; c < cend; ++c, ++k) {
        const int s = k % nst;
        if (i != cur_i) {
          cur_i = i;
          row.dn = 1; row.n1 = row.n2 = 0;
          row.lead[0] = row.lead[1] = 0;
          int m0, m1;
          if (MODE == RS_GATHER || (MODE == RS_BWD_REDUCE && p.gather_dst != nullptr)) {
            row.dbase[0] = (int64_t(b) * p.H + i) * p.W * p.C;
          } else if (MODE == RS_APPLY || MODE == RS_BWD_APPLY) {
            const FrameMap& m = p.dmap;
            const int64_t img = int64_t(b) * m.frame_pix * m.C;
            if (m.kind == 0) {
              const int nm = mirrors(i, m.H, m.reflect, m0, m1);
              row.dn = 1 + nm;
              row.dbase[0] = img + (int64_t(i + m.pt) * m.P + m.pl) * m.C;
              if (nm > 0) row.dbase[1] = img + (int64_t(m0 + m.pt) * m.P + m.pl) * m.C;
              if (nm > 1) row.dbase[2] = img + (int64_t(m1 + m.pt) * m.P + m.pl) * m.C;
            } else {
              const int64_t r = int64_t((i >> 1) + m.pt) * m.P + m.pl;
              row.dbase[0] = img + (int64_t((i & 1) * 2) * m.plane_pix + r) * m.C;
              row.dbase[1] = img + (int64_t((i & 1) * 2 + 1) * m.plane_pix + r) * m.C;
            }
          }
          if (MODE != RS_APPLY) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const GradSrc& g = p.g[q];
              long long* rows = q ? row.s2 : row.s1;
              int n = 0;
              if (g.ptr != nullptr) {
                const int64_t img = int64_t(b) * g.Hs * g.Ws * p.C;
                const int nm = mirrors(i, p.H, g.fold, m0, m1);
                n = 1 + nm;
                rows[0] = img + (int64_t(i + g.oy) * g.Ws + g.ox) * p.C;
                if (nm > 0) rows[1] = img + (int64_t(m0 + g.oy) * g.Ws + g.ox) * p.C;
                if (nm > 1) rows[2] = img + (int64_t(m1 + g.oy) * g.Ws + g.ox) * p.C;
              }
              if (q) row.n2 = n; else row.n1 = n;
            }
          }
        }
        mbar_wait(&empty_bar[s], ((k / nst) & 1) ^ 1, 41);
        const int j0 = jc * p.CW;
        const int cw = min(p.CW, p.W - j0);
        int lo = 0, hi = cw;
        if (row.dn > 1 || row.n1 > 1 || row.n2 > 1) {
          hi = 0;
        } else if (band > 0) {
          lo = min(cw, max(0, band + 1 - j0));
          hi = max(lo, min(cw, p.W - 1 - band - j0));
        }
        row.i = i; row.j0 = j0; row.cw = cw; row.lo = lo; row.hi = hi;
        // gradient sources with a reflected border: the first / last chunk of a row also stages the border columns, so
        // the consumers fold them in out of shared memory (a global load inside the consumer loop waits behind
        // everything the pipeline has in flight: ~7 us each)
        int lead[NS], tail[NS];
        uint32_t total = 0;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          const int gq = q - kGradStream0;  // stream q carries gradient source gq (or none)
          const int f = (MODE != RS_APPLY && gq >= 0 && gq < 2 && p.g[gq].ptr != nullptr) ? p.g[gq].fold : 0;
          lead[q] = jc == 0 ? f : 0;
          tail[q] = jc == cpr - 1 ? f : 0;
          if (gq >= 0 && gq < 2) row.lead[gq] = lead[q];
          total += uint32_t(cw + lead[q] + tail[q]) * p.C * 2;
        }
        // a row that folds a mirrored source row in: that row's segment travels in the stage's extra slot
        const int mq = (MODE == RS_APPLY || !p.mslot) ? -1 : (row.n1 > 1 ? 0 : (row.n2 > 1 ? 1 : -1));
        const int mqs = mq + kGradStream0;  // its stream index
        if (mq >= 0) total += uint32_t(cw + lead[mqs] + tail[mqs]) * p.C * 2;
        descs[s] = row;
        mbar_arrive_expect_tx(&full_bar[s], total);  // release: the descriptor is visible with the data
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          const StreamDesc& d = p.s[q];
          const sg_bf16* src = d.base + int64_t(d.act_index ? ba : b) * d.img_stride + int64_t(i + d.oy) * d.row_stride +
                               int64_t(j0 + d.ox - lead[q]) * p.C;
          const uint32_t bytes = uint32_t(cw + lead[q] + tail[q]) * p.C * 2;
          if (p.evict_first) bulk_load_1d_hint(smem + (s * nslot + q) * p.chunk_bytes, src, bytes, &full_bar[s], pol);
          else bulk_load_1d(smem + (s * nslot + q) * p.chunk_bytes, src, bytes, &full_bar[s]);
          if (mq >= 0 && q == mqs) {  // same columns of the first mirrored row
            const long long* mrow = mq ? row.s2 : row.s1;
            const sg_bf16* msrc = reinterpret_cast<const sg_bf16*>(p.g[mq].ptr) + mrow[1] + int64_t(j0 - lead[q]) * p.C;
            bulk_load_1d(smem + (s * nslot + NS) * p.chunk_bytes, msrc, bytes, &full_bar[s]);
          }
        }
        if (++jc == cpr) { jc = 0; ++i; }
      }
      RS_STAMP(6, true);
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
  // Per-channel coefficients, computed cooperatively: one thread per channel does the (independent) global loads and
  // the arithmetic, the results are staged in shared memory and every thread then picks up its eight channels.  (Each
  // thread loading its own 8 x 3 values cost ~14k cycles of serialised load latency per launch, a third of the pass.)
  float* coef = reinterpret_cast<float*>(smem + kPipeBytes);                  // [6][C]: mean, rstd, scale, beta, a1, a2
  float2* sum_scratch = reinterpret_cast<float2*>(smem + kPipeBytes + kCoefBytes);
  // ---- round 1: issue the global loads of the prologue
  const bool do_fin = MODE == RS_APPLY && p.stats_part != nullptr;
  const float2* sum_src = do_fin ? reinterpret_cast<const float2*>(p.stats_part) + int64_t(ba) * p.stats_T * p.C
                                 : (MODE == RS_BWD_APPLY ? reinterpret_cast<const float2*>(p.sums_part) + int64_t(b) * p.sums_nblk * p.C
                                                         : nullptr);
  const int sum_T = do_fin ? p.stats_T : (MODE == RS_BWD_APPLY ? p.sums_nblk : 0);
  SumPre pre;
  if (do_fin || MODE == RS_BWD_APPLY) fixed_order_sum_issue(sum_src, sum_T, p.C, threadIdx.x, pre);
  float2 st_c = make_float2(0.f, 0.f);
  float g_c = 1.f, be_c = 0.f;
  const bool have_st = !do_fin && p.stats != nullptr;
  if (MODE != RS_GATHER && int(threadIdx.x) < p.C) {  // C <= kMaxC = kConsumers: one channel per thread
    const int c = threadIdx.x;
    if (have_st) st_c = ld_nc_f2(reinterpret_cast<const float2*>(p.stats) + int64_t(ba) * p.C + c);
    if (p.gamma != nullptr) g_c = ld_nc_f1(p.gamma + c);
    if (p.beta != nullptr) be_c = ld_nc_f1(p.beta + c);
  }
  // ---- round 2: the loads have landed -> release the producer's bulk stream, then the fixed-order sums
  // (statistics finalize / backward partial sums) and the per-channel coefficients.  The bulk stream starts one or
  // two L2 round trips late; started first, the ~28 MB that all blocks request at once delay these small loads by
  // 4-8 us (measured with tests/gpu/rows_probe.cu), a third of the pass.
  const bool has_sum = do_fin || MODE == RS_BWD_APPLY;
  if (has_sum) fixed_order_sum_accumulate(sum_src, sum_T, p.C, sum_scratch, threadIdx.x, pre);
  else if (MODE != RS_GATHER && int(threadIdx.x) < p.C) coef[3 * p.C + threadIdx.x] = be_c + 0.f * (st_c.x + g_c);  // data dependency
  __syncwarp();
  if (lane == 0) mbar_arrive(&go_bar);
  const float2* fin = nullptr;
  if (do_fin) {
    // fused statistics finalize: the per-tile partials of the producing convolution are added in a fixed order;
    // block 0 of the image also publishes the result for the backward pass
    fin = fixed_order_sum_combine(p.C, sum_scratch, threadIdx.x);
    if (blockIdx.x == 0)
      for (int c = threadIdx.x; c < p.C; c += kConsumers) reinterpret_cast<float2*>(p.stats_out)[int64_t(ba) * p.C + c] = fin[c];
  }
  const float2* bsum = nullptr;
  if (MODE == RS_BWD_APPLY) {
    // the reduce pass left one partial (sum dzh, sum dzh * xhat) per block of this image
    bsum = fixed_order_sum_combine(p.C, sum_scratch, threadIdx.x);
    if (blockIdx.x == 0 && p.sums != nullptr)
      for (int c = threadIdx.x; c < p.C; c += kConsumers) reinterpret_cast<float2*>(p.sums)[int64_t(b) * p.C + c] = bsum[c];
  }
  if (MODE != RS_GATHER) {
    if (int(threadIdx.x) < p.C) {
      const int c = threadIdx.x;
      float mu = 0.f, rs = 1.f;
      if (fin != nullptr || have_st) {
        const float2 st = fin != nullptr ? fin[c] : st_c;
        mu = st.x / n;
        rs = rsqrtf(fmaxf(st.y / n - mu * mu, 0.f) + p.eps);
      }
      coef[c] = mu;
      coef[p.C + c] = rs;
      coef[2 * p.C + c] = g_c * rs;
      coef[3 * p.C + c] = be_c;  // z = (y - mean)*scale + beta: exactly beta when H*W == 1
      float q1 = 0.f, q2 = 0.f;
      if (MODE == RS_BWD_APPLY) {
        q1 = bsum[c].x / n;        // mean of dzh
        q2 = bsum[c].y / n * rs;   // mean of dzh * xhat, times rstd (applied to y - mean directly)
      }
      coef[4 * p.C + c] = q1;
      coef[5 * p.C + c] = q2;
    }
    named_bar_sync(3, kConsumers);
  }
  // activation as a slope for the non-positive side: relu 0, leaky alpha, identity 1 (tanh never reaches the glue)
  const float gneg = p.act == SG_ACT_RELU ? 0.f : (p.act == SG_ACT_LRELU ? p.alpha : 1.f);
  float mean[8], rstd[8], scale[8], beta[8], a1[8], a2[8];
  if (MODE != RS_GATHER) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      mean[e] = coef[c];
      rstd[e] = coef[p.C + c];
      scale[e] = coef[2 * p.C + c];
      beta[e] = coef[3 * p.C + c];
      a1[e] = coef[4 * p.C + c];
      a2[e] = coef[5 * p.C + c];
    }
  }
  RS_STAMP(2, threadIdx.x == 0);
  const uint32_t chunk_bytes = p.chunk_bytes, stage_bytes = nslot * p.chunk_bytes;
  const int W = p.W, C = p.C;
  const int dC = (MODE == RS_GATHER || MODE == RS_BWD_REDUCE) ? C : p.dmap.C;
  const int pshift = 31 - __clz(pstep);  // pstep = 512 / C8 is a power of two
  const uint32_t sbase = smem_u32(smem) + threadIdx.x * 16;
  const bool gather = MODE == RS_BWD_REDUCE && p.gather_dst != nullptr;
  const bool has_act = gneg != 1.f;
  sg_bf16* const dst_c0 = (gather ? p.gather_dst : p.dst) + c0;
  const ChunkDesc* cd = nullptr;

  // one pixel (8 channels) of this thread: `sa` = its vector in stream 0 of the stage, `dptr` = where the lean
  // (no border bookkeeping) result goes, `j` = image column
  uint32_t lb0 = 0, lb1 = 0;  // byte offsets of the chunk's first pixel inside the staged rows of gradient source 0 / 1
  auto pixel = [&](const bool lean, const uint32_t sa, sg_bf16* dptr, const int j, const int mq, const uint32_t moff) {
    // shared-memory addresses of this pixel in the streams of gradient source 0 / 1
    const uint32_t sg0 = sa + kGradStream0 * chunk_bytes + lb0, sg1 = sa + (kGradStream0 + 1) * chunk_bytes + lb1;
    uint4 r0 = lds128(MODE == RS_GATHER ? sg0 : sa), r1 = make_uint4(0, 0, 0, 0), r2 = r1;
    if (NS > 1) r1 = lds128(MODE == RS_APPLY ? sa + chunk_bytes : (MODE == RS_GATHER ? sg1 : sg0));
    if (NS > 2) r2 = lds128(sg1);
    float y[8], d[8];
    if (MODE == RS_APPLY) {
      unpack8(r0, y);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = fmaf(y[e] - mean[e], scale[e], beta[e]);
        y[e] = z > 0.f ? z : z * gneg;
      }
      if (NS > 1) {
        unpack8(r1, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] += d[e];
      }
      if (lean) *reinterpret_cast<uint4*>(dptr) = pack8(y);
      else dst_store8(cd, p.dmap, dst_c0, j, pack8(y));
    } else if (MODE == RS_GATHER) {
      unpack8(r0, d);
      unpack8(r1, y);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] += y[e];
      if (!lean) {
        if (src_has_extra(cd->n1, p.g[0], j, W)) src_extra8(cd->s1, cd->n1, p.g[0], j, W, C, c0, d, sg0, mq == 0, sa + moff);
        if (src_has_extra(cd->n2, p.g[1], j, W)) src_extra8(cd->s2, cd->n2, p.g[1], j, W, C, c0, d, sg1, mq == 1, sa + moff);
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
        if (src_has_extra(cd->n1, p.g[0], j, W)) src_extra8(cd->s1, cd->n1, p.g[0], j, W, C, c0, d, sg0, mq == 0, sa + moff);
        if (src_has_extra(cd->n2, p.g[1], j, W)) src_extra8(cd->s2, cd->n2, p.g[1], j, W, C, c0, d, sg1, mq == 1, sa + moff);
      }
      if (MODE == RS_BWD_REDUCE && gather) {
        // fused residual-gradient gather: store the sum, and take the statistics over the value AS STORED so that
        // the apply pass (which reads it back) stays exactly consistent with these sums
        const uint4 w = pack8(d);
        *reinterpret_cast<uint4*>(dptr) = w;
        unpack8(w, d);
      }
      // zpre = (y - mean)*scale + beta;  xhat = (y - mean)*rstd.  rstd is folded into the channel constants: the
      // reduce pass accumulates sum(dzh * (y - mean)) and scales it once at the end, the apply pass gets
      // a2 = mean(dzh * xhat) * rstd.  The apply form keeps a single-pixel norm (y == mean, dzh == a1)
      // back-propagating exactly zero, as the reference does at 128x128 (Appendix B).  Layers without an
      // activation (gneg == 1: the second norm of every residual block) skip the mask altogether.
      if (has_act) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float yc = y[e] - mean[e];
          const float dz = fmaf(yc, scale[e], beta[e]) > 0.f ? d[e] : d[e] * gneg;
          if (MODE == RS_BWD_APPLY) {
            d[e] = scale[e] * ((dz - a1[e]) - yc * a2[e]);
          } else {
            a1[e] += dz;
            a2[e] = fmaf(dz, yc, a2[e]);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float yc = y[e] - mean[e];
          if (MODE == RS_BWD_APPLY) {
            d[e] = scale[e] * ((d[e] - a1[e]) - yc * a2[e]);
          } else {
            a1[e] += d[e];
            a2[e] = fmaf(d[e], yc, a2[e]);
          }
        }
      }
      if (MODE == RS_BWD_APPLY) {
        if (lean) *reinterpret_cast<uint4*>(dptr) = pack8(d);
        else dst_store8(cd, p.dmap, dst_c0, j, pack8(d));
      }
    }
  };

  int k = 0;
  for (int c = cbeg; c < cend; ++c, ++k) {
    const int s = k % nst;
    cd = &descs[s];
    mbar_wait(&full_bar[s], (k / nst) & 1, 42);
    RS_STAMP(3, threadIdx.x == 0 && k == 0);
    const int4 h0 = *reinterpret_cast<const int4*>(cd);      // i, j0, cw, lo
    const int hi = cd->hi;
    const int j0 = h0.y, cw = h0.z, lo = h0.w;
    lb0 = uint32_t(cd->lead[0]) * C * 2;
    lb1 = uint32_t(cd->lead[1]) * C * 2;
    // this thread visits pixels px0 + t * pstep, t in [0, nt); t in [tlo, thi) are lean
    const int nt = cw > px0 ? ((cw - px0 + pstep - 1) >> pshift) : 0;
    const int tlo = min(nt, lo > px0 ? ((lo - px0 + pstep - 1) >> pshift) : 0);
    const int thi = max(tlo, min(nt, hi > px0 ? ((hi - px0 + pstep - 1) >> pshift) : 0));
    // incremental destination pointer of this thread (element units)
    int j = j0 + px0;
    sg_bf16* dptr;
    int dstep;
    if (MODE == RS_GATHER || MODE == RS_BWD_REDUCE || dkind == 0) {
      dptr = dst_c0 + cd->dbase[0] + int64_t(j) * dC;
      dstep = pstep * dC;
    } else {  // phase planes: the column parity of this thread is fixed because pstep is even
      dptr = dst_c0 + cd->dbase[j & 1] + int64_t(j >> 1) * dC;
      dstep = (pstep >> 1) * dC;
    }
    uint32_t sa = sbase + s * stage_bytes;
    // a source folds a mirrored row in: its vectors sit in the stage's extra slot, same layout as the primary row
    const int mq = (MODE == RS_APPLY || !p.mslot) ? -1 : (cd->n1 > 1 ? 0 : (cd->n2 > 1 ? 1 : -1));
    const uint32_t moff = NS * chunk_bytes + (mq == 1 ? lb1 : lb0);  // from `sa` to the pixel's vector in the mirror slot
    int t = 0;
    for (; t < tlo; ++t, sa += kConsumers * 16, dptr += dstep, j += pstep) pixel(false, sa, dptr, j, mq, moff);
#pragma unroll 2
    for (; t < thi; ++t, sa += kConsumers * 16, dptr += dstep, j += pstep) pixel(true, sa, dptr, j, -1, 0u);
    for (; t < nt; ++t, sa += kConsumers * 16, dptr += dstep, j += pstep) pixel(false, sa, dptr, j, mq, moff);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  RS_STAMP(4, threadIdx.x == 0);
  if (MODE == RS_BWD_REDUCE) {
    // threads that share a channel group differ in px0 (kConsumers / C8 of them): stage their partials in the
    // (now idle) pipeline buffers as [px0][e][k][cg] (consecutive lanes -> consecutive words) and add them
    // in a fixed order
    float* red = reinterpret_cast<float*>(smem);
    named_bar_sync(2, kConsumers);  // every consumer is past its last chunk: the stage buffers are free
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[((px0 * 8 + e) * 2 + 0) * C8 + cg] = a1[e];
      red[((px0 * 8 + e) * 2 + 1) * C8 + cg] = a2[e] * rstd[e];  // sum(dzh * (y - mean)) -> sum(dzh * xhat)
    }
    named_bar_sync(2, kConsumers);
    for (int t = threadIdx.x; t < 2 * C; t += kConsumers) {
      const int ch = t >> 1, kk = t & 1;
      const int idx = (((ch & 7) * 2) + kk) * C8 + (ch >> 3);
      float acc = 0.f;
      for (int q = 0; q < pstep; ++q) acc += red[q * 16 * C8 + idx];
      p.sums_part[(int64_t(b) * gridDim.x + blockIdx.x) * C * 2 + t] = acc;  // blocks without chunks still write zeros
    }
  }
  RS_STAMP(5, threadIdx.x == 0);
}

// =======================================================================================================
// Fused instance-norm backward: reduce AND apply of one layer in ONE launch.
//
//   pass 1 (forward over the block's chunks):  dzh = act'(.) * (g1 + g2),  partial (sum dzh, sum dzh * xhat); the
//           residual-stream gradient g1 + g2 is stored on the way when asked for;
//   image barrier: the blocks of an image publish their partials (plain stores) and meet on an arrival counter --
//           every block of the launch is resident (one block per SM, grid <= 148), so this cannot deadlock;
//   pass 2 (BACKWARD over the same chunks):    dy = gamma * rstd * (dzh - mean(dzh) - xhat * mean(dzh * xhat)) -> dY frame.
//
// The last kStages chunks of pass 1 are still in the pipeline's shared memory when pass 2 starts, and pass 2 consumes
// them first, so ~40 % of the layer is never read a second time; the rest is re-read most-recent-first (what an LRU
// L2 can still hold).  Against the two-launch form this saves a launch, a prologue, and that part of the traffic:
// 59 us -> see tests/gpu/rows_probe.cu.  Deterministic: partials are added in a fixed order, no atomics on data.
template <int NS>
__global__ void __launch_bounds__(kStreamThreads, 1) row_fused_bwd_kernel(const RowStreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  __shared__ uint64_t go_bar;
  __shared__ ChunkDesc descs[kStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ba = b < p.nb_act ? b : b - p.act_wrap;
  const int cpr = (p.W + p.CW - 1) / p.CW;
  const int cbeg = chunk_boundary(blockIdx.x, gridDim.x, p.H, cpr, p.mrows, p.wm);
  const int cend = chunk_boundary(blockIdx.x + 1, gridDim.x, p.H, cpr, p.mrows, p.wm);
  const int n = max(0, cend - cbeg);          // chunks of this block
  const int S = min(p.stages, max(n, 1));     // stages in use; the last S chunks of pass 1 stay resident for pass 2
  const int nslot = NS + p.mslot;             // slots per stage (the last one holds a mirrored source row)
  const int dkind = p.dmap.kind;
  const bool gather = p.gather_dst != nullptr;

  RS_STAMP(0, threadIdx.x == 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumers / 32);
    }
    mbar_init(&go_bar, kConsumers / 32);
    fence_barrier_init();
  }
  pdl_launch_dependents();
  pdl_wait();
  __syncthreads();
  RS_STAMP(1, threadIdx.x == 0);

  if (warp == kConsumers / 32) {
    // ------------------------------------------------------------ producer
    if (lane == 0 && n > 0) {
      const int band = max(p.g[0].ptr ? p.g[0].fold : 0, p.g[1].ptr ? p.g[1].fold : 0);
      const uint64_t pol = l2_policy_evict_first();
      mbar_wait(&go_bar, 0, 40);
      ChunkDesc row;
      int cur_i = -1;
      uint32_t epar = 0xffffffffu;  // per-stage parity of the next empty-barrier wait (a fresh barrier passes parity 1)
      const int total = n + max(0, n - S);
      for (int k = 0; k < total; ++k) {
        // load k: pass 1 walks the chunks forward; pass 2 re-loads the non-resident ones backward into the stages
        // in the order the consumers free them
        const int cc = k < n ? k : n - S - 1 - (k - n);
        const int s = k < n ? k % S : (n - 1 - (k - n)) % S;
        const bool second = k >= n;
        const int c = cbeg + cc;
        const int i = c / cpr, jc = c - i * cpr;
        if (i != cur_i) {
          cur_i = i;
          row.dn = 1; row.n1 = row.n2 = 0;
          row.lead[0] = row.lead[1] = 0;
          int m0, m1;
          const FrameMap& m = p.dmap;  // dY frames carry no reflected border (zero borders stay zero)
          const int64_t img = int64_t(b) * m.frame_pix * m.C;
          if (m.kind == 0) {
            row.dbase[0] = img + (int64_t(i + m.pt) * m.P + m.pl) * m.C;
            row.dbase[1] = row.dbase[0];
          } else {
            const int64_t r = int64_t((i >> 1) + m.pt) * m.P + m.pl;
            row.dbase[0] = img + (int64_t((i & 1) * 2) * m.plane_pix + r) * m.C;
            row.dbase[1] = img + (int64_t((i & 1) * 2 + 1) * m.plane_pix + r) * m.C;
          }
          row.dbase[2] = (int64_t(b) * p.H + i) * p.W * p.C;  // row of the gathered residual-stream gradient
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const GradSrc& g = p.g[q];
            long long* rows = q ? row.s2 : row.s1;
            int nn = 0;
            if (g.ptr != nullptr) {
              const int64_t gimg = int64_t(b) * g.Hs * g.Ws * p.C;
              const int nm = mirrors(i, p.H, g.fold, m0, m1);
              nn = 1 + nm;
              rows[0] = gimg + (int64_t(i + g.oy) * g.Ws + g.ox) * p.C;
              if (nm > 0) rows[1] = gimg + (int64_t(m0 + g.oy) * g.Ws + g.ox) * p.C;
              if (nm > 1) rows[2] = gimg + (int64_t(m1 + g.oy) * g.Ws + g.ox) * p.C;
            }
            if (q) row.n2 = nn; else row.n1 = nn;
          }
        }
        mbar_wait(&empty_bar[s], (epar >> s) & 1u, 41);
        epar ^= 1u << s;
        const int j0 = jc * p.CW;
        const int cw = min(p.CW, p.W - j0);
        int lo = 0, hi = cw;
        if (row.n1 > 1 || row.n2 > 1) {
          hi = 0;
        } else if (band > 0) {
          lo = min(cw, max(0, band + 1 - j0));
          hi = max(lo, min(cw, p.W - 1 - band - j0));
        }
        row.i = i; row.j0 = j0; row.cw = cw; row.lo = lo; row.hi = hi;
        int lead[NS], tail[NS];  // reflected border columns staged with the first / last chunk of a row (see row_stream_kernel)
        uint32_t total = 0;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          const int f = (q >= 1 && p.g[q - 1].ptr != nullptr) ? p.g[q - 1].fold : 0;
          lead[q] = jc == 0 ? f : 0;
          tail[q] = jc == cpr - 1 ? f : 0;
          if (q >= 1) row.lead[q - 1] = lead[q];
          total += uint32_t(cw + lead[q] + tail[q]) * p.C * 2;
        }
        const int mq = !p.mslot ? -1 : (row.n1 > 1 ? 0 : (row.n2 > 1 ? 1 : -1));  // a mirrored source row: the extra slot
        const int mqs = mq + 1;
        if (mq >= 0) total += uint32_t(cw + lead[mqs < NS ? mqs : 0] + tail[mqs < NS ? mqs : 0]) * p.C * 2;
        descs[s] = row;
        mbar_arrive_expect_tx(&full_bar[s], total);
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          const StreamDesc& d = p.s[q];
          const sg_bf16* src = d.base + int64_t(d.act_index ? ba : b) * d.img_stride + int64_t(i + d.oy) * d.row_stride +
                               int64_t(j0 + d.ox - lead[q]) * p.C;
          const uint32_t bytes = uint32_t(cw + lead[q] + tail[q]) * p.C * 2;
          // pass 1 data is read again by pass 2: default policy; pass 2 is the last reader: evict first
          if (second && p.evict_first) bulk_load_1d_hint(smem + (s * nslot + q) * p.chunk_bytes, src, bytes, &full_bar[s], pol);
          else bulk_load_1d(smem + (s * nslot + q) * p.chunk_bytes, src, bytes, &full_bar[s]);
          if (mq >= 0 && q == mqs) {
            const long long* mrow = mq ? row.s2 : row.s1;
            const sg_bf16* msrc = reinterpret_cast<const sg_bf16*>(p.g[mq].ptr) + mrow[1] + int64_t(j0 - lead[q]) * p.C;
            bulk_load_1d(smem + (s * nslot + NS) * p.chunk_bytes, msrc, bytes, &full_bar[s]);
          }
        }
      }
      RS_STAMP(6, true);
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int C8 = p.C >> 3;
  const int cg = threadIdx.x % C8;
  const int c0 = cg * 8;
  const int px0 = threadIdx.x / C8, pstep = kConsumers / C8;
  const float npix = float(p.H * p.W);
  float* coef = reinterpret_cast<float*>(smem + kPipeBytes);  // [6][C]: mean, rstd, scale, beta, a1, a2
  float2* sum_scratch = reinterpret_cast<float2*>(smem + kPipeBytes + kCoefBytes);
  // ---- prologue: per-channel constants of the forward pass (one channel per thread), then let the producer go
  {
    float2 st_c = make_float2(0.f, 0.f);
    float g_c = 1.f, be_c = 0.f;
    if (int(threadIdx.x) < p.C) {
      const int c = threadIdx.x;
      if (p.stats != nullptr) st_c = ld_nc_f2(reinterpret_cast<const float2*>(p.stats) + int64_t(ba) * p.C + c);
      if (p.gamma != nullptr) g_c = ld_nc_f1(p.gamma + c);
      if (p.beta != nullptr) be_c = ld_nc_f1(p.beta + c);
      float mu = 0.f, rs = 1.f;
      if (p.stats != nullptr) {
        mu = st_c.x / npix;
        rs = rsqrtf(fmaxf(st_c.y / npix - mu * mu, 0.f) + p.eps);
      }
      coef[c] = mu;
      coef[p.C + c] = rs;
      coef[2 * p.C + c] = g_c * rs;
      coef[3 * p.C + c] = be_c;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&go_bar);  // after the (data-dependent) stores above: the small loads have landed
    named_bar_sync(3, kConsumers);
  }
  const float gneg = p.act == SG_ACT_RELU ? 0.f : (p.act == SG_ACT_LRELU ? p.alpha : 1.f);
  float mean[8], rstd[8], scale[8], beta[8], a1[8], a2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    mean[e] = coef[c];
    rstd[e] = coef[p.C + c];
    scale[e] = coef[2 * p.C + c];
    beta[e] = coef[3 * p.C + c];
    a1[e] = a2[e] = 0.f;
  }
  RS_STAMP(2, threadIdx.x == 0);
  const uint32_t chunk_bytes = p.chunk_bytes, stage_bytes = nslot * p.chunk_bytes;
  const int W = p.W, C = p.C, dC = p.dmap.C;
  const int pshift = 31 - __clz(pstep);
  const uint32_t sbase = smem_u32(smem) + threadIdx.x * 16;
  const bool has_act = gneg != 1.f;
  sg_bf16* const gat_c0 = gather ? p.gather_dst + c0 : nullptr;
  sg_bf16* const dst_c0 = p.dst + c0;
  const ChunkDesc* cd = nullptr;

  // one pixel (8 channels): PH 1 = reduce pass (accumulates a1 / a2, stores the gathered gradient at gptr),
  // PH 2 = apply pass (a1 / a2 hold the image means; stores dy at dptr)
  uint32_t lb0 = 0, lb1 = 0;  // byte offsets of the chunk's first pixel inside the staged rows of gradient source 0 / 1
  auto pixel = [&](auto PHc, const bool lean, const uint32_t sa, sg_bf16* ptr, const int j, const int mq,
                   const uint32_t moff) {
    constexpr int PH = decltype(PHc)::value;
    const uint32_t sg0 = sa + chunk_bytes + lb0, sg1 = sa + 2 * chunk_bytes + lb1;
    const uint4 r0 = lds128(sa), r1 = lds128(sg0);
    float y[8], d[8];
    unpack8(r0, y);
    unpack8(r1, d);
    if (NS > 2) {
      float t[8];
      unpack8(lds128(sg1), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] += t[e];
    }
    if (!lean) {
      if (src_has_extra(cd->n1, p.g[0], j, W)) src_extra8(cd->s1, cd->n1, p.g[0], j, W, C, c0, d, sg0, mq == 0, sa + moff);
      if (src_has_extra(cd->n2, p.g[1], j, W)) src_extra8(cd->s2, cd->n2, p.g[1], j, W, C, c0, d, sg1, mq == 1, sa + moff);
    }
    if (gather) {  // the residual-stream gradient as stored (bf16): both passes work on exactly this value
      const uint4 w = pack8(d);
      if (PH == 1) *reinterpret_cast<uint4*>(ptr) = w;
      unpack8(w, d);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yc = y[e] - mean[e];
      float dz = d[e];
      if (has_act) dz = fmaf(yc, scale[e], beta[e]) > 0.f ? dz : dz * gneg;
      if (PH == 2) {
        d[e] = scale[e] * ((dz - a1[e]) - yc * a2[e]);
      } else {
        a1[e] += dz;
        a2[e] = fmaf(dz, yc, a2[e]);
      }
    }
    if (PH == 2) *reinterpret_cast<uint4*>(ptr) = pack8(d);
  };
  auto chunk = [&](auto PHc, const int s) {
    constexpr int PH = decltype(PHc)::value;
    cd = &descs[s];
    const int4 h0 = *reinterpret_cast<const int4*>(cd);  // i, j0, cw, lo
    const int hi = cd->hi;
    const int j0 = h0.y, cw = h0.z, lo = h0.w;
    lb0 = uint32_t(cd->lead[0]) * C * 2;
    lb1 = uint32_t(cd->lead[1]) * C * 2;
    const int nt = cw > px0 ? ((cw - px0 + pstep - 1) >> pshift) : 0;
    const int tlo = min(nt, lo > px0 ? ((lo - px0 + pstep - 1) >> pshift) : 0);
    const int thi = max(tlo, min(nt, hi > px0 ? ((hi - px0 + pstep - 1) >> pshift) : 0));
    int j = j0 + px0;
    sg_bf16* ptr;
    int step;
    if (PH == 1) {  // gathered gradient: plain [B][H][W][C] (pointer unused when nothing is gathered)
      ptr = gat_c0 + cd->dbase[2] + int64_t(j) * C;
      step = pstep * C;
    } else if (dkind == 0) {
      ptr = dst_c0 + cd->dbase[0] + int64_t(j) * dC;
      step = pstep * dC;
    } else {  // phase planes: the column parity of this thread is fixed because pstep is even
      ptr = dst_c0 + cd->dbase[j & 1] + int64_t(j >> 1) * dC;
      step = (pstep >> 1) * dC;
    }
    uint32_t sa = sbase + s * stage_bytes;
    const int mq = !p.mslot ? -1 : (cd->n1 > 1 ? 0 : (cd->n2 > 1 ? 1 : -1));  // mirrored source row in the extra slot
    const uint32_t moff = NS * chunk_bytes + (mq == 1 ? lb1 : lb0);
    int t = 0;
    for (; t < tlo; ++t, sa += kConsumers * 16, ptr += step, j += pstep) pixel(PHc, false, sa, ptr, j, mq, moff);
#pragma unroll 2
    for (; t < thi; ++t, sa += kConsumers * 16, ptr += step, j += pstep) pixel(PHc, true, sa, ptr, j, -1, 0u);
    for (; t < nt; ++t, sa += kConsumers * 16, ptr += step, j += pstep) pixel(PHc, false, sa, ptr, j, mq, moff);
  };
  using PH1 = std::integral_constant<int, 1>;
  using PH2 = std::integral_constant<int, 2>;

  // ---- pass 1: forward over the chunks; the last S stay in shared memory
  uint32_t fpar = 0;  // per-stage parity of the next full-barrier completion
  for (int k = 0; k < n; ++k) {
    const int s = k % S;
    mbar_wait(&full_bar[s], (fpar >> s) & 1u, 42);
    fpar ^= 1u << s;
    RS_STAMP(3, threadIdx.x == 0 && k == 0);
    chunk(PH1{}, s);
    if (k < n - S) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
  RS_STAMP(4, threadIdx.x == 0);
  // ---- block partial sums, added in a fixed order through an 8 KB buffer (the pipeline memory is still in use):
  //      four rounds, each covering a quarter of the px0 groups
  {
    float* red = reinterpret_cast<float*>(sum_scratch);  // [R][8][2][C8] floats = 2048
    const int R = (pstep >> 2) > 0 ? (pstep >> 2) : 1;   // px0 groups per round (pstep is 8, 16, 32 or 64)
    const int rounds = (pstep + R - 1) / R;
    float acc0 = 0.f, acc1 = 0.f;  // thread t < 2C owns entry t (and t + kConsumers when 2C > kConsumers)
    for (int r = 0; r < rounds; ++r) {
      named_bar_sync(2, kConsumers);
      if (px0 / R == r) {
        const int q = px0 - r * R;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          red[((q * 8 + e) * 2 + 0) * C8 + cg] = a1[e];
          red[((q * 8 + e) * 2 + 1) * C8 + cg] = a2[e] * rstd[e];  // sum(dzh * (y - mean)) -> sum(dzh * xhat)
        }
      }
      named_bar_sync(2, kConsumers);
      for (int t = threadIdx.x, u = 0; t < 2 * C; t += kConsumers, ++u) {
        const int ch = t >> 1, kk = t & 1;
        const int idx = (((ch & 7) * 2) + kk) * C8 + (ch >> 3);
        float a = 0.f;
        for (int q = 0; q < R && r * R + q < pstep; ++q) a += red[q * 16 * C8 + idx];
        if (u == 0) acc0 += a; else acc1 += a;
      }
    }
    float* part = p.sums_part + (int64_t(b) * gridDim.x + blockIdx.x) * C * 2;
    for (int t = threadIdx.x, u = 0; t < 2 * C; t += kConsumers, ++u) part[t] = u == 0 ? acc0 : acc1;
  }
  // ---- image barrier: all blocks of this image have published their partials
  named_bar_sync(2, kConsumers);
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(p.sync_ctr + b, 1);
    const long long t0 = clock64();
    while (true) {
      int v;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.sync_ctr + b) : "memory");
      if (v >= int(gridDim.x)) break;
      __nanosleep(64);
      if (clock64() - t0 > (1ll << 31)) {  // a block of this image never arrived: fail loudly instead of hanging
        g_tc_watchdog_flag = 43;
        __threadfence_system();
        __trap();
      }
    }
  }
  named_bar_sync(2, kConsumers);
  if (n == 0) return;  // nothing to apply (the partial above was all zeros)
  {
    SumPre pre;
    const float2* src = reinterpret_cast<const float2*>(p.sums_part) + int64_t(b) * gridDim.x * p.C;
    fixed_order_sum_issue<true>(src, int(gridDim.x), p.C, threadIdx.x, pre);
    fixed_order_sum_accumulate<true>(src, int(gridDim.x), p.C, sum_scratch, threadIdx.x, pre);
    const float2* bsum = fixed_order_sum_combine(p.C, sum_scratch, threadIdx.x);
    if (int(threadIdx.x) < p.C) {
      const int c = threadIdx.x;
      const float2 q = bsum[c];
      if (blockIdx.x == 0 && p.sums != nullptr) reinterpret_cast<float2*>(p.sums)[int64_t(b) * p.C + c] = q;
      coef[4 * p.C + c] = q.x / npix;                      // mean of dzh
      coef[5 * p.C + c] = q.y / npix * coef[p.C + c];      // mean of dzh * xhat, times rstd
    }
    named_bar_sync(3, kConsumers);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a1[e] = coef[4 * p.C + c0 + e];
      a2[e] = coef[5 * p.C + c0 + e];
    }
  }
  // ---- pass 2: backward over the chunks; the first S are the ones pass 1 left in shared memory
  for (int j = 0; j < n; ++j) {
    const int s = (n - 1 - j) % S;
    if (j >= S) {
      mbar_wait(&full_bar[s], (fpar >> s) & 1u, 44);
      fpar ^= 1u << s;
    }
    chunk(PH2{}, s);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  RS_STAMP(5, threadIdx.x == 0);
}

// Returns the blocks per image of the launch (> 0), 0 if there was nothing to do, or -(cudaError) on failure.
template <int MODE, int NS>
static int launch_row_stream_ns(const RowStreamParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set[64] = {};  // the attribute is per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -int(e);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(row_stream_kernel<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return -int(e);
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  e = launch_kernel_pdl(row_stream_kernel<MODE, NS>, grid, dim3(kStreamThreads), smem, st, pdl_enabled(), p);
  return e == cudaSuccess ? int(grid.x) : -int(e);
}

// Chunk geometry and grid of a launch; returns 0, 1 if there is nothing to do, or a negative cudaError.
static int plan_row_stream(RowStreamParams& p, dim3* grid, size_t* smem) {
  p.ns = 0;
  while (p.ns < kMaxStreams && p.s[p.ns].base != nullptr) ++p.ns;  // active streams are a prefix
  if (p.ns == 0) return 1;
  if (p.C > kMaxC || (p.C & 63) != 0 || p.B < 1 || p.H < 1 || p.W < 1) return -int(cudaErrorInvalidValue);
  {
    // L2 policy of the streamed loads: evict-first.  Keeping the backward reduce pass's streams resident for the apply
    // pass that follows was measured (tests/gpu/rows_probe.cu, "reduce -> apply") and does not help: 67 MB streamed
    // cyclically do not survive in the 126 MB L2, and the reduce pass itself runs 15 % slower without the hint.
    static const int ef = []() { const char* e = getenv("SGGAN_ROWS_EVICT_FIRST"); return (e && e[0] == '0') ? 0 : 1; }();
    p.evict_first = ef;
  }
#ifdef SG_ROWS_DEBUG
  p.dbg_launch = g_rows_dbg_launch++;
#endif
  // chunk width: as wide as the pipeline memory allows, then evened out over the row (128 pixels at 48 per chunk
  // would be 48 + 48 + 32)
  const int maxfold = std::max(p.g[0].ptr ? p.g[0].fold : 0, p.g[1].ptr ? p.g[1].fold : 0);  // staged border columns
  // sources with a folded border: one more slot per stage for the mirrored row, three stages instead of four (measured
  // equal, tests/gpu/rows_probe.cu s3 vs s4) so that the chunks keep their width
  p.mslot = maxfold > 0 ? 1 : 0;
  p.stages = p.mslot ? std::min(3, kStages) : kStages;
  int cwmax = kPipeBytes / (p.stages * (p.ns + p.mslot)) / (p.C * 2) - 2 * maxfold;
  if (cwmax < 1) return -int(cudaErrorInvalidValue);
  if (cwmax > p.W) cwmax = p.W;
  const int cpr = (p.W + cwmax - 1) / cwmax;
  p.CW = (p.W + cpr - 1) / cpr;
  p.chunk_bytes = (p.CW + 2 * maxfold) * p.C * 2;
  *smem = size_t(kPipeBytes) + 128 + kCoefBytes + kSumBytes;
  // one persistent block per SM; blocks never span images (per-image statistics); no block without chunks
  int gx = 148 / p.B;
  if (gx < 1) gx = 1;
  const int nchunks = p.H * cpr;
  if (gx > nchunks) gx = nchunks;
  *grid = dim3(gx, p.B);
  {
    static const int wm = []() { const char* e = getenv("SGGAN_ROWS_WM"); return e ? atoi(e) : 2; }();
    const int dst_reflect = (p.dmap.kind == 0) ? p.dmap.reflect : 0;
    p.mrows = std::max(maxfold, dst_reflect);
    p.wm = wm < 1 ? 1 : wm;
  }
  return 0;
}

template <int MODE>
static int launch_row_stream(RowStreamParams& p, cudaStream_t st) {
  dim3 grid;
  size_t smem;
  const int r = plan_row_stream(p, &grid, &smem);
  if (r != 0) return r < 0 ? r : 0;
  if (p.ns == 1) return launch_row_stream_ns<MODE, 1>(p, grid, smem, st);
  if (p.ns == 2) return launch_row_stream_ns<MODE, 2>(p, grid, smem, st);
  return launch_row_stream_ns<MODE, 3>(p, grid, smem, st);
}

template <int NS>
static int launch_row_fused_ns(const RowStreamParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -int(e);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(row_fused_bwd_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return -int(e);
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  e = launch_kernel_pdl(row_fused_bwd_kernel<NS>, grid, dim3(kStreamThreads), smem, st, pdl_enabled(), p);
  return e == cudaSuccess ? 0 : -int(e);
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

int launch_in_apply(const InApplyParams& a, cudaStream_t st) {
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
  p.stats_part = a.stats_part; p.stats_T = a.stats_T; p.stats_out = a.stats_out;
  p.dst = a.dst; p.dmap = a.dmap;
  const int r = launch_row_stream<RS_APPLY>(p, st);
  return r < 0 ? r : 0;
}

static void bwd_params(const InBwdParams& a, RowStreamParams& p) {
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.nb_act = a.nb_act; p.act_wrap = a.act_wrap;
  p.s[0] = plain_stream(a.Y, a.H, a.W, a.C, 1);
  p.s[1] = a.g1.ptr ? grad_stream(a.g1, a.C) : null_stream();
  p.s[2] = a.g2.ptr ? grad_stream(a.g2, a.C) : null_stream();
  p.g[0] = a.g1; p.g[1] = a.g2;
  p.stats = a.stats; p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.act = a.act; p.alpha = a.act_alpha;
  p.sums = a.sums; p.sums_part = a.sums_part; p.sums_nblk = a.sums_nblk; p.dst = a.dst; p.dmap = a.dmap;
  p.gather_dst = a.gather_dst;
}
int launch_in_bwd_reduce(const InBwdParams& a, cudaStream_t st) {
  RowStreamParams p = {};
  bwd_params(a, p);
  if (a.sums_part == nullptr) return -int(cudaErrorInvalidValue);
  return launch_row_stream<RS_BWD_REDUCE>(p, st);  // blocks per image = partial sums per image
}
int launch_in_bwd_apply(const InBwdParams& a, cudaStream_t st) {
  RowStreamParams p = {};
  bwd_params(a, p);
  if (a.sums_part == nullptr || a.sums_nblk < 1) return -int(cudaErrorInvalidValue);
  const int r = launch_row_stream<RS_BWD_APPLY>(p, st);
  return r < 0 ? r : 0;
}
// One launch for both passes (row_fused_bwd_kernel).  a.sync_ctr: one int per image, zero before the launch.  The
// blocks of an image wait for each other, so every block must be resident: one block per SM and grid <= 148 (checked).
int launch_in_bwd_fused(const InBwdParams& a, cudaStream_t st) {
  RowStreamParams p = {};
  bwd_params(a, p);
  if (a.sums_part == nullptr || a.sync_ctr == nullptr) return -int(cudaErrorInvalidValue);
  p.sync_ctr = a.sync_ctr;
  dim3 grid;
  size_t smem;
  const int r = plan_row_stream(p, &grid, &smem);
  if (r != 0) return r < 0 ? r : 0;
  if (p.ns < 2 || int(grid.x * grid.y) > 148) return -int(cudaErrorInvalidConfiguration);
  if (p.ns == 2) return launch_row_fused_ns<2>(p, grid, smem, st);
  return launch_row_fused_ns<3>(p, grid, smem, st);
}
size_t in_bwd_partials_bytes(int C) { return size_t(160) * C * 2 * sizeof(float); }  // <= 148 blocks per launch

int launch_grad_gather(const GradSrc& g1, const GradSrc& g2, int B, int H, int W, int C, sg_bf16* out, cudaStream_t st) {
  RowStreamParams p = {};
  p.B = B; p.H = H; p.W = W; p.C = C; p.nb_act = B; p.act_wrap = 0;
  p.s[0] = g1.ptr ? grad_stream(g1, C) : null_stream();
  p.s[1] = g2.ptr ? grad_stream(g2, C) : null_stream();
  p.s[2] = null_stream();
  p.g[0] = g1; p.g[1] = g2;
  p.dst = out;
  const int r = launch_row_stream<RS_GATHER>(p, st);
  return r < 0 ? r : 0;
}

}  // namespace sggan
