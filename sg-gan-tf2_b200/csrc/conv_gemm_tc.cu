// conv_gemm_tc.cu -- the tensor-core kernels of the SG-GAN step, hand-written for sm_100a:
//
//   conv_gemm_tc_kernel    implicit-GEMM convolution (forward conv, dgrad, deconv phases) over
//                          pitch-linearised frames: TMA boxes -> smem (SWIZZLE_128B) ->
//                          tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> staged epilogue (bias, act,
//                          bf16 tile in smem, bulk async stores, per-(image,channel) sum / sum^2
//                          partials for instance norm).  Used for Cout > 256, fp32 outputs and, with
//                          kind::tf32 on fp32 frames, for the fp32-accurate operator tier.
//   conv_gemm_pair_kernel  the 256-channel layers (the residual blocks: 80 % of the FLOPs) on CTA pairs,
//                          tcgen05.mma.cta_group::2 with M = 256 over the pair, each CTA staging half of every
//                          weight tile; persistent, TMEM double-buffered, staged epilogue.  MMAs retire at the
//                          hardware rate (128 cycles per 128x256x16).  <true>: the dgrad form that also
//                          accumulates the norm-backward sums of the layer below in its epilogue.
//   conv_gemm_swap_kernel  the same operation with the operand roles exchanged (weights = M, 256 pixels
//                          = N), persistent with two TMEM accumulators: layers with Cout <= 128 and
//                          the 7x7 output convolution in shift-sum form.
//   wgrad_gemm_pair_kernel weight gradient of the 256 x 256-channel layers on CTA pairs, two taps per CTA
//                          sharing the dY half tile.
//   wgrad_gemm_tc_kernel   the other weight gradients: MN-major operands (pixels are K), split-K over
//                          images x pixel chunks into fp32 partial tiles, added in a fixed order by
//                          wgrad_reduce_kernel.
//
// Every tcgen05 / TMA issue loop runs under elect_one_sync() (tc_common.cuh): under `if (lane == 0)` ptxas wraps each
// UTCHMMA in an ELECT / BRA.U.ANY loop, which costs 22 % of the issue rate (156 instead of 128 cycles per MMA).
//
// Replaces the cuDNN/Eigen calls behind tf.keras.layers.Conv2D / Conv2DTranspose and their
// gradients on the reference path (module.py:211-216,232-265,284-311; model.py:196-197).
//
// The convolution is operand-bandwidth bound unless re-use is organised on chip (measured:
// 392 TFLOP/s with one TMA box per tap).  Two forms of re-use are built in:
//   * the CTA tile is MT = 256 output positions x BN channels held as two 128-row accumulators in
//     TMEM, so each weight tile (B) loaded into smem feeds two MMAs;
//   * taps whose pixel offsets are consecutive (the kw taps of one filter row) form a "run": the
//     A tile of a run is loaded ONCE as MT + 8 rows, and tap q reads it through a shared-memory
//     descriptor whose start address is shifted by q rows (q * 128 B).  SWIZZLE_128B is a function
//     of the absolute smem address, so a row-shifted descriptor (base_offset 0) addresses the same
//     swizzled data -- verified on hardware by tests/gpu/tc_probe.cu `shift`.
// A and B therefore travel in separate mbarrier rings fed by two producer warps.
//
// Warp roles of conv_gemm_tc_kernel (480 threads): warp 0 = A producer, warp 1 = TMEM allocator + MMA
// issuer, warp 2 = B producer, warps 3..10 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31; two warps
// per quarter), warps 11..14 = statistics + store warps of the staged epilogue.  All kernels here are
// launched with programmatic stream serialization: their set-up overlaps the predecessor's tail and
// pdl_wait() precedes the first global access.
#include "conv_gemm_tc.h"
#include "tc_common.cuh"
#include "tmap.h"

#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

namespace sggan {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                    // bf16 elements per 128-byte swizzled row
constexpr int kABytes = kTileM * kChunkK * 2;  // 16 KB: one 128-row A box
constexpr int kHaloRows = 8;                   // one extra swizzle atom of rows covers runs of up to 8 taps
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;  // pipeline bytes (leaves room for alignment slack)
constexpr int kEpiWarps = 8;                           // two warps per TMEM lane quarter (16 measured no faster:
                                                       // the epilogue is bound by TMEM / smem / store traffic)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kStatThreads = 128;                      // statistics + bulk-store warps of the staged epilogue
constexpr int kConvThreads = 96 + kEpiThreads + kStatThreads;  // 3 control warps + epilogue warps + statistics warps
constexpr int kWgradThreads = 192;

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == SG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == SG_ACT_LRELU) return v > 0.f ? v : alpha * v;
  if (act == SG_ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

static inline uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// Column sums (sum, sum of squares) of a staged bf16 tile S[128][BN] (row pitch `pitch` bytes) by `nthr`
// threads (tid in [0, nthr)); each thread owns a column pair over 128 / parts rows, the parts are added in
// a fixed order through `comb` ([parts][BN], synchronised on named barrier `bar_id`).  gst -> [BN] float2.
__device__ __forceinline__ void staged_stats(const uint8_t* S, int pitch, int BN, int nthr, int tid, float2* comb,
                                             int bar_id, float2* gst) {
  const int pw = pitch >> 2, cpairs = BN >> 1;
  const int parts = nthr / cpairs, rows_per = kTileM / parts;
  const int cp = tid % cpairs, part = tid / cpairs;
  const uint32_t* col = reinterpret_cast<const uint32_t*>(S + part * rows_per * pitch) + cp;
  float s1l = 0.f, s2l = 0.f, s1h = 0.f, s2h = 0.f, t1l = 0.f, t2l = 0.f, t1h = 0.f, t2h = 0.f;
#pragma unroll 8
  for (int r = 0; r < rows_per; r += 2) {
    const uint32_t w0 = col[r * pw], w1 = col[(r + 1) * pw];
    const float x0 = __uint_as_float(w0 << 16), y0 = __uint_as_float(w0 & 0xffff0000u);
    const float x1 = __uint_as_float(w1 << 16), y1 = __uint_as_float(w1 & 0xffff0000u);
    s1l += x0; s2l += x0 * x0; s1h += y0; s2h += y0 * y0;
    t1l += x1; t2l += x1 * x1; t1h += y1; t2h += y1 * y1;
  }
  if (parts == 1) {
    gst[2 * cp] = make_float2(s1l + t1l, s2l + t2l);
    gst[2 * cp + 1] = make_float2(s1h + t1h, s2h + t2h);
  } else {
    comb[part * BN + 2 * cp] = make_float2(s1l + t1l, s2l + t2l);
    comb[part * BN + 2 * cp + 1] = make_float2(s1h + t1h, s2h + t2h);
    named_bar_sync(bar_id, nthr);
    if (tid < BN) {
      float2 acc = comb[tid];
      for (int pp = 1; pp < parts; ++pp) {
        const float2 o = comb[pp * BN + tid];
        acc.x += o.x; acc.y += o.y;
      }
      gst[tid] = acc;
    }
  }
}

// =============================================================================================
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA8,
                    const __grid_constant__ CUtensorMap tmB, const ConvGemmParams p, const int sa_stages,
                    const int sb_stages, const uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space
  __shared__ uint64_t a_full[kMaxStages], a_empty[kMaxStages], b_full[kMaxStages], b_empty[kMaxStages];
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) float sbias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + (int64_t(blockIdx.z) * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const int BN = p.BN, MT = p.MT, NA = MT / kTileM;
  const int a_stage_bytes = (MT + kHaloRows) * 128, b_stage_bytes = BN * 128;
  uint8_t* smA = smem;
  uint8_t* smB = smem + sa_stages * a_stage_bytes;
  const int mstep = p.shift_kw > 0 ? MT - (p.shift_kw - 1) : MT;
  const int m0 = blockIdx.x * mstep, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int ck = p.tf32 ? kChunkK / 2 : kChunkK;  // elements per 128-byte swizzled row: 64 bf16 or 32 fp32 (tf32)
  const int cchunks = p.Cin / ck;
  const int agroups = p.nruns * cchunks;  // A tiles this CTA consumes
  // staged epilogue (see below): bf16 output and a full tile of channels
  const bool staged = p.shift_kw == 0 && !p.out_f32 && (p.omap.C & 7) == 0 && n0 + BN <= p.Cout;

  if (threadIdx.x == 0) {
    for (int s = 0; s < sa_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < sb_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA8);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&tmem_base_sh, tmem_cols);
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  for (int t = threadIdx.x; t < 256; t += kConvThreads)
    sbias[t] = (p.bias != nullptr && t < BN && n0 + t < p.Cout) ? __ldg(p.bias + n0 + t) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ A producer: one tile per (run, chunk)
    if (elect_one_sync()) {
      for (int g = 0; g < agroups; ++g) {
        const int s = g % sa_stages;
        const uint32_t ph = (g / sa_stages) & 1;
        mbar_wait(&a_empty[s], ph ^ 1, 1);
        const int r = g / cchunks, cc = g - r * cchunks;
        uint8_t* sa = smA + s * a_stage_bytes;
        const int row0 = m0 + p.run_off[r];
        mbar_arrive_expect_tx(&a_full[s], a_stage_bytes);
        for (int a = 0; a < NA; ++a) tma_load_3d(&tmA, &a_full[s], sa + a * kABytes, cc * ck, row0 + a * kTileM, b);
        tma_load_3d(&tmA8, &a_full[s], sa + NA * kABytes, cc * ck, row0 + MT, b);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ B producer: one weight tile per (tap, chunk)
    if (elect_one_sync()) {
      int it = 0;
      for (int r = 0, t0 = 0; r < p.nruns; t0 += p.run_len[r], ++r)
        for (int cc = 0; cc < cchunks; ++cc)
          for (int q = 0; q < p.run_len[r]; ++q, ++it) {
            const int s = it % sb_stages;
            const uint32_t ph = (it / sb_stages) & 1;
            mbar_wait(&b_empty[s], ph ^ 1, 4);
            mbar_arrive_expect_tx(&b_full[s], b_stage_bytes);
            tma_load_2d(&tmB, &b_full[s], smB + s * b_stage_bytes, cc * ck, int(p.run_w[t0 + q]) * p.CoutPad + n0);
          }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const bool tf32 = p.tf32 != 0;
      const uint32_t idesc = tf32 ? idesc_tf32_f32(kTileM, BN) : idesc_bf16_f32(kTileM, BN, 0, 0);
      int it = 0, g = 0;
      for (int r = 0; r < p.nruns; ++r)
        for (int cc = 0; cc < cchunks; ++cc, ++g) {
          const int sa_i = g % sa_stages;
          mbar_wait(&a_full[sa_i], (g / sa_stages) & 1, 2);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smA + sa_i * a_stage_bytes);
          for (int q = 0; q < p.run_len[r]; ++q, ++it) {
            const int sb_i = it % sb_stages;
            mbar_wait(&b_full[sb_i], (it / sb_stages) & 1, 5);
            tc_fence_after();
            const uint64_t bdesc = desc_kmajor_sw128(smem_u32(smB + sb_i * b_stage_bytes));
            for (int a = 0; a < NA; ++a) {
              // tap q of the run: the A view starts q rows (q * 128 B) into the tile
              const uint64_t adesc = desc_kmajor_sw128(a_base + uint32_t(a * kTileM + q) * 128u);
              if (tf32) {
#pragma unroll
                for (int k = 0; k < 4; ++k)  // UMMA_K = 8 tf32 = 32 bytes inside the swizzle row
                  umma_tf32(tmem_acc + uint32_t(a * BN), adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (it | k) != 0);
              } else {
#pragma unroll
                for (int k = 0; k < kChunkK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes inside the swizzle row
                  umma_bf16(tmem_acc + uint32_t(a * BN), adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc,
                            (it | k) != 0);
              }
            }
            umma_commit(&b_empty[sb_i]);  // frees the weight slot once these MMAs retire
          }
          umma_commit(&a_empty[sa_i]);
        }
      umma_commit(&acc_bar);  // accumulators complete
      if (dbg) dbg[2] = clock64();
    }
  } else if (warp >= 3 + kEpiWarps) {
    // ------------------------------------------------------------ statistics + store warps (staged epilogue)
    // While the epilogue warps drain the NEXT accumulator out of TMEM, these four warps finish the tile
    // that was just staged in shared memory: one bulk async copy per valid output position, and the
    // column sums for the instance-norm statistics (the epilogue warps finish the last tile themselves).
    if (staged) {
      const int t = threadIdx.x - (96 + kEpiThreads);  // 0 .. 127
      const int pitch = BN * 2 + 16;
      uint8_t* aux = smem + NA * kTileM * pitch;
      float2* comb0 = reinterpret_cast<float2*>(aux);                     // 2 KB: partial sums of these warps
      const long long* rowoff = reinterpret_cast<const long long*>(aux + 6144);  // [NA][128] output offsets
      // all but the last accumulator: the epilogue warps finish the last one themselves
      for (int a = 0; a + 1 < NA; ++a) {
        named_bar_sync(2 + a, kEpiThreads + kStatThreads);
        const uint8_t* S = smem + a * kTileM * pitch;
        {
          const long long off = rowoff[a * kTileM + t];
          if (off >= 0) {
            bulk_store_1d(reinterpret_cast<__nv_bfloat16*>(p.out) + off, S + t * pitch, uint32_t(BN * 2));
            bulk_commit_group();
          }
        }
        if (p.stats != nullptr) {
          const int tile = p.stats_t0 + blockIdx.x * NA + a;
          staged_stats(S, pitch, BN, kStatThreads, t, comb0, 4,
                       reinterpret_cast<float2*>(p.stats) + (int64_t(b) * p.stats_T + tile) * p.Cout + n0);
        }
      }
      bulk_wait_group_read0();  // shared memory must outlive the reads of the copies
      if (dbg && t == 0) dbg[7] = clock64();
    }
  } else {
    // ------------------------------------------------------------ epilogue (kEpiWarps warps)
    // The warps that share a TMEM lane quarter split the 32-column chunks between them, so every
    // scheduler has several epilogue warps to interleave.  The per-element work is kept branch-free: bias
    // comes from smem (zero padded), padded weight rows make columns >= Cout exactly zero, the
    // activation is chosen once per chunk.
    mbar_wait(&acc_bar, 0, 3);
    tc_fence_after();
    const int q = warp & 3;
    const int half = (warp - 3) >> 2;  // 0 .. kEpiWarps/4 - 1: which share of the 32-column chunks
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 96;  // 0 .. kEpiThreads - 1
    if (dbg && et == 0) dbg[3] = clock64();
    float* tsm = reinterpret_cast<float*>(smem);  // [BN][129] transposed fp32 tile for the statistics
    const bool has_stats = p.stats != nullptr;
    const int act = p.act;
    const float alpha = p.act_alpha;
    if (p.shift_kw > 0) {
      // ---- shift-sum epilogue: stage the (kw, co) partial products of all MT rows, then add across rows
      float* S = reinterpret_cast<float*>(smem);  // [MT][33]
      if (half < NA) {
        float v[32];
        tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(half * BN), v);
        float* srow = S + (half * kTileM + row) * 33;
#pragma unroll
        for (int e = 0; e < 32; ++e) srow[e] = v[e];
      }
      named_bar_sync(1, kEpiThreads);
      const int r = et;
      if (r < mstep) {
        const int m = m0 + r;
        const int i = m / p.P, j = m - i * p.P;
        if (m < p.M && i < p.Hv && j < p.Wv) {
          const int64_t ob = (int64_t(b) * p.omap.frame_pix + frame_pixel(p.omap, i, j)) * p.omap.C;
          for (int co = 0; co < p.Cout; ++co) {
            float x = sbias[co];
            for (int kw = 0; kw < p.shift_kw; ++kw) x += S[(r + kw) * 33 + kw * 4 + co];
            x = apply_act(x, act, alpha);
            if (p.out_f32) reinterpret_cast<float*>(p.out)[ob + co] = x;
            else reinterpret_cast<__nv_bfloat16*>(p.out)[ob + co] = __float2bfloat16_rn(x);
          }
        }
      }
    } else if (staged) {
      // ---- staged epilogue (bf16 output, full channel tile): the tile is written row-major into shared
      // memory (the pipeline buffers are idle by now).  Every valid output position then leaves as ONE
      // bulk async copy of BN * 2 bytes, and the statistics are column sums of the staged bf16 values.  All
      // but the last accumulator are handed to the statistics warps, so that work overlaps the TMEM drain
      // of the next accumulator.  No strided global stores, no fp32 transposition tile.
      const int pitch = BN * 2 + 16;  // bytes; +16 keeps the 16-byte row writes of a quarter warp on distinct banks
      uint8_t* aux = smem + NA * kTileM * pitch;  // 2 KB + 4 KB partial sums, 2 KB output offsets
      long long* rowoff = reinterpret_cast<long long*>(aux + 6144);
      for (int a = 0; a < NA; ++a) {
        uint8_t* S = smem + a * kTileM * pitch;
        const int m = m0 + a * kTileM + row;
        const int i = m / p.P, j = m - i * p.P;
        const int oi = i * p.o_scale + p.o_a, oj = j * p.o_scale + p.o_b;
        const bool valid = (m < p.M) && (i < p.Hv) && (j < p.Wv) && (oi < p.omap.H) && (oj < p.omap.W);
        for (int c0 = half * 32; c0 < BN; c0 += 32 * (kEpiWarps / 4)) {
          float v[32];
          tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(a * BN + c0), v);
          const float4* sb4 = reinterpret_cast<const float4*>(sbias + c0);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bb = sb4[g];
            v[4 * g] += bb.x; v[4 * g + 1] += bb.y; v[4 * g + 2] += bb.z; v[4 * g + 3] += bb.w;
          }
          if (act == SG_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
          } else if (act == SG_ACT_LRELU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], alpha * v[e]);
          } else if (act == SG_ACT_TANH) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = tanhf(v[e]);
          }
          uint4* srow = reinterpret_cast<uint4*>(S + row * pitch + c0 * 2);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = valid ? pack_bf16x2(v[8 * g], v[8 * g + 1]) : 0u;  // rows outside the image count as zeros
            w.y = valid ? pack_bf16x2(v[8 * g + 2], v[8 * g + 3]) : 0u;
            w.z = valid ? pack_bf16x2(v[8 * g + 4], v[8 * g + 5]) : 0u;
            w.w = valid ? pack_bf16x2(v[8 * g + 6], v[8 * g + 7]) : 0u;
            srow[g] = w;
          }
        }
        if (half == 0) {
          const int64_t obase = (int64_t(b) * p.omap.frame_pix + (valid ? frame_pixel(p.omap, oi, oj) : 0)) * p.omap.C;
          rowoff[a * kTileM + row] = valid ? obase + n0 : -1;
        }
        fence_proxy_async_smem();  // bulk copies (async proxy) read what this thread just wrote
        if (a + 1 < NA) named_bar_arrive(2 + a, kEpiThreads + kStatThreads);  // hand over; go on with the next accumulator
        if (dbg && et == 0) dbg[4 + a] = clock64();
      }
      named_bar_sync(1, kEpiThreads);  // the last tile is completely staged
      if (half == 0) {
        const long long off = rowoff[(NA - 1) * kTileM + row];
        if (off >= 0) {
          bulk_store_1d(reinterpret_cast<__nv_bfloat16*>(p.out) + off, smem + ((NA - 1) * kTileM + row) * pitch,
                        uint32_t(BN * 2));
          bulk_commit_group();
        }
      }
      if (has_stats) {
        const int tile = p.stats_t0 + blockIdx.x * NA + (NA - 1);
        staged_stats(smem + (NA - 1) * kTileM * pitch, pitch, BN, kEpiThreads, et,
                     reinterpret_cast<float2*>(aux + 2048), 1,
                     reinterpret_cast<float2*>(p.stats) + (int64_t(b) * p.stats_T + tile) * p.Cout + n0);
      }
      bulk_wait_group_read0();
      tc_fence_before();
    } else
    for (int a = 0; a < NA; ++a) {
      const int m = m0 + a * kTileM + row;
      const int i = m / p.P, j = m - i * p.P;
      const int oi = i * p.o_scale + p.o_a, oj = j * p.o_scale + p.o_b;
      const bool valid = (m < p.M) && (i < p.Hv) && (j < p.Wv) && (oi < p.omap.H) && (oj < p.omap.W);
      const int64_t obase = (int64_t(b) * p.omap.frame_pix + (valid ? frame_pixel(p.omap, oi, oj) : 0)) * p.omap.C;
      const bool vec_ok = (!p.out_f32) && ((p.omap.C & 7) == 0);
      const float msk = valid ? 1.f : 0.f;
      for (int c0 = half * 32; c0 < BN; c0 += 32 * (kEpiWarps / 4)) {
        float v[32];
        tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(a * BN + c0), v);
        const int nb = n0 + c0;
        const float4* sb4 = reinterpret_cast<const float4*>(sbias + c0);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bb = sb4[g];
          v[4 * g] += bb.x; v[4 * g + 1] += bb.y; v[4 * g + 2] += bb.z; v[4 * g + 3] += bb.w;
        }
        if (act == SG_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
        } else if (act == SG_ACT_LRELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], alpha * v[e]);  // 0 < alpha < 1
        } else if (act == SG_ACT_TANH) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = tanhf(v[e]);
        }
        uint32_t pk[16];
        if (!p.out_f32) {
          // statistics are taken over the values as stored (bf16), so that the normalisation that
          // follows is exact for what it reads (H*W == 1 must give exactly beta)
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2) {
            pk[e2] = pack_bf16x2(v[2 * e2], v[2 * e2 + 1]);
            v[2 * e2] = __uint_as_float(pk[e2] << 16);
            v[2 * e2 + 1] = __uint_as_float(pk[e2] & 0xffff0000u);
          }
        }
        if (has_stats) {
#pragma unroll
          for (int e = 0; e < 32; ++e) tsm[(c0 + e) * 129 + row] = v[e] * msk;
        }
        if (valid) {
          if (vec_ok && nb + 32 <= p.Cout) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + obase + nb);
#pragma unroll
            for (int g = 0; g < 4; ++g) dst[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          } else if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.out) + obase;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (nb + e < p.Cout) dst[nb + e] = v[e];
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + obase;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (nb + e < p.Cout) dst[nb + e] = __float2bfloat16_rn(v[e]);
          }
        }
      }
      if (has_stats) {
        named_bar_sync(1, kEpiThreads);
        for (int c = et; c < BN; c += kEpiThreads) {
          const int n = n0 + c;
          if (n >= p.Cout) continue;
          const float* col = tsm + c * 129;
          float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll 8
          for (int r = 0; r < 128; r += 2) {
            const float x = col[r], y = col[r + 1];
            s1 += x; s2 += x * x;
            s1b += y; s2b += y * y;
          }
          const int tile = p.stats_t0 + blockIdx.x * NA + a;
          float2* dst = reinterpret_cast<float2*>(p.stats) + (int64_t(b) * p.stats_T + tile) * p.Cout + n;
          *dst = make_float2(s1 + s1b, s2 + s2b);
        }
        if (a + 1 < NA) named_bar_sync(1, kEpiThreads);  // tsm is rewritten by the next accumulator
      }
      if (dbg && et == 0) dbg[4 + a] = clock64();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
  if (dbg && threadIdx.x == 0) dbg[6] = clock64();
}

// =============================================================================================
// CTA-pair variant for the 256-channel layers (the residual blocks: 80 % of the step's FLOPs).
//
// A cluster of two CTAs runs tcgen05.mma.cta_group::2 with M = 256 (128 output positions per CTA, any
// two 128-row tiles of the batch), N = 256: each CTA stages only HALF of every weight tile (its 128 of
// the 256 output channels), so the weight stream per SM -- what bounds the single-CTA kernel in L2 --
// halves without needing a second accumulator.  That frees TMEM for DOUBLE BUFFERING (2 x 256 columns):
// the kernel is persistent, and the epilogue of tile k (TMEM -> bias/act/statistics -> bf16 stores,
// ~9k cycles) runs while the MMAs of tile k + 1 are issued.  The single-CTA kernel spends ~29 % of its
// time in that epilogue with the tensor pipe idle.
//
// Barriers: a_full / b_full live in the leader CTA (rank 0) and count the TMA bytes of BOTH CTAs;
// a_empty / b_empty / acc_full exist in both CTAs and are signalled by multicast tcgen05.commit;
// acc_empty lives in the leader and collects one arrival per epilogue warp of both CTAs.
constexpr int kPairEpiTile = 32 * 33 * 4;                       // one warp's transposition tile (fallback epilogue)
// staged epilogue: the 128 x 256 bf16 tile row-major (pitch + 16 B: conflict-free 16-byte row writes), then 4 KB of partial
// column sums and 1 KB of output offsets; the fallback epilogue (fp32 output / partial channel tile) fits inside
constexpr int kPairStagePitch = 256 * 2 + 16;
constexpr int kPairEpiBytes = kTileM * kPairStagePitch + 4096 + 1024;
static_assert(kPairEpiBytes >= kEpiWarps * kPairEpiTile + 2 * 4 * 256 * 8, "fallback epilogue scratch");
constexpr int kPairAStages = 3, kPairBStages = 6;  // (maximum; a launch that also folds the norm-backward sums runs 3 + 4)
constexpr int kPairAStage = (kTileM + kHaloRows) * 128, kPairBStage = 128 * 128;
constexpr int kPairSmem = kPairAStages * kPairAStage + kPairBStages * kPairBStage + kPairEpiBytes + 1024;
// norm-backward fold: the Y rows matching a slab of the tile (32 positions x 512 B, one per lane of the staging warp),
// double buffered
constexpr int kPairYRows = 32, kPairYSlab = kPairYRows * 512, kPairYBufs = 2, kPairYSlabs = kTileM / kPairYRows;
constexpr int kPairFoldAStages = 3, kPairFoldBStages = 4;
constexpr int kPairFoldSmem = kPairFoldAStages * kPairAStage + kPairFoldBStages * kPairBStage + kPairEpiBytes + kPairYBufs * kPairYSlab + 1024;
static_assert(kPairFoldSmem <= kPairSmem, "the fold variant must fit the same shared-memory opt-in");

// FOLD: the launch also accumulates the norm-backward sums of the layer below (ConvGemmParams::nr_*); it then runs 3 + 4
// pipeline stages instead of 3 + 6 to make room for the Y slab buffers (2 + 5 measured 4 us slower per launch).  Compile-time, because the ring indices
// (g % stages, g / stages) sit in the single-thread MMA issue loop: as run-time divisions they cost 15 % of the kernel.
template <bool FOLD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA8,
                      const __grid_constant__ CUtensorMap tmBh, const ConvGemmParams p, const int T128,
                      const int npairs) {
  constexpr int a_stages = FOLD ? kPairFoldAStages : kPairAStages, b_stages = FOLD ? kPairFoldBStages : kPairBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space
  __shared__ uint64_t a_full[kPairAStages], a_empty[kPairAStages], b_full[kPairBStages], b_empty[kPairBStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint64_t y_full[kPairYBufs], y_empty[kPairYBufs];  // norm-backward fold: Y slab buffers
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) float sbias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  uint8_t* smA = smem;
  uint8_t* smB = smem + a_stages * kPairAStage;
  float* epi = reinterpret_cast<float*>(smB + b_stages * kPairBStage);
  uint8_t* ybuf = reinterpret_cast<uint8_t*>(epi) + kPairEpiBytes;  // only carved when the launch folds (nr_Y != null)
  constexpr bool fold = FOLD;
  const int cchunks = p.Cin / kChunkK;
  const int total_tiles = p.B * T128;
  // debug stamps, 16 per CTA: [0] start, [1] set-up done, [2+k] MMAs of tile k issued (leader),
  // [8+k] epilogue of tile k done, [15] end
  long long* dbg = p.dbg ? p.dbg + int64_t(blockIdx.x) * 16 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kPairBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * kEpiWarps); }
    for (int s = 0; s < kPairYBufs; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], kEpiWarps); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA8);
    tma_prefetch_desc(&tmBh);
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_sh, 512);
  pdl_launch_dependents();
  pdl_wait();  // the set-up above overlapped the previous kernel's tail; global memory is touched only below
  for (int t = threadIdx.x; t < 256; t += kConvThreads)
    sbias[t] = (p.bias != nullptr && t < p.Cout) ? __ldg(p.bias + t) : 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ A producer (both CTAs: own 128 rows)
    if (elect_one_sync()) {
      int g = 0;
      for (int j = cl; j < npairs; j += ncl) {
        const int gt = 2 * j + int(rank);
        const int gtc = gt < total_tiles ? gt : 0;
        const int b = gtc / T128, m0 = (gtc - b * T128) * kTileM;
        for (int r = 0; r < p.nruns; ++r)
          for (int cc = 0; cc < cchunks; ++cc, ++g) {
            const int s = g % a_stages;
            mbar_wait(&a_empty[s], ((g / a_stages) & 1) ^ 1, 1);
            uint8_t* sa = smA + s * kPairAStage;
            const int row0 = m0 + p.run_off[r];
            if (rank == 0) mbar_arrive_expect_tx(&a_full[s], 2 * kPairAStage);
            const uint32_t bar = mapa_u32(smem_u32(&a_full[s]), 0);
            tma_load_3d_pair(&tmA, bar, sa, cc * kChunkK, row0, b);
            tma_load_3d_pair(&tmA8, bar, sa + kABytes, cc * kChunkK, row0 + kTileM, b);
          }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ B producer (both CTAs: own 128 channels)
    if (elect_one_sync()) {
      int it = 0;
      for (int j = cl; j < npairs; j += ncl)
        for (int r = 0, t0 = 0; r < p.nruns; t0 += p.run_len[r], ++r)
          for (int cc = 0; cc < cchunks; ++cc)
            for (int q = 0; q < p.run_len[r]; ++q, ++it) {
              const int s = it % b_stages;
              mbar_wait(&b_empty[s], ((it / b_stages) & 1) ^ 1, 4);
              if (rank == 0) mbar_arrive_expect_tx(&b_full[s], 2 * kPairBStage);
              tma_load_2d_pair(&tmBh, mapa_u32(smem_u32(&b_full[s]), 0), smB + s * kPairBStage, cc * kChunkK,
                               int(p.run_w[t0 + q]) * p.CoutPad + int(rank) * 128);
            }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one_sync()) {  // rank is CTA-uniform: the warp is converged at the election
      const uint32_t idesc = idesc_bf16_f32(256, 256, 0, 0);
      int it = 0, g = 0, k = 0;
      for (int j = cl; j < npairs; j += ncl, ++k) {
        const int buf = k & 1;
        mbar_wait(&acc_empty[buf], ((k >> 1) & 1) ^ 1, 6);  // both CTAs drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_acc + uint32_t(buf * 256);
        uint32_t acc = 0;
        for (int r = 0; r < p.nruns; ++r)
          for (int cc = 0; cc < cchunks; ++cc, ++g) {
            const int sa_i = g % a_stages;
            mbar_wait(&a_full[sa_i], (g / a_stages) & 1, 2);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smA + sa_i * kPairAStage);
            for (int q = 0; q < p.run_len[r]; ++q, ++it) {
              const int sb_i = it % b_stages;
              mbar_wait(&b_full[sb_i], (it / b_stages) & 1, 5);
              tc_fence_after();
              const uint64_t bdesc = desc_kmajor_sw128(smem_u32(smB + sb_i * kPairBStage));
              const uint64_t adesc = desc_kmajor_sw128(a_base + uint32_t(q) * 128u);
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) {
                umma_bf16_pair(d, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc, acc);
                acc = 1;
              }
              umma_commit_pair(&b_empty[sb_i]);
            }
            umma_commit_pair(&a_empty[sa_i]);
          }
        umma_commit_pair(&acc_full[buf]);
        if (dbg && k < 6) dbg[2 + k] = clock64();
      }
    }
  } else if (warp == 3 + kEpiWarps) {
    // ------------------------------------------------------------ norm-backward fold: Y rows of the tile, a slab at a time
    // Position m of the padded grid belongs to source pixel (reflect(i - pad), reflect(j - pad)).  Lane r of this warp owns
    // position r of the slab; runs of positions whose source rows are consecutive in memory (the interior of a grid row;
    // reflected border positions stand alone) are found with two ballots, and only the FIRST lane of a run issues a bulk
    // copy, for the whole run: 1-4 copies per slab instead of 32 (per-lane copies serialise in the issue loop: measured
    // ~3k cycles per slab, which made the kernel epilogue-bound).
    if (fold) {
      int qn = 0;
      for (int j = cl; j < npairs; j += ncl) {
        const int gt = 2 * j + int(rank);
        const bool tile_ok = gt < total_tiles;
        const int gtc = tile_ok ? gt : 0;
        const int b = gtc / T128, t128 = gtc - b * T128;
        const sg_bf16* yimg = p.nr_Y + int64_t(b) * p.nr_H * p.nr_W * 256;
        for (int qq = 0; qq < kPairYSlabs; ++qq, ++qn) {
          const int s = qn % kPairYBufs;
          if (lane == 0) mbar_wait(&y_empty[s], ((qn / kPairYBufs) & 1) ^ 1, 7);
          __syncwarp();
          const int m = t128 * kTileM + qq * kPairYRows + lane;
          const int i = m / p.P, jj = m - i * p.P;
          const bool valid = tile_ok && (m < p.M) && (i < p.Hv) && (jj < p.Wv);
          int src = -1;
          if (valid) {
            int ri = i - p.nr_pad, rj = jj - p.nr_pad;
            ri = ri < 0 ? -ri : (ri >= p.nr_H ? 2 * (p.nr_H - 1) - ri : ri);
            rj = rj < 0 ? -rj : (rj >= p.nr_W ? 2 * (p.nr_W - 1) - rj : rj);
            src = ri * p.nr_W + rj;
          }
          const int prev = __shfl_up_sync(0xffffffffu, src, 1);
          const bool start = valid && (lane == 0 || prev < 0 || src != prev + 1);
          const unsigned vm = __ballot_sync(0xffffffffu, valid), sm = __ballot_sync(0xffffffffu, start);
          if (lane == 0) mbar_arrive_expect_tx(&y_full[s], uint32_t(__popc(vm)) * 512u);
          __syncwarp();
          if (start) {
            const unsigned after = lane == 31 ? 0u : ((sm | ~vm) >> (lane + 1));  // later positions that end this run
            const int len = after ? __ffs(after) : 32 - lane;
            bulk_load_1d(ybuf + s * kPairYSlab + lane * 512, yimg + int64_t(src) * 256, uint32_t(len) * 512u, &y_full[s]);
          }
        }
      }
    }
  } else if (warp < 3 + kEpiWarps) {
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 x 256 tile)
    const int q = warp & 3;
    const int half = (warp - 3) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 96;
    float* tsw = epi + (warp - 3) * (32 * 33);
    float2* red = reinterpret_cast<float2*>(epi + kEpiWarps * 32 * 33);  // [2][4][256]
    const bool has_stats = p.stats != nullptr;
    const int act = p.act;
    const float alpha = p.act_alpha;
    const bool vec_ok = (!p.out_f32) && ((p.omap.C & 7) == 0);
    const uint32_t empty_addr0 = mapa_u32(smem_u32(&acc_empty[0]), 0), empty_addr1 = mapa_u32(smem_u32(&acc_empty[1]), 0);
    // staged epilogue (bf16 output, all 256 channels): the tile goes row-major into shared memory, every valid output
    // position then leaves as ONE 512-byte bulk async copy and the statistics are column sums of the staged bf16 values
    // (as in the single-CTA kernel).  ~5k cycles per tile instead of ~12k with per-thread global stores and the fp32
    // transposition tiles: what matters is the LAST tile of a CTA, whose epilogue no MMA hides.
    const bool staged = !p.out_f32 && (p.omap.C & 7) == 0 && p.Cout == 256;
    uint8_t* S = reinterpret_cast<uint8_t*>(epi);
    float2* comb = reinterpret_cast<float2*>(S + kTileM * kPairStagePitch);               // [2][256]
    long long* rowoff = reinterpret_cast<long long*>(S + kTileM * kPairStagePitch + 4096);  // [128]
    int k = 0, yq = 0;
    for (int j = cl; j < npairs; j += ncl, ++k) {
      const int buf = k & 1;
      const int gt = 2 * j + int(rank);
      const bool tile_ok = gt < total_tiles;
      const int gtc = tile_ok ? gt : 0;
      const int b = gtc / T128, t128 = gtc - b * T128;
      const int m = t128 * kTileM + row;
      const int i = m / p.P, jj = m - i * p.P;
      const int oi = i * p.o_scale + p.o_a, oj = jj * p.o_scale + p.o_b;
      const bool valid = tile_ok && (m < p.M) && (i < p.Hv) && (jj < p.Wv) && (oi < p.omap.H) && (oj < p.omap.W);
      const int64_t obase = (int64_t(b) * p.omap.frame_pix + (valid ? frame_pixel(p.omap, oi, oj) : 0)) * p.omap.C;
      const float msk = valid ? 1.f : 0.f;
      mbar_wait(&acc_full[buf], (k >> 1) & 1, 3);
      tc_fence_after();
      if (staged) {
        // the previous tile's bulk copies have finished READING the staging tile (issued by the half == 0 threads)
        bulk_wait_group_read0();
        named_bar_sync(1, kEpiThreads);
        for (int c0 = half * 32; c0 < 256; c0 += 32 * (kEpiWarps / 4)) {
          float v[32];
          tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(buf * 256 + c0), v);
          const float4* sb4 = reinterpret_cast<const float4*>(sbias + c0);
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const float4 bb = sb4[g4];
            v[4 * g4] += bb.x; v[4 * g4 + 1] += bb.y; v[4 * g4 + 2] += bb.z; v[4 * g4 + 3] += bb.w;
          }
          if (act == SG_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
          } else if (act == SG_ACT_LRELU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], alpha * v[e]);
          } else if (act == SG_ACT_TANH) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = tanhf(v[e]);
          }
          uint4* srow = reinterpret_cast<uint4*>(S + row * kPairStagePitch + c0 * 2);
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint4 w;
            w.x = valid ? pack_bf16x2(v[8 * g4], v[8 * g4 + 1]) : 0u;  // rows outside the image count as zeros
            w.y = valid ? pack_bf16x2(v[8 * g4 + 2], v[8 * g4 + 3]) : 0u;
            w.z = valid ? pack_bf16x2(v[8 * g4 + 4], v[8 * g4 + 5]) : 0u;
            w.w = valid ? pack_bf16x2(v[8 * g4 + 6], v[8 * g4 + 7]) : 0u;
            srow[g4] = w;
          }
        }
        // this warp's share of the accumulator has been read: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0);
        if (half == 0) rowoff[row] = valid ? obase : -1;
        fence_proxy_async_smem();  // the bulk copies (async proxy) read what this thread just wrote
        named_bar_sync(1, kEpiThreads);
        if (half == 0) {
          const long long off = rowoff[row];
          if (off >= 0) {
            bulk_store_1d(reinterpret_cast<__nv_bfloat16*>(p.out) + off, S + row * kPairStagePitch, 512u);
            bulk_commit_group();
          }
        }
        if (has_stats && tile_ok)  // tile_ok is uniform over the CTA
          staged_stats(S, kPairStagePitch, 256, kEpiThreads, et, comb, 2,
                       reinterpret_cast<float2*>(p.stats) + (int64_t(b) * p.stats_T + p.stats_t0 + t128) * p.Cout);
        if (fold) {
          // (sum dz, sum dz * xhat) of this tile per channel: thread (cp, hh) owns channels 2cp, 2cp + 1 and rows
          // hh * 16 .. hh * 16 + 15 of every 32-row slab; dX comes from the staged bf16 tile (the value as stored, what the
          // apply pass will read back), Y from the quarter buffers.  Same expressions as glue_rows.cu BWD_REDUCE.
          const int cp = et & 127, hh = et >> 7;
          float mu0 = 0.f, mu1 = 0.f, rs0 = 1.f, rs1 = 1.f, sc0 = 1.f, sc1 = 1.f, be0 = 0.f, be1 = 0.f;
          if (tile_ok) {
            const float n = float(p.nr_H) * float(p.nr_W);
            const float4 st = __ldg(reinterpret_cast<const float4*>(p.nr_stats) + (int64_t(b) * 256 + 2 * cp) / 2);
            const float2 ga = __ldg(reinterpret_cast<const float2*>(p.nr_gamma) + cp);
            const float2 bt = __ldg(reinterpret_cast<const float2*>(p.nr_beta) + cp);
            mu0 = st.x / n; rs0 = rsqrtf(fmaxf(st.y / n - mu0 * mu0, 0.f) + p.nr_eps);
            mu1 = st.z / n; rs1 = rsqrtf(fmaxf(st.w / n - mu1 * mu1, 0.f) + p.nr_eps);
            sc0 = ga.x * rs0; sc1 = ga.y * rs1; be0 = bt.x; be1 = bt.y;
          }
          const float gneg = p.nr_gneg;
          float a10 = 0.f, a20 = 0.f, a11 = 0.f, a21 = 0.f;
          for (int qq = 0; qq < kPairYSlabs; ++qq, ++yq) {
            const int s = yq % kPairYBufs;
            mbar_wait(&y_full[s], (yq / kPairYBufs) & 1, 8);
            const uint8_t* yb = ybuf + s * kPairYSlab;
            // branch-free over the 16 rows (independent loads first, so that they pipeline): rows outside the image carry
            // no copy -- their dX is zero in the staged tile and their Y is masked here
            const int rq0 = hh * (kPairYRows / 2), rt0 = qq * kPairYRows + rq0;
            uint32_t okm = 0;
#pragma unroll
            for (int r = 0; r < kPairYRows / 2; ++r) okm |= (rowoff[rt0 + r] >= 0 ? 1u : 0u) << r;
            uint32_t dwv[kPairYRows / 2], ywv[kPairYRows / 2];
#pragma unroll
            for (int r = 0; r < kPairYRows / 2; ++r) {
              dwv[r] = *reinterpret_cast<const uint32_t*>(S + (rt0 + r) * kPairStagePitch + cp * 4);
              ywv[r] = *reinterpret_cast<const uint32_t*>(yb + (rq0 + r) * 512 + cp * 4);
            }
#pragma unroll
            for (int r = 0; r < kPairYRows / 2; ++r) {
              const bool ok = (okm >> r) & 1u;
              const uint32_t dw = dwv[r], yw = ok ? ywv[r] : 0u;
              const float d0 = __uint_as_float(dw << 16), d1 = __uint_as_float(dw & 0xffff0000u);
              const float yc0 = ok ? __uint_as_float(yw << 16) - mu0 : 0.f, yc1 = ok ? __uint_as_float(yw & 0xffff0000u) - mu1 : 0.f;
              const float dz0 = fmaf(yc0, sc0, be0) > 0.f ? d0 : d0 * gneg;
              const float dz1 = fmaf(yc1, sc1, be1) > 0.f ? d1 : d1 * gneg;
              a10 += dz0; a20 = fmaf(dz0, yc0, a20);
              a11 += dz1; a21 = fmaf(dz1, yc1, a21);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&y_empty[s]);
          }
          // the two row halves, in order; comb is free (a dgrad launch takes no forward statistics)
          float4* c4 = reinterpret_cast<float4*>(comb);
          if (hh == 1) c4[cp] = make_float4(a10, a20 * rs0, a11, a21 * rs1);
          named_bar_sync(2, kEpiThreads);
          if (hh == 0 && tile_ok) {
            const float4 o = c4[cp];
            float4* dst = reinterpret_cast<float4*>(p.nr_part) + (int64_t(b) * T128 + t128) * 128 + cp;
            *dst = make_float4(a10 + o.x, a20 * rs0 + o.y, a11 + o.z, a21 * rs1 + o.w);
          }
        }
        if (dbg && et == 0 && k < 6) dbg[8 + k] = clock64();
        continue;
      }
      for (int c0 = half * 32; c0 < 256; c0 += 32 * (kEpiWarps / 4)) {
        float v[32];
        tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(buf * 256 + c0), v);
        const float4* sb4 = reinterpret_cast<const float4*>(sbias + c0);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const float4 bb = sb4[g4];
          v[4 * g4] += bb.x; v[4 * g4 + 1] += bb.y; v[4 * g4 + 2] += bb.z; v[4 * g4 + 3] += bb.w;
        }
        if (act == SG_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
        } else if (act == SG_ACT_LRELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], alpha * v[e]);
        } else if (act == SG_ACT_TANH) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = tanhf(v[e]);
        }
        uint32_t pk[16];
        if (!p.out_f32) {
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2) {
            pk[e2] = pack_bf16x2(v[2 * e2], v[2 * e2 + 1]);
            v[2 * e2] = __uint_as_float(pk[e2] << 16);
            v[2 * e2 + 1] = __uint_as_float(pk[e2] & 0xffff0000u);
          }
        }
        if (valid) {
          if (vec_ok && c0 + 32 <= p.Cout) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + obase + c0);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) dst[g4] = make_uint4(pk[4 * g4], pk[4 * g4 + 1], pk[4 * g4 + 2], pk[4 * g4 + 3]);
          } else if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.out) + obase;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e < p.Cout) dst[c0 + e] = v[e];
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + obase;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e < p.Cout) dst[c0 + e] = __float2bfloat16_rn(v[e]);
          }
        }
        if (has_stats) {
          // column sums over this warp's 32 rows through a private transposition tile (pitch 33: no conflicts)
#pragma unroll
          for (int e = 0; e < 32; ++e) tsw[e * 33 + lane] = v[e] * msk;
          __syncwarp();
          float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
          for (int r = 0; r < 32; r += 2) {
            const float x = tsw[lane * 33 + r], y = tsw[lane * 33 + r + 1];
            s1 += x; s2 += x * x;
            s1b += y; s2b += y * y;
          }
          red[(buf * 4 + q) * 256 + c0 + lane] = make_float2(s1 + s1b, s2 + s2b);
          __syncwarp();
        }
      }
      // this warp's share of the accumulator has been read: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0);
      if (has_stats) {
        named_bar_sync(1, kEpiThreads);  // also orders the re-use of red[buf] two tiles later
        if (tile_ok && et < p.Cout) {
          const float2 r0 = red[(buf * 4 + 0) * 256 + et], r1 = red[(buf * 4 + 1) * 256 + et];
          const float2 r2 = red[(buf * 4 + 2) * 256 + et], r3 = red[(buf * 4 + 3) * 256 + et];
          float2* dst = reinterpret_cast<float2*>(p.stats) + (int64_t(b) * p.stats_T + p.stats_t0 + t128) * p.Cout + et;
          *dst = make_float2((r0.x + r1.x) + (r2.x + r3.x), (r0.y + r1.y) + (r2.y + r3.y));
        }
      }
      if (dbg && et == 0 && k < 6) dbg[8 + k] = clock64();
    }
    if (staged) bulk_wait_group_read0();  // shared memory must outlive the reads of the last tile's copies
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_acc, 512);
  }
  if (dbg && threadIdx.x == 0) dbg[15] = clock64();
}

// =============================================================================================
// Transposed ("swap A/B") persistent kernel for layers with at most 128 output channels.
//
// An M = 128 pixels x N = Cout <= 128 instruction does not fill the tensor pipe (tc_probe mma_rate: 74 cycles at N = 128,
// 63 at N = 64 against the ideal 64 / 32; the "156 cycles whatever N is" this kernel was first designed around was the
// issue-loop artefact described at the top of the file).  Here
// the roles are exchanged: the WEIGHT tile is the M operand (128 channel rows, zero/garbage rows above
// Cout are never stored) and 256 PIXELS are the N operand, so every instruction is a full 128 x 256 x 16.
// Both operands are K-major SWIZZLE_128B tiles either way, and the row-shifted descriptor that walks the
// kw taps of a run now shifts the N operand.  The accumulator is D[channel lane][pixel column]:
//   * bias / activation / instance-norm statistics are per LANE: plain register sums, no transposition;
//   * the bf16 result is transposed through shared memory (2-byte stores, conflict-free) and copied out
//     with coalesced 16-byte stores.
// One 128 x 256 fp32 accumulator is 256 TMEM columns, so the kernel is persistent with TWO accumulators:
// the epilogue of tile k overlaps the MMAs of tile k + 1, and barrier/TMEM set-up is paid once per SM
// instead of once per tile (these layers have short K loops: set-up + epilogue used to be half their time).
constexpr int kSwapPix = 256;
constexpr int kSwapPStage = (kSwapPix + kHaloRows) * 128, kSwapWStage = 128 * 128;
constexpr int kSwapMaxWStages = 6;

__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_swap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA8,
                      const __grid_constant__ CUtensorMap tmW, const ConvGemmParams p, const int T256,
                      const int p_stages, const int w_stages, const int nblk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t p_full[4], p_empty[4], w_full[kSwapMaxWStages], w_empty[kSwapMaxWStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smP = smem;
  uint8_t* smW = smem + p_stages * kSwapPStage;
  uint8_t* S = smW + w_stages * kSwapWStage;          // staged bf16 tile [256 pixels][pitch]
  const int cw = p.Cout < 128 ? p.Cout : 128;              // channels per tile (multiple of 8)
  const int pitch = cw * 2 + 16;
  long long* rowoff = reinterpret_cast<long long*>(S + kSwapPix * pitch);  // [256] output element offsets
  float2* comb = reinterpret_cast<float2*>(rowoff + kSwapPix);             // [2][128] partial statistics
  const int cchunks = p.Cin / kChunkK;
  const int tstep = p.shift_kw > 0 ? kSwapPix - (p.shift_kw - 1) : kSwapPix;  // shift-sum tiles overlap by kw - 1 pixels
  const int total_tiles = p.B * T256 * nblk;  // tile t: channel block t % nblk (fastest, so the blocks of one pixel
                                              // tile run side by side and share it in L2), pixel tile t / nblk

  if (threadIdx.x == 0) {
    for (int s = 0; s < p_stages; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
    for (int s = 0; s < w_stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA8);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc(&tmem_base_sh, 512);
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;

  if (warp == 0) {
    // ------------------------------------------------------------ pixel-tile producer: one tile per (run, chunk)
    if (elect_one_sync()) {
      int g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int pt = t / nblk;
        const int b = pt / T256, m0 = (pt - b * T256) * tstep;
        for (int r = 0; r < p.nruns; ++r)
          for (int cc = 0; cc < cchunks; ++cc, ++g) {
            const int s = g % p_stages;
            mbar_wait(&p_empty[s], ((g / p_stages) & 1) ^ 1, 1);
            uint8_t* sp = smP + s * kSwapPStage;
            const int row0 = m0 + p.run_off[r];
            mbar_arrive_expect_tx(&p_full[s], kSwapPStage);
            tma_load_3d(&tmA, &p_full[s], sp, cc * kChunkK, row0, b);
            tma_load_3d(&tmA, &p_full[s], sp + kABytes, cc * kChunkK, row0 + kTileM, b);
            tma_load_3d(&tmA8, &p_full[s], sp + 2 * kABytes, cc * kChunkK, row0 + 2 * kTileM, b);
          }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ weight-tile producer: one tile per (tap, chunk)
    if (elect_one_sync()) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n0 = (t % nblk) * 128;
        for (int r = 0, t0 = 0; r < p.nruns; t0 += p.run_len[r], ++r)
          for (int cc = 0; cc < cchunks; ++cc)
            for (int q = 0; q < p.run_len[r]; ++q, ++it) {
              const int s = it % w_stages;
              mbar_wait(&w_empty[s], ((it / w_stages) & 1) ^ 1, 4);
              mbar_arrive_expect_tx(&w_full[s], kSwapWStage);
              tma_load_2d(&tmW, &w_full[s], smW + s * kSwapWStage, cc * kChunkK,
                          int(p.run_w[t0 + q]) * p.CoutPad + n0);
            }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc = idesc_bf16_f32(128, kSwapPix, 0, 0);
      int it = 0, g = 0, k = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++k) {
        const int buf = k & 1;
        mbar_wait(&acc_empty[buf], ((k >> 1) & 1) ^ 1, 6);
        tc_fence_after();
        const uint32_t d = tmem_acc + uint32_t(buf * kSwapPix);
        uint32_t acc = 0;
        for (int r = 0; r < p.nruns; ++r)
          for (int cc = 0; cc < cchunks; ++cc, ++g) {
            const int sp_i = g % p_stages;
            mbar_wait(&p_full[sp_i], (g / p_stages) & 1, 2);
            tc_fence_after();
            const uint32_t p_base = smem_u32(smP + sp_i * kSwapPStage);
            for (int q = 0; q < p.run_len[r]; ++q, ++it) {
              const int sw_i = it % w_stages;
              mbar_wait(&w_full[sw_i], (it / w_stages) & 1, 5);
              tc_fence_after();
              const uint64_t wdesc = desc_kmajor_sw128(smem_u32(smW + sw_i * kSwapWStage));
              const uint64_t pdesc = desc_kmajor_sw128(p_base + uint32_t(q) * 128u);  // tap q: q pixel rows in
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) {
                umma_bf16(d, wdesc + uint64_t(kk * 2), pdesc + uint64_t(kk * 2), idesc, acc);
                acc = 1;
              }
              umma_commit(&w_empty[sw_i]);
            }
            umma_commit(&p_empty[sp_i]);
          }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 3 + kEpiWarps) {
    // ------------------------------------------------------------ epilogue: lane = output channel
    const int q = warp & 3;
    const int half = (warp - 3) >> 2;
    const int ch = q * 32 + lane;
    const int et = threadIdx.x - 96;
    const bool has_stats = p.stats != nullptr;
    const int act = p.act;
    const float alpha = p.act_alpha;
    const int ppr = cw >> 3;  // 16-byte pieces per output position and channel block
    __nv_bfloat16* Sh = reinterpret_cast<__nv_bfloat16*>(S);
    const int pitch_h = pitch >> 1;
    int k = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++k) {
      const int buf = k & 1;
      const int pt = t / nblk, n0 = (t - pt * nblk) * 128;
      const int b = pt / T256, t256 = pt - b * T256;
      const int m0 = t256 * tstep;
      const bool ch_ok = n0 + ch < p.Cout;
      const float bch = (p.bias != nullptr && ch_ok) ? __ldg(p.bias + n0 + ch) : 0.f;
      mbar_wait(&acc_full[buf], (k >> 1) & 1, 3);
      tc_fence_after();
      if (p.shift_kw > 0) {
        // ---- shift-sum (7x7 output convolution): lane = kw * 4 + co holds that tap's partial product for
        // every pixel column; out[pixel][co] = sum_kw D[kw*4+co][pixel + kw].  Stage the 32 lanes as fp32
        // rows (pitch 257: conflict-free both ways), then one thread per pixel adds across kw.
        float* T = reinterpret_cast<float*>(S);
        if (q == 0) {
          for (int c0 = half * 32; c0 < kSwapPix; c0 += 64) {
            float v[32];
            tmem_ld32(tmem_acc + uint32_t(buf * kSwapPix + c0), v);
#pragma unroll
            for (int e = 0; e < 32; ++e) T[lane * 257 + c0 + e] = v[e];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        named_bar_sync(1, kEpiThreads);
        const int r = et;
        if (r < tstep) {
          const int m = m0 + r;
          const int i = m / p.P, j = m - i * p.P;
          if (m < p.M && i < p.Hv && j < p.Wv) {
            const int64_t ob = (int64_t(b) * p.omap.frame_pix + frame_pixel(p.omap, i, j)) * p.omap.C;
            for (int co = 0; co < p.Cout; ++co) {
              float x = p.bias != nullptr ? __ldg(p.bias + co) : 0.f;
              for (int kw = 0; kw < p.shift_kw; ++kw) x += T[(kw * 4 + co) * 257 + r + kw];
              x = apply_act(x, act, alpha);
              if (p.out_f32) reinterpret_cast<float*>(p.out)[ob + co] = x;
              else reinterpret_cast<__nv_bfloat16*>(p.out)[ob + co] = __float2bfloat16_rn(x);
            }
          }
        }
        named_bar_sync(1, kEpiThreads);  // T is rewritten by the next tile
        continue;
      }
      float s1 = 0.f, s2 = 0.f;
      for (int c0 = half * 32; c0 < kSwapPix; c0 += 64) {
        // validity of the 32 pixels of this chunk (the same for every channel lane)
        const int m = m0 + c0 + lane;
        const int i = m / p.P, j = m - i * p.P;
        const int oi = i * p.o_scale + p.o_a, oj = j * p.o_scale + p.o_b;
        const bool valid = (m < p.M) && (i < p.Hv) && (j < p.Wv) && (oi < p.omap.H) && (oj < p.omap.W);
        const uint32_t mask = __ballot_sync(0xffffffffu, valid);
        if (q == 0)
          rowoff[c0 + lane] = valid ? (int64_t(b) * p.omap.frame_pix + frame_pixel(p.omap, oi, oj)) * p.omap.C + n0 : -1;
        float v[32];
        tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(buf * kSwapPix + c0), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += bch;
        if (act == SG_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
        } else if (act == SG_ACT_LRELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], alpha * v[e]);
        } else if (act == SG_ACT_TANH) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = tanhf(v[e]);
        }
        __nv_bfloat16* sp = Sh + c0 * pitch_h + ch;
        if (ch_ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v[e]);
            sp[e * pitch_h] = h;
            v[e] = __bfloat162float(h);  // statistics over the values as stored
          }
        }
        if (has_stats) {
          if (mask == 0xffffffffu) {
#pragma unroll
            for (int e = 0; e < 32; ++e) { s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float x = (mask >> e) & 1u ? v[e] : 0.f;
              s1 += x; s2 = fmaf(x, x, s2);
            }
          }
        }
      }
      // this warp's share of the accumulator has been read: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (has_stats) comb[half * 128 + ch] = make_float2(s1, s2);
      named_bar_sync(1, kEpiThreads);  // tile staged, offsets and partial sums written
      if (has_stats && et < 128 && n0 + et < p.Cout) {
        const float2 a0 = comb[et], a1 = comb[128 + et];
        reinterpret_cast<float2*>(p.stats)[(int64_t(b) * p.stats_T + p.stats_t0 + t256) * p.Cout + n0 + et] =
            make_float2(a0.x + a1.x, a0.y + a1.y);
      }
      // coalesced copy-out: consecutive threads take consecutive 16-byte pieces of consecutive positions
      for (int piece = et; piece < kSwapPix * ppr; piece += kEpiThreads) {
        const int r = piece / ppr, part = piece - r * ppr;
        const long long off = rowoff[r];
        if (off >= 0)
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off + part * 8) =
              *reinterpret_cast<const uint4*>(S + r * pitch + part * 16);
      }
      named_bar_sync(1, kEpiThreads);  // the staging buffers are rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 512);
  }
}

// =============================================================================================
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                     const WgradParams p, const int stages, const uint32_t tmem_cols, const int NA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  // NA accumulators (128 x-channels each) share every dY tile: the kernel is L2-bandwidth bound, and two
  // accumulators cut the bytes per FLOP by a quarter
  const int a_bytes = NA * kABytes;
  const int stage_bytes = a_bytes + BN * 128;
  const int mtiles = p.x_pair ? 1 : p.Cx / (kTileM * NA);
  const int mt = blockIdx.x % mtiles, tap = blockIdx.x / mtiles;
  const int n0 = blockIdx.y * BN;
  const int nchunk = (p.Mpix + 63) / 64;  // 64-pixel K chunks per image
  const int total = p.B * nchunk;
  const int per = (total + p.ksplit - 1) / p.ksplit;
  const int kbeg = blockIdx.z * per;
  const int kend = min(total, kbeg + per);
  if (kbeg >= kend) return;  // uniform for the whole CTA
  const int ksteps = kend - kbeg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1) tmem_alloc(&tmem_base_sh, tmem_cols);
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;

  if (warp == 0) {
    if (elect_one_sync()) {
      const int xoff = p.x_off[tap], yoff = p.y_off[tap];
      const int xoff2 = p.x_pair ? p.x_off2[tap] : xoff;
      const int xc0 = p.x_pair ? 0 : mt * kTileM * NA;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 11);
        const int c = kbeg + ks;
        const int b = c / nchunk, mc = (c - b * nchunk) * 64;
        uint8_t* sa = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        if (p.x_pair) {
          tma_load_3d(&tmX, &full_bar[s], sa, 0, mc + xoff, b);
          tma_load_3d(&tmX, &full_bar[s], sa + 8192, 0, mc + xoff2, b);
        } else {
          for (int h = 0; h < 2 * NA; ++h) tma_load_3d(&tmX, &full_bar[s], sa + h * 8192, xc0 + h * 64, mc + xoff, b);
        }
        for (int h = 0; h < BN / 64; ++h)
          tma_load_3d(&tmY, &full_bar[s], sa + a_bytes + h * 8192, n0 + h * 64, mc + yoff, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = idesc_bf16_f32(kTileM, BN, 1, 1);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&full_bar[s], ph, 12);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint64_t bdesc = desc_mnmajor_sw128(sa + a_bytes, 8192);
        for (int a = 0; a < NA; ++a) {
          const uint64_t adesc = desc_mnmajor_sw128(sa + a * kABytes, 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 pixels (= 16 rows of 128 B) per MMA
            umma_bf16(tmem_acc + uint32_t(a * BN), adesc + uint64_t(k * (2048 >> 4)), bdesc + uint64_t(k * (2048 >> 4)),
                      idesc, (ks | k) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    mbar_wait(&acc_bar, 0, 13);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* obase = p.part ? p.part + int64_t(blockIdx.z) * p.part_stride : p.dW;
    for (int a = 0; a < NA; ++a) {
    const int xg = (mt * NA + a) * kTileM + row;
    float* dst = obase + int64_t(tap) * p.dw_tap_stride + int64_t(xg) * p.dw_sx;
    // y channels contiguous in dW and 16-byte aligned rows: 128-bit vector accesses
    const bool vec = (p.dw_sy == 1) && ((p.dw_sx & 3) == 0) && ((p.dw_tap_stride & 3) == 0) &&
                     ((p.part_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15) == 0);
    const bool plain = p.part != nullptr;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(a * BN + c0), v);  // warp-collective: no divergence here
      if (xg < p.nx_valid) {
        if (vec && n0 + c0 + 32 <= p.ny_valid) {
          float* d4 = dst + n0 + c0;
          if (plain) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(d4 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4 + 4 * g), "f"(v[4 * g]),
                           "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3])
                           : "memory");
          }
        } else {
          const int64_t sy = p.dw_sy;
          float* d1 = dst + int64_t(n0 + c0) * sy;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (n0 + c0 + e < p.ny_valid) {
              if (plain) d1[e * sy] = v[e];
              else atomicAdd(d1 + e * sy, v[e]);
            }
        }
      }
    }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
}

// =============================================================================================
// CTA-pair weight gradient for the 256 x 256-channel layers (the residual blocks).
//
// The single-CTA kernel above streams 64 KB of operands from L2 per 8 MMAs (X [64 px][256] + dY [64 px][256]): 8 KB per
// MMA against the ~46 B/clk an SM gets out of L2 when all of them pull (tc_probe tma_share) = 178 cycles per MMA at
// best, 230 measured, against the 128 the tensor pipe needs.  Here a cluster of two CTAs runs cta_group::2 MMAs with
// M = 256 input channels over the pair (128 each) and N = 256 output channels, of which each CTA stages only ITS HALF of
// the dY tile; the two accumulators of a CTA (2 x 256 TMEM columns) belong to TWO TAPS that share that dY tile.  Per SM
// and 64-pixel chunk: X(tap a) 16 KB + X(tap b) 16 KB + dY half 16 KB = 48 KB per 8 MMAs = 6 KB per MMA.
// Work unit: (pair of taps with the same dY offset, 256-channel block, split-K slice); an odd tap runs alone with one
// accumulator.  Partials leave exactly like the single-CTA kernel's (plain stores, fixed-order wgrad_reduce).
constexpr int kWgPairStages = 4;
constexpr int kWgPairOperand = 64 * 128 * 2;          // 64 pixels x 128 channels, bf16
constexpr int kWgPairStage = 3 * kWgPairOperand;      // X tap a | X tap b | dY half
constexpr int kWgPairSmem = kWgPairStages * kWgPairStage + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgradThreads, 1)
wgrad_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                       const WgradParams p, const int ngroups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full_bar[kWgPairStages];   // leader: both CTAs' operand bytes of a stage have landed
  __shared__ uint64_t empty_bar[kWgPairStages];  // both CTAs: the pair's MMAs on a stage have retired (multicast commit)
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cl = blockIdx.x >> 1;
  const int grp = cl % ngroups, xblk = cl / ngroups;
  const int ta = p.pair_a[grp], tb = int(p.pair_b[grp]) - 1;  // tb < 0: a single tap, one accumulator
  const int NA = tb >= 0 ? 2 : 1;
  const int n0 = blockIdx.y * 256;
  const int nchunk = (p.Mpix + 63) / 64;
  const int total = p.B * nchunk;
  const int per = (total + p.ksplit - 1) / p.ksplit;
  const int kbeg = blockIdx.z * per;
  const int kend = min(total, kbeg + per);
  const int ksteps = max(kend - kbeg, 0);  // prepare_wgrad_gemm gives every slice work; both CTAs of a pair agree

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_sh, 512);
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;
  const uint32_t stage_tx = uint32_t(NA + 1) * kWgPairOperand;  // bytes one CTA loads per stage

  if (warp == 0) {
    // ------------------------------------------------------------ producer (both CTAs: own 128 x / 128 y channels)
    if (elect_one_sync()) {
      const int xa = p.x_off[ta], xb = tb >= 0 ? p.x_off[tb] : 0, yoff = p.y_off[ta];
      const int xc = xblk * 256 + int(rank) * 128, yc = n0 + int(rank) * 128;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % kWgPairStages;
        mbar_wait(&empty_bar[s], ((ks / kWgPairStages) & 1) ^ 1, 14);
        const int c = kbeg + ks;
        const int b = c / nchunk, mc = (c - b * nchunk) * 64;
        uint8_t* sa = smem + s * kWgPairStage;
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_tx);
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
        for (int h = 0; h < 2; ++h) tma_load_3d_pair(&tmX, bar, sa + h * 8192, xc + h * 64, mc + xa, b);
        if (tb >= 0)
          for (int h = 0; h < 2; ++h) tma_load_3d_pair(&tmX, bar, sa + kWgPairOperand + h * 8192, xc + h * 64, mc + xb, b);
        for (int h = 0; h < 2; ++h) tma_load_3d_pair(&tmY, bar, sa + 2 * kWgPairOperand + h * 8192, yc + h * 64, mc + yoff, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && elect_one_sync()) {
      const uint32_t idesc = idesc_bf16_f32(256, 256, 1, 1);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % kWgPairStages;
        mbar_wait(&full_bar[s], (ks / kWgPairStages) & 1, 15);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kWgPairStage);
        const uint64_t bdesc = desc_mnmajor_sw128(sa + 2 * kWgPairOperand, 8192);
        for (int a = 0; a < NA; ++a) {
          const uint64_t adesc = desc_mnmajor_sw128(sa + a * kWgPairOperand, 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 pixels (= 16 rows of 128 B) per MMA
            umma_bf16_pair(tmem_acc + uint32_t(a * 256), adesc + uint64_t(k * (2048 >> 4)), bdesc + uint64_t(k * (2048 >> 4)),
                           idesc, (ks | k) != 0);
        }
        umma_commit_pair(&empty_bar[s]);
      }
      umma_commit_pair(&acc_bar);
    }
  } else if (ksteps > 0) {
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 rows of both taps)
    mbar_wait(&acc_bar, 0, 16);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* obase = p.part ? p.part + int64_t(blockIdx.z) * p.part_stride : p.dW;
    const bool vec = (p.dw_sy == 1) && ((p.dw_sx & 3) == 0) && ((p.dw_tap_stride & 3) == 0) && ((p.part_stride & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(obase) & 15) == 0);
    const bool plain = p.part != nullptr;
    const int xg = xblk * 256 + int(rank) * 128 + row;
    for (int a = 0; a < NA; ++a) {
      float* dst = obase + int64_t(a ? tb : ta) * p.dw_tap_stride + int64_t(xg) * p.dw_sx;
      for (int c0 = 0; c0 < 256; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(a * 256 + c0), v);  // warp-collective
        if (xg < p.nx_valid) {
          if (vec && n0 + c0 + 32 <= p.ny_valid) {
            float* d4 = dst + n0 + c0;
            if (plain) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(d4 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            } else {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4 + 4 * g), "f"(v[4 * g]),
                             "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3])
                             : "memory");
            }
          } else {
            const int64_t sy = p.dw_sy;
            float* d1 = dst + int64_t(n0 + c0) * sy;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (n0 + c0 + e < p.ny_valid) {
                if (plain) d1[e * sy] = v[e];
                else atomicAdd(d1 + e * sy, v[e]);
              }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();  // the peer may still be signalling this CTA's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_acc, 512);
  }
}

// =============================================================================================
// Host side
// Sort the taps by pixel offset and group consecutive offsets into runs (at most kHaloRows taps each).
static void build_runs(ConvGemmParams& p) {
  std::vector<std::pair<int, int>> t;
  for (int i = 0; i < p.ntaps; ++i) t.push_back({p.tap_off[i], int(p.tap_w[i])});
  std::sort(t.begin(), t.end());
  p.nruns = 0;
  int i = 0;
  while (i < p.ntaps) {
    int len = 1;
    while (i + len < p.ntaps && len < kHaloRows && t[i + len].first == t[i].first + len) ++len;
    p.run_off[p.nruns] = t[i].first;
    p.run_len[p.nruns] = uint8_t(len);
    ++p.nruns;
    i += len;
  }
  for (int k = 0; k < p.ntaps; ++k) p.run_w[k] = uint8_t(t[k].second);
}

int prepare_conv_gemm(const ConvGemmParams& pin, ConvGemmLaunch* L) {
  ConvGemmParams p = pin;
  const int esz = p.tf32 ? 4 : 2, ck = p.tf32 ? 32 : 64;  // element size, elements per 128-byte row
  if (p.Cin % ck != 0 || p.Cin <= 0) return -10;
  if (p.tf32 && (!p.out_f32 || p.shift_kw > 0)) return -15;
  if (!(p.BN == 32 || p.BN == 64 || p.BN == 128 || p.BN == 256)) return -11;
  if (p.CoutPad % p.BN != 0 || p.ntaps < 1 || p.ntaps > SGGAN_MAX_TAPS) return -12;
  build_runs(p);
  // two accumulators per CTA when that still leaves enough CTAs to fill the 148 SMs
  const int64_t ctas256 = int64_t((p.M + 255) / 256) * (p.CoutPad / p.BN) * p.B;
  if (p.MT != 128 && p.MT != 256) p.MT = ctas256 >= 120 ? 256 : 128;
  const int a_stage = (p.MT + kHaloRows) * 128, b_stage = p.BN * 128;
  // split the smem budget: at least 2 A stages, the rest to B (up to 8), then leftover back to A
  int sa = 2, sb = (kSmemBudget - sa * a_stage) / b_stage;
  if (sb > kMaxStages) sb = kMaxStages;
  if (sb < 2) return -13;
  sa = (kSmemBudget - sb * b_stage) / a_stage;
  if (sa > 4) sa = 4;
  L->p = p;
  L->sa_stages = sa;
  L->sb_stages = sb;
  L->tmem_cols = tmem_cols_for((p.MT / 128) * p.BN);
  L->smem = size_t(sa) * a_stage + size_t(sb) * b_stage + 1024;
  if (L->smem < size_t(p.BN) * 129 * 4 + 1024) L->smem = size_t(p.BN) * 129 * 4 + 1024;  // epilogue statistics tile
  {
    const size_t staged = size_t(p.MT) * (p.BN * 2 + 16) + 8192 + 1024;  // staged epilogue: bf16 tiles + partial sums + offsets
    if (L->smem < staged) L->smem = staged;
  }
  if (p.shift_kw > 0 && L->smem < size_t(p.MT) * 33 * 4 + 1024) L->smem = size_t(p.MT) * 33 * 4 + 1024;
  {
    const int mstep = p.shift_kw > 0 ? p.MT - (p.shift_kw - 1) : p.MT;
    L->grid_x = (p.M + mstep - 1) / mstep;
  }
  if (p.shift_kw > 0 && (p.BN != 32 || p.shift_kw * 4 > 32 || p.Cout > 4 || p.stats != nullptr)) return -14;
  L->grid_y = p.CoutPad / p.BN;
  L->grid_z = p.B;
  L->stat_tiles = L->grid_x * (p.MT / 128);
  // CTA-pair persistent kernel for the 256-channel layers when there is at least one tile per SM
  L->pair = 0;
  {
    const int T128 = (p.M + 127) / 128;
    // default (SGGAN_CONV_PAIR=0 selects the single-CTA kernel): with the issue loops under elect.sync the pair's MMAs
    // retire at the ideal 128 cycles each (probe PAIR stamps: 18432 cycles per 144-MMA tile) and only the last tile's
    // epilogue is exposed: 58 us against 73 us for the residual convolution (1.34 against 1.06 PFLOP/s)
    const char* env = getenv("SGGAN_CONV_PAIR");
    const bool allow = !(env && env[0] == '0');
    // also below one wave (batch-1 inference: 65 tiles): a pair finishes a 128-row tile in ~28k cycles (18.4k of MMAs + set-up
    // + epilogue), the single-CTA kernel needs ~60k for its two tiles, and neither fills the 148 SMs
    constexpr int kPairMinTiles = 8;
    if (allow && !p.tf32 && p.BN == 256 && p.CoutPad == 256 && p.shift_kw == 0 && int64_t(T128) * p.B >= kPairMinTiles) {
      L->pair = 1;
      L->T128 = T128;
      L->npairs = (T128 * p.B + 1) / 2;
      L->stat_tiles = T128;
      L->p.MT = 128;
    }
    // norm-backward fold: only on the pair kernel's staged epilogue, for a plain (unit-scale) bf16 dgrad without statistics
    L->nr_ok = (L->pair && p.nr_Y != nullptr && p.nr_part != nullptr && !p.out_f32 && (p.omap.C & 7) == 0 && p.Cout == 256 &&
                p.stats == nullptr && p.o_scale == 1 && p.o_a == 0 && p.o_b == 0) ? 1 : 0;
    if (!L->nr_ok) L->p.nr_Y = nullptr;
  }
  // transposed persistent kernel for layers with at most 128 output channels (bf16 output, one channel block)
  L->swap = 0;
  {
    const char* env = getenv("SGGAN_CONV_SWAP");
    const bool allow = !(env && env[0] == '0');
    const char* env256 = getenv("SGGAN_CONV_SWAP256");
    const bool wide_ok = (env256 && env256[0] == '1') && p.Cout % 128 == 0 && p.CoutPad == p.Cout;  // several channel blocks
    const bool plain_ok = ((p.CoutPad <= 128 && p.Cout <= 128) || wide_ok) && (p.Cout & 7) == 0 && p.shift_kw == 0 &&
                          !p.out_f32 && (p.omap.C & 7) == 0;
    const bool shift_ok = p.shift_kw > 0 && p.BN == 32 && p.shift_kw * 4 <= 32 && p.Cout <= 4 && p.stats == nullptr;
    const int tstep = p.shift_kw > 0 ? 256 - (p.shift_kw - 1) : 256;
    const int T = (p.M + tstep - 1) / tstep;
    if (allow && !p.tf32 && !L->pair && (plain_ok || shift_ok) && p.wt_taps * p.CoutPad >= 128 &&
        int64_t(T) * p.B * ((p.Cout + 127) / 128) >= 32) {
      L->swap = 1;
      L->T256 = T;
      L->stat_tiles = T;
      const int cw = p.Cout < 128 ? p.Cout : 128;
      L->swap_nblk = (p.Cout + 127) / 128;
      const int epi = shift_ok ? 32 * 257 * 4 + 1024 : 256 * (cw * 2 + 16) + 256 * 8 + 2048;
      // split the pipeline memory: 3 pixel stages (33 KB each; one serves run_len x 4 MMAs) when at least 3 weight
      // stages (16 KB each; one serves 4 MMAs) still fit, else 2
      const int pipe = 226 * 1024 - 1024 - epi;
      const char* ews = getenv("SGGAN_SWAP_WSTAGES");
      int ps = 3, ws = (pipe - ps * kSwapPStage) / kSwapWStage;
      if (ews) { ws = atoi(ews); ps = (pipe - ws * kSwapWStage) / kSwapPStage; }
      if (ws < 3) { ps = 2; ws = (pipe - ps * kSwapPStage) / kSwapWStage; }
      if (ws > kSwapMaxWStages) ws = kSwapMaxWStages;
      if (ps > 4) ps = 4;
      if (ps < 2 || ws < 2) L->swap = 0;
      L->swap_pstages = ps;
      L->swap_wstages = ws;
      L->swap_smem = size_t(ps) * kSwapPStage + size_t(ws) * kSwapWStage + epi + 1024;
    }
  }
  const uint64_t rs = uint64_t(p.a_row_stride) * esz, fs = uint64_t(p.a_frame_pix) * p.a_row_stride * esz;
  int r = make_tmap_3d(&L->tmA, p.A, p.Cin, p.a_frame_pix, p.B, rs, fs, ck, 128, p.tf32);
  if (r) return -1000 - r;
  r = make_tmap_3d(&L->tmA8, p.A, p.Cin, p.a_frame_pix, p.B, rs, fs, ck, kHaloRows, p.tf32);
  if (r) return -1500 - r;
  r = make_tmap_2d(&L->tmB, p.Wt, p.Cin, uint64_t(p.wt_taps) * p.CoutPad, uint64_t(p.Cin) * esz, ck, p.BN, p.tf32);
  if (r) return -2000 - r;
  if (L->pair || L->swap) {
    r = make_tmap_bf16_2d(&L->tmBh, p.Wt, p.Cin, uint64_t(p.wt_taps) * p.CoutPad, uint64_t(p.Cin) * 2, 64, 128);
    if (r) return -2500 - r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget + 1024);
    if (e != cudaSuccess) return -3000 - int(e);
    e = cudaFuncSetAttribute(conv_gemm_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem);
    if (e != cudaSuccess) return -3100 - int(e);
    e = cudaFuncSetAttribute(conv_gemm_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairFoldSmem);
    if (e != cudaSuccess) return -3100 - int(e);
    e = cudaFuncSetAttribute(conv_gemm_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return -3200 - int(e);
    attr_set = true;
  }
  return 0;
}

int run_conv_gemm(const ConvGemmLaunch& L, cudaStream_t st) {
  if (L.pair) {
    const int clusters = L.npairs < 74 ? L.npairs : 74;  // one CTA pair per TPC (148 SMs)
    cudaError_t e = L.nr_ok ? launch_kernel_pdl(conv_gemm_pair_kernel<true>, dim3(2 * clusters), dim3(kConvThreads),
                                                size_t(kPairFoldSmem), st, pdl_enabled(), L.tmA, L.tmA8, L.tmBh, L.p, L.T128, L.npairs)
                            : launch_kernel_pdl(conv_gemm_pair_kernel<false>, dim3(2 * clusters), dim3(kConvThreads),
                                                size_t(kPairSmem), st, pdl_enabled(), L.tmA, L.tmA8, L.tmBh, L.p, L.T128, L.npairs);
    return e == cudaSuccess ? 0 : -4100 - int(e);
  }
  if (L.swap) {
    const int tiles = L.T256 * L.p.B * L.swap_nblk;
    cudaError_t e = launch_kernel_pdl(conv_gemm_swap_kernel, dim3(tiles < 148 ? tiles : 148), dim3(kConvThreads), L.swap_smem,
                                      st, pdl_enabled(), L.tmA, L.tmA8, L.tmBh, L.p, L.T256, L.swap_pstages,
                                      L.swap_wstages, L.swap_nblk);
    return e == cudaSuccess ? 0 : -4200 - int(e);
  }
  dim3 grid(L.grid_x, L.grid_y, L.grid_z);
  cudaError_t e = launch_kernel_pdl(conv_gemm_tc_kernel, grid, dim3(kConvThreads), L.smem, st, pdl_enabled(), L.tmA, L.tmA8,
                                    L.tmB, L.p, L.sa_stages, L.sb_stages, L.tmem_cols);
  return e == cudaSuccess ? 0 : -4000 - int(e);
}

// Tap groups of the CTA-pair weight-gradient kernel: two taps that read dY at the same offset share a group (and the dY
// tile); 0 = the shape is not eligible.  pair_a[g] = first tap, pair_b[g] = second tap + 1 (0: none).
int wgrad_pair_groups(const WgradParams& p, uint8_t* pair_a, uint8_t* pair_b) {
  static const bool allow = []() { const char* e = getenv("SGGAN_WGRAD_PAIR"); return !(e && e[0] == '0'); }();
  if (!allow || p.x_pair || p.Cx <= 0 || p.Cx % 256 != 0 || p.BN != 256 || p.Cy % 256 != 0) return 0;
  bool used[SGGAN_MAX_TAPS] = {};
  int n = 0;
  for (int t = 0; t < p.ntaps; ++t) {
    if (used[t]) continue;
    used[t] = true;
    int mate = -1;
    for (int u = t + 1; u < p.ntaps; ++u)
      if (!used[u] && p.y_off[u] == p.y_off[t]) { mate = u; break; }
    if (mate >= 0) used[mate] = true;
    if (pair_a) { pair_a[n] = uint8_t(t); pair_b[n] = uint8_t(mate + 1); }
    ++n;
  }
  return n;
}

int prepare_wgrad_gemm(const WgradParams& p, WgradLaunch* L) {
  if (p.Cx <= 0 || (p.x_pair ? p.Cx != 64 : p.Cx % 128 != 0)) return -20;
  if (!(p.BN == 64 || p.BN == 128 || p.BN == 256) || p.Cy % p.BN != 0) return -21;
  if (p.ntaps < 1 || p.ntaps > SGGAN_MAX_TAPS || p.ksplit < 1) return -22;
  L->p = p;
  L->pair_groups = wgrad_pair_groups(p, L->p.pair_a, L->p.pair_b);
  if (L->pair_groups > 0) {
    L->na = 2;
    L->stages = kWgPairStages;
    L->tmem_cols = 512;
    L->smem = kWgPairSmem;
    L->grid_x = 2 * L->pair_groups * (p.Cx / 256);
    L->grid_y = p.Cy / 256;
    const int total = p.B * ((p.Mpix + 63) / 64);
    const int per = (total + p.ksplit - 1) / p.ksplit;
    L->p.ksplit = (total + per - 1) / per;
    L->grid_z = L->p.ksplit;
    int r = make_tmap_bf16_3d(&L->tmX, p.X, p.Cx, p.x_frame_pix, p.B, uint64_t(p.x_row_stride) * 2,
                              uint64_t(p.x_frame_pix) * p.x_row_stride * 2, 64, 64);
    if (r) return -1000 - r;
    r = make_tmap_bf16_3d(&L->tmY, p.Y, p.Cy, p.y_frame_pix, p.B, uint64_t(p.y_row_stride) * 2,
                          uint64_t(p.y_frame_pix) * p.y_row_stride * 2, 64, 64);
    if (r) return -2000 - r;
    static bool pair_attr_set = false;
    if (!pair_attr_set) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgPairSmem);
      if (e != cudaSuccess) return -3300 - int(e);
      pair_attr_set = true;
    }
    return 0;
  }
  L->na = (!p.x_pair && p.Cx % 256 == 0) ? 2 : 1;
  {
    int st = kSmemBudget / (L->na * kABytes + p.BN * 128);
    L->stages = st > kMaxStages ? kMaxStages : st;
  }
  L->tmem_cols = tmem_cols_for(L->na * p.BN);
  L->smem = size_t(L->stages) * (L->na * kABytes + p.BN * 128) + 1024;
  L->grid_x = (p.x_pair ? 1 : p.Cx / (128 * L->na)) * p.ntaps;
  L->grid_y = p.Cy / p.BN;
  {
    const int total = p.B * ((p.Mpix + 63) / 64);
    const int per = (total + p.ksplit - 1) / p.ksplit;
    L->p.ksplit = (total + per - 1) / per;  // every slice has work (the reduce adds all of them)
  }
  L->grid_z = L->p.ksplit;
  int r = make_tmap_bf16_3d(&L->tmX, p.X, p.Cx, p.x_frame_pix, p.B, uint64_t(p.x_row_stride) * 2,
                            uint64_t(p.x_frame_pix) * p.x_row_stride * 2, 64, 64);
  if (r) return -1000 - r;
  r = make_tmap_bf16_3d(&L->tmY, p.Y, p.Cy, p.y_frame_pix, p.B, uint64_t(p.y_row_stride) * 2,
                        uint64_t(p.y_frame_pix) * p.y_row_stride * 2, 64, 64);
  if (r) return -2000 - r;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget + 1024);
    if (e != cudaSuccess) return -3000 - int(e);
    attr_set = true;
  }
  return 0;
}

int run_wgrad_gemm(const WgradLaunch& L, cudaStream_t st) {
  dim3 grid(L.grid_x, L.grid_y, L.grid_z);
  if (L.pair_groups > 0) {
    cudaError_t e = launch_kernel_pdl(wgrad_gemm_pair_kernel, grid, dim3(kWgradThreads), L.smem, st, pdl_enabled(), L.tmX, L.tmY,
                                      L.p, L.pair_groups);
    return e == cudaSuccess ? 0 : -4300 - int(e);
  }
  cudaError_t e = launch_kernel_pdl(wgrad_gemm_tc_kernel, grid, dim3(kWgradThreads), L.smem, st, pdl_enabled(), L.tmX, L.tmY,
                                    L.p, L.stages, L.tmem_cols, L.na);
  return e == cudaSuccess ? 0 : -4000 - int(e);
}

template <bool STREAM>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, int64_t stride,
                                                           int64_t n, float* dW) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 acc = reinterpret_cast<float4*>(dW)[i];
    for (int z = 0; z < nsplit; ++z) {
      const float4 v = STREAM ? __ldcs(reinterpret_cast<const float4*>(part + z * stride) + i)  // read once: evict first
                              : __ldg(reinterpret_cast<const float4*>(part + z * stride) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dW)[i] = acc;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float a = dW[i];
      for (int z = 0; z < nsplit; ++z) a += part[z * stride + i];
      dW[i] = a;
    }
}
void launch_wgrad_reduce(const WgradLaunch& L, int64_t numel, cudaStream_t st) {
  int blocks = int((numel / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  static const bool stream = []() { const char* e = getenv("SGGAN_REDUCE_STREAM"); return !(e && e[0] == '0'); }();
  if (stream)
    launch_kernel_pdl(wgrad_reduce_kernel<true>, dim3(blocks), dim3(256), 0, st, pdl_enabled(), (const float*)L.p.part,
                      L.p.ksplit, L.p.part_stride, numel, L.p.dW);
  else
    launch_kernel_pdl(wgrad_reduce_kernel<false>, dim3(blocks), dim3(256), 0, st, pdl_enabled(), (const float*)L.p.part,
                      L.p.ksplit, L.p.part_stride, numel, L.p.dW);
}

int read_tc_watchdog() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_watchdog_flag, sizeof(int));
  return v;
}

}  // namespace sggan
