// kparams.h -- plain-old-data launch parameters of every kernel in the SG-GAN step.
// No CUDA types: pointers are raw device addresses, so the host-side planner (engine.cu) and
// the tests can fill and inspect them without a GPU.
#pragma once
#include <stdint.h>

#define SGGAN_MAX_TAPS 49

typedef uint16_t sg_bf16;  // raw bfloat16 bits

enum { SG_ACT_NONE = 0, SG_ACT_RELU = 1, SG_ACT_LRELU = 2, SG_ACT_TANH = 3 };

// ---------------------------------------------------------------------------------------------
// Frames: the HBM layout of every activation / gradient that a convolution consumes.
//
// A frame stores one H x W x C image per `frame_pix` pixels.  kind 0 is a single plane with
// pitch P whose logical pixel (0,0) sits at row pt, column pl; everything outside the logical
// image is border (zeros, or the reflection written by the producer when reflect > 0) or pitch
// slack.  kind 1 holds the four 2x2 phases of the image as separate planes (plane_pix apart),
// plane (a,b) containing pixels (2i+a, 2j+b) at (i+pt, j+pl) -- the layout a stride-2
// convolution reads with unit stride.  Because producer and consumer share the pitch, tap
// (kh,kw) of a convolution is a constant pixel offset: that is what lets one dense 2-D TMA box
// feed the tensor cores with no im2col and no separate padding pass (tf.pad at
// module.py:210,214,230,262 costs nothing here).
struct FrameMap {
  int64_t frame_pix;  // pixels per image
  int C;              // channels stored per pixel
  int H, W;           // logical size
  int kind;           // 0 plane, 1 phase-split
  int P;              // pitch in pixels
  int pt, pl;         // origin inside the plane
  int plane_pix;      // kind 1: pixels per phase plane
  int reflect;        // kind 0: reflect border width the producer must write (0 = zero border)
};

#if defined(__CUDACC__)
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD inline
#endif
SG_HD int64_t frame_pixel(const FrameMap& f, int i, int j) {
  if (f.kind == 0) return int64_t(i + f.pt) * f.P + (j + f.pl);
  return int64_t((i & 1) * 2 + (j & 1)) * f.plane_pix + int64_t((i >> 1) + f.pt) * f.P + ((j >> 1) + f.pl);
}

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM convolution over pitch-linearised frames (conv_gemm_tc.cu).
//
//   acc[b, m, n] = sum_t sum_c A[b, m + tap_off[t], c] * Wt[tap_w[t], n, c]        m = i*P + j
//   out[b, (i*o_scale+o_a, j*o_scale+o_b) via omap, n] = act(acc + bias[n])     (i < Hv, j < Wv)
//
// "Channels" c are whatever 64-element rows the tensor map exposes: real channels (row stride =
// Cin) or, for the 3-channel layers, a sliding window of 8 pixels x 8 padded channels (row
// stride 8), which turns the kw taps of a 7x7 / 3x3 filter into K of one tap.
struct ConvGemmParams {
  const sg_bf16* A;
  int64_t a_frame_pix;   // rows (pixels) per image visible to TMA
  int64_t a_row_stride;  // elements between consecutive rows (= real channel count per pixel)
  int Cin;               // K per tap, multiple of 64
  int B;
  const sg_bf16* Wt;  // [wt_taps][CoutPad][Cin]
  int wt_taps;        // number of tap slabs in Wt
  int ntaps;          // taps used by this launch
  int CoutPad;        // rows per tap slab, multiple of BN
  int Cout;           // valid output channels
  int BN;             // N tile: 32, 64, 128, 256
  int tap_off[SGGAN_MAX_TAPS];
  uint8_t tap_w[SGGAN_MAX_TAPS];  // slab index in Wt for tap t
  // Derived by prepare_conv_gemm: taps sorted by offset and grouped into runs of consecutive pixel
  // offsets.  One A tile (MT + 8 rows) is loaded per (run, channel chunk) and every tap of the run
  // reads it through a row-shifted shared-memory descriptor (the kw taps of a filter row share one load).
  int nruns;
  int run_off[SGGAN_MAX_TAPS];
  uint8_t run_len[SGGAN_MAX_TAPS];
  uint8_t run_w[SGGAN_MAX_TAPS];  // slab index of the sorted taps, runs back to back
  int MT;                         // rows of the CTA tile: 128 (one accumulator) or 256 (two)
  int M;                          // linear output positions per image
  int P;                          // pitch used to decode m -> (i, j)
  int Hv, Wv;                     // valid rows / columns
  void* out;
  int out_f32;  // 0 bf16, 1 fp32
  FrameMap omap;
  int o_scale, o_a, o_b;
  const float* bias;  // [Cout] or null
  float* stats;       // per-tile partial (sum, sum of squares) of the stored result, [B][stats_T][Cout][2], or null;
                      // the norm-apply pass adds the tiles in a fixed order (deterministic forward)
  int stats_T;        // tiles per image over all launches of the layer
  int stats_t0;       // first tile index of this launch
  int act;
  float act_alpha;
  // Shift-sum mode (the 7x7, 64 -> 3 output convolution): the N dimension holds (kw, co) pairs (co padded to
  // 4), each filter ROW is one tap, and the epilogue adds the kw partial products of neighbouring rows:
  //   out[m, co] = act(bias[co] + sum_kw acc[m + kw, kw*4 + co]).  CTA tiles then advance by MT - (shift_kw-1).
  // 7x fewer tensor-core instructions than one tap per (kh, kw) at N = 32 (tcgen05.mma has a ~156-cycle
  // floor per instruction regardless of N).
  int shift_kw;
  // tf32 mode (the fp32-accurate tier): A and Wt hold fp32 elements (pre-rounded to tf32), K per tap is a multiple of 32,
  // the tensor cores run kind::tf32, the output is fp32.  conv_gemm_tc_kernel only.
  int tf32;
  // CTA-pair kernel, dgrad launches only (nr_Y != null): this launch produces the gradient w.r.t. the padded input grid
  // of a convolution whose input is act(instance_norm(Y)) + a reflect border of nr_pad pixels.  The epilogue then also
  // accumulates that norm layer's backward sums per 128-position tile and channel: (sum dz, sum dz * xhat) with
  // dz = act'(z) * dX at the source pixel of every (reflected) grid position -- the "reduce" pass of the norm backward
  // (glue_rows.cu BWD_REDUCE) without re-reading dX.  nr_part: [B][tiles per image][C][2]; nr_gneg: slope of the
  // activation on its non-positive side (ReLU 0, LeakyReLU alpha, none 1).
  const sg_bf16* nr_Y;      // [B][nr_H][nr_W][Cout] raw convolution output of the layer below
  const float* nr_stats;    // [B][Cout][2] (sum y, sum y^2)
  const float* nr_gamma;
  const float* nr_beta;
  float* nr_part;
  float nr_eps, nr_gneg;
  int nr_H, nr_W, nr_pad;
  long long* dbg;  // optional per-CTA clock64 stamps [grid][8] (tests/gpu/tc_probe.cu); null in production
};

// ---------------------------------------------------------------------------------------------
// Weight gradient (wgrad_gemm_tc in conv_gemm_tc.cu): MN-major GEMM, K = pixels, split-K.
//
//   dW[t, x, y] += sum_b sum_{m < Mpix} X[b, m + x_off[t], x] * Y[b, m + y_off[t], y]
//
// X takes the M role (128-row tiles), Y the N role.  pair mode (x_pair != 0): the 128 M rows are
// two 64-element windows of X taken at pixel offsets x_off[t] and x_off2[t] (used by the
// 3-channel layers, whose real channel count is below the 128-row MMA shape).
struct WgradParams {
  const sg_bf16* X;
  int64_t x_frame_pix, x_row_stride;
  int Cx;  // multiple of 128 (64 in pair mode)
  const sg_bf16* Y;
  int64_t y_frame_pix, y_row_stride;
  int Cy;  // multiple of BN
  int BN;  // 64, 128, 256
  int B;
  int ntaps;
  int x_off[SGGAN_MAX_TAPS];
  int x_off2[SGGAN_MAX_TAPS];
  int y_off[SGGAN_MAX_TAPS];
  int x_pair;
  uint8_t pair_a[SGGAN_MAX_TAPS], pair_b[SGGAN_MAX_TAPS];  // CTA-pair kernel: tap groups (filled by prepare_wgrad_gemm)
  int nx_valid, ny_valid;  // rows / columns of the tile actually accumulated into dW
  int Mpix;
  float* dW;
  int64_t dw_tap_stride, dw_sx, dw_sy;
  int ksplit;
  // Split-K partials: when `part` is set, slice z of the K range stores its tile with plain 128-bit stores to
  // part + z * part_stride (same element strides as dW) and launch_wgrad_reduce adds the slices in a fixed
  // order (deterministic, and ~5x cheaper than 128 KB of fp32 atomics per CTA).  Null: red.global into dW.
  float* part;
  int64_t part_stride;
};

// ---------------------------------------------------------------------------------------------
// Glue kernels (glue.cu)

// Where a gradient w.r.t. a layer's (post-activation) output comes from: a plain [B][Hs][Ws][C]
// buffer whose logical pixel (0,0) sits at (oy, ox); fold > 0 adds the reflected border back
// (gradient of tf.pad REFLECT, Appendix A.4).
struct GradSrc {
  const void* ptr;  // null = absent
  int f32;          // element type: 0 bf16, 1 fp32
  int Hs, Ws;
  int oy, ox;
  int fold;
};

// Instance norm (+ activation, + residual) of a raw convolution output into the next frame.
//   z = act(gamma * (y - mean) * rstd + beta) (+ res);   mean / var from stats (sum, sum^2).
// tfa.layers.InstanceNormalization at module.py:212,216,233,...; Activation / LeakyReLU / `y + x`
// at module.py:213,217,...
struct InApplyParams {
  const sg_bf16* Y;  // [B][H][W][C] raw conv output
  int B, H, W, C;
  const float* stats;  // [B][C][2]
  // optional: per-tile partial statistics [B][stats_T][C][2] of the producing convolution; the kernel adds them in
  // a fixed order itself (and writes the result to `stats_out` for the backward pass) instead of a separate launch
  const float* stats_part;
  int stats_T;
  float* stats_out;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  float act_alpha;
  const sg_bf16* res;  // residual frames (interior read) or null
  FrameMap rmap;
  sg_bf16* dst;
  FrameMap dmap;
};

// Backward of the same: two passes.
//   dzh = act'(.) * (dz1 + dz2);   sums[b][c] = (sum dzh, sum dzh * xhat)          (reduce)
//   dy  = gamma * rstd * (dzh - mean(dzh) - xhat * mean(dzh * xhat))  -> dY frame  (apply)
// Activations of "virtual" image b >= nb_act are those of image b - act_wrap (the generator-loss
// pass through the discriminator re-uses the fake half of the batch).
struct InBwdParams {
  const sg_bf16* Y;
  int B, H, W, C;
  int nb_act, act_wrap;
  const float* stats;  // forward (sum, sum^2) per activation image
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  float act_alpha;
  GradSrc g1, g2;
  float* sums;       // [B][C][2]: reduced (sum dzh, sum dzh * xhat), published by the apply pass (for dgamma / dbeta)
  float* sums_part;  // [B][sums_nblk][C][2] scratch: per-block partials of the reduce pass (in_bwd_partials_bytes(C))
  int sums_nblk;     // apply pass: value returned by launch_in_bwd_reduce
  int* sync_ctr;     // fused form (launch_in_bwd_fused): [B] arrival counters, zero before the launch
  sg_bf16* dst;
  FrameMap dmap;
  // optional (reduce pass): also store the summed, border-folded gradient g1 + g2 as plain [B][H][W][C] bf16 -- the
  // residual-stream gradient gather of the generator blocks, fused with the statistics of the layer that reads it
  sg_bf16* gather_dst;
};
