// kparams.h -- plain-old-data launch parameters shared by the CUDA kernels (csrc/*.cu) and by
// the CPU emulation of the same kernels that the tests use to check the host-side planning
// (tests/emu).  No CUDA types here: pointers are raw addresses in whichever memory the build
// targets.
#pragma once
#include <stdint.h>

#define SGGAN_MAX_TAPS 49

typedef uint16_t sg_bf16;  // raw bfloat16 bits

// Activation codes used by epilogues / glue kernels.
enum { SG_ACT_NONE = 0, SG_ACT_RELU = 1, SG_ACT_LRELU = 2, SG_ACT_TANH = 3 };

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM convolution over "pitch-linearised frames".
//
// The input of a convolution lives in frames: image b occupies a_frame_pix pixels of Cin bf16
// channels, rows have pitch P pixels and already contain whatever border (reflect / zero) the
// convolution needs.  Output position m = i*P + j (i row, j column on the SAME pitch) reads, for
// tap t, the input pixel m + tap_off[t]; hence every A tile of 128 consecutive m is one dense
// [128 x 64ch] TMA box.  Columns j >= Wv and rows i >= Hv are pitch slack: computed, never
// stored, excluded from the statistics.
//
//   out[b, i, j, n] = act( bias[n] + sum_t sum_c A[b, m + tap_off[t], c] * Wt[t, n, c] )
//
// Strided / transposed convolutions reduce to this form by phase-splitting the frames (host
// plan), so this one kernel serves conv fwd, dgrad and deconv.
struct ConvGemmParams {
  const sg_bf16* A;     // frames [B][a_frame_pix][Cin]
  int64_t a_frame_pix;  // pixels per frame (TMA zero-fills beyond)
  int Cin;              // multiple of 64
  int B;
  const sg_bf16* Wt;  // [ntaps][CoutPad][Cin]  (K-major rows)
  int ntaps;
  int CoutPad;  // rows per tap in Wt, multiple of BN
  int Cout;     // valid output channels (n < Cout stored)
  int BN;       // N tile: 32, 64, 128 or 256
  int tap_off[SGGAN_MAX_TAPS];
  int M;       // linear output positions per image
  int P;       // pitch used to decode m -> (i, j)
  int Hv, Wv;  // valid rows / columns
  // output addressing (elements): out + b*out_bstride + i*out_sy + j*out_sx + out_off + n
  void* out;
  int out_f32;  // 0: bf16 output, 1: fp32 output
  int64_t out_bstride, out_sy, out_sx, out_off;
  const float* bias;  // [Cout] or null
  float* stats;       // [B][Cout][2] running (sum, sum of squares) of the fp32 result, or null
  int act;
  float act_alpha;  // leaky slope
};

// ---------------------------------------------------------------------------------------------
// Weight gradient as an MN-major GEMM, K = pixels, split-K with fp32 atomics.
//
//   dW[t, x, y] += sum_b sum_{m < Mpix} X[b, m + x_off[t], x] * Y[b, m + y_off, y]
//
// X (activation frames) takes the M role, Y (output-gradient frames) the N role; both share the
// pitch so tap offsets are constants.  Y must be zero at slack positions.
struct WgradParams {
  const sg_bf16* X;
  int64_t x_frame_pix;
  int Cx;  // multiple of 128
  const sg_bf16* Y;
  int64_t y_frame_pix;
  int Cy;  // multiple of BN
  int BN;  // 64, 128 or 256
  int B;
  int ntaps;
  int x_off[SGGAN_MAX_TAPS];
  int y_off;
  int Mpix;  // linear positions per image covered (rounded up to 64 inside)
  float* dW;
  int64_t dw_tap_stride, dw_sx, dw_sy;  // element strides of dW for (tap, x channel, y channel)
  int ksplit;                           // CTAs along the split-K (grid.z)
};
