// engine.cu -- plan + launch sequences of the SG-GAN training step (model.py:169-200) on one GPU.
//
// Data layout in HBM (all carved from the caller's workspace, nothing allocated here):
//   * flat fp32 params / grads / Adam m / Adam v per net, tensors back to back in Keras creation
//     order -- the gradient buffers are what a data-parallel caller all-reduces;
//   * per conv layer: bf16 GEMM weight slabs (forward + dgrad), input frame X, raw output Y,
//     output-gradient frame dY, input-gradient buffer dX, instance-norm statistics;
//   * frames are zero-initialised once; producers only ever write logical pixels (and reflected
//     borders), so zero borders / pitch slack stay zero for the life of the handle.
#include "engine.h"

#include <cstdlib>

#include <math.h>
#include <string.h>

namespace sggan {

static FrameMap plain_map(int H, int W, int C) {
  FrameMap m;
  memset(&m, 0, sizeof(m));
  m.frame_pix = int64_t(H) * W;
  m.C = C;
  m.H = H;
  m.W = W;
  m.kind = 0;
  m.P = W;
  return m;
}
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// geometry
int layer_geometry(Layer& l) {
  const int k = l.k;
  memset(&l.xmap, 0, sizeof(FrameMap));
  memset(&l.dymap, 0, sizeof(FrameMap));
  l.CoutK = round_up(l.Cout, 64);
  l.dx_f32 = 0;
  switch (l.type) {
    case LT_S1:
    case LT_WIN_C1:
    case LT_OUT7: {
      const int p = l.pad == PAD_VALID ? 0 : (k - 1) / 2;
      l.Hout = l.Hin + 2 * p - k + 1;
      l.Wout = l.Win + 2 * p - k + 1;
      if (l.Hout < 1 || l.Wout < 1) return SGGAN_E_INVALID;
      l.P = l.Wout + k - 1;
      l.xmap.frame_pix = int64_t(l.Hin + 2 * p) * l.P;
      l.xmap.C = l.type == LT_WIN_C1 ? 8 : l.Cin;
      l.xmap.H = l.Hin; l.xmap.W = l.Win; l.xmap.kind = 0; l.xmap.P = l.P; l.xmap.pt = p; l.xmap.pl = p;
      l.xmap.reflect = l.pad == PAD_REFLECT ? p : 0;
      l.dymap.frame_pix = int64_t(l.Hout + 2 * (k - 1)) * l.P;
      l.dymap.C = l.type == LT_OUT7 ? 8 : l.CoutK;
      l.dymap.H = l.Hout; l.dymap.W = l.Wout; l.dymap.kind = 0; l.dymap.P = l.P; l.dymap.pt = k - 1; l.dymap.pl = 0;
      l.dxH = l.Hout + k - 1; l.dxW = l.P; l.dx_oy = p; l.dx_ox = p; l.dx_fold = l.pad == PAD_REFLECT ? p : 0;
      break;
    }
    case LT_S2:
    case LT_WIN_H0: {
      if (k != 3) return SGGAN_E_INVALID;
      if (l.pad == PAD_VALID) { l.Hout = (l.Hin - 3) / 2 + 1; l.Wout = (l.Win - 3) / 2 + 1; }
      else { if ((l.Hin | l.Win) & 1) return SGGAN_E_INVALID; l.Hout = l.Hin / 2; l.Wout = l.Win / 2; }
      if (l.Hout < 1 || l.Wout < 1) return SGGAN_E_INVALID;
      l.P = l.Wout + 2;
      const int rows2 = l.Hout + 2;
      if ((l.Hin + 1) / 2 > rows2 || (l.Win + 1) / 2 > l.P) return SGGAN_E_INVALID;
      l.xmap.plane_pix = rows2 * l.P;
      l.xmap.frame_pix = int64_t(4) * l.xmap.plane_pix;
      l.xmap.C = l.type == LT_WIN_H0 ? 8 : l.Cin;
      l.xmap.H = l.Hin; l.xmap.W = l.Win; l.xmap.kind = 1; l.xmap.P = l.P;
      l.dymap.frame_pix = int64_t(l.Hout + 2) * l.P;
      l.dymap.C = l.CoutK; l.dymap.H = l.Hout; l.dymap.W = l.Wout; l.dymap.kind = 0; l.dymap.P = l.P; l.dymap.pt = 1;
      l.dxH = l.Hin; l.dxW = l.Win; l.dx_oy = 0; l.dx_ox = 0; l.dx_fold = 0;
      if (l.type == LT_WIN_H0) l.dx_f32 = 1;
      break;
    }
    case LT_DECONV: {
      if (k != 3) return SGGAN_E_INVALID;
      l.Hout = 2 * l.Hin; l.Wout = 2 * l.Win;
      l.P = l.Win + 2;
      l.xmap.frame_pix = int64_t(l.Hin + 2) * l.P;
      l.xmap.C = l.Cin; l.xmap.H = l.Hin; l.xmap.W = l.Win; l.xmap.kind = 0; l.xmap.P = l.P; l.xmap.pt = 1;
      const int rows2 = l.Hin + 2;
      l.dymap.plane_pix = rows2 * l.P;
      l.dymap.frame_pix = int64_t(4) * l.dymap.plane_pix;
      l.dymap.C = l.CoutK; l.dymap.H = l.Hout; l.dymap.W = l.Wout; l.dymap.kind = 1; l.dymap.P = l.P;
      l.dxH = l.Hin; l.dxW = l.Win; l.dx_oy = 0; l.dx_ox = 0; l.dx_fold = 0;
      break;
    }
  }
  // weight slab shapes
  memset(&l.packf, 0, sizeof(PackParams));
  memset(&l.packd, 0, sizeof(PackParams));
  l.packf.KH = l.packd.KH = k; l.packf.KW = l.packd.KW = k;
  l.packf.Cin = l.packd.Cin = l.Cin; l.packf.Cout = l.packd.Cout = l.Cout;
  l.unpack_mode = -1;
  l.wscratch_elems = 0;
  {  // upper bound on the number of 128-row tiles the forward launches produce per image
    const int m = (l.type == LT_DECONV ? l.Hin : l.Hout) * l.P;
    l.stats_T_max = (l.type == LT_DECONV ? 4 : 1) * ((m + 127) / 128 + 1);
    l.stats_T = 0;
  }
  switch (l.type) {
    case LT_S1:
    case LT_S2:
      l.CoutN = l.CoutK; l.CinN = l.Cin;
      l.packf.mode = 0; l.packf.T = k * k; l.packf.N = l.CoutN; l.packf.K = l.Cin;
      l.packd.mode = 1; l.packd.T = k * k; l.packd.N = l.CinN; l.packd.K = l.CoutK;
      break;
    case LT_DECONV:
      l.CoutN = l.CoutK; l.CinN = l.Cin;
      l.packf.mode = 2; l.packf.T = 9; l.packf.N = l.CoutN; l.packf.K = l.Cin;
      l.packd.mode = 3; l.packd.T = 9; l.packd.N = l.CinN; l.packd.K = l.CoutK;
      break;
    case LT_WIN_C1:
      l.CoutN = l.CoutK; l.CinN = 0;
      l.packf.mode = 4; l.packf.T = k; l.packf.N = l.CoutN; l.packf.K = 64;
      l.packd.T = 0;
      l.unpack_mode = 0; l.wscratch_elems = ((k + 1) / 2) * 128 * 64;
      break;
    case LT_OUT7:
      l.CoutN = 32; l.CinN = l.Cin;
      l.packf.mode = 7; l.packf.T = k; l.packf.N = 32; l.packf.K = l.Cin;  // rows = (kw, co) pairs, one slab per kh
      l.packd.mode = 5; l.packd.T = k; l.packd.N = l.Cin; l.packd.K = 64;
      l.unpack_mode = 1; l.wscratch_elems = ((k + 1) / 2) * 128 * 64;
      break;
    case LT_WIN_H0:
      l.CoutN = l.CoutK; l.CinN = 32;
      l.packf.mode = 6; l.packf.T = 2 * k; l.packf.N = l.CoutN; l.packf.K = 64;
      l.packd.mode = 1; l.packd.T = k * k; l.packd.N = 32; l.packd.K = l.CoutK;
      l.unpack_mode = 2; l.wscratch_elems = k * 128 * 64;
      break;
  }
  return 0;
}

static void conv_common(ConvGemmParams& p, const Layer& l) {
  memset(&p, 0, sizeof(p));
  p.o_scale = 1;
  p.act = SG_ACT_NONE;
  p.act_alpha = l.alpha;
}
static int bn_for(int n) { return n >= 256 ? 256 : n; }

// forward launches.  `out`/`omap`: destination of the conv result (raw Y, the next frame, or an fp32
// plain tensor).
static int layer_prepare_fwd_impl(Layer& l, const float* bias, void* out, const FrameMap& omap, int out_f32, bool dry);
int layer_prepare_fwd(Layer& l, const float* bias, void* out, const FrameMap& omap, int out_f32, bool dry) {
  int r = layer_prepare_fwd_impl(l, bias, out, omap, out_f32, dry);
  if (r || dry) return r;
  int T = 0;
  for (auto& L : l.fwd) { L.p.stats_t0 = T; T += L.stat_tiles; }
  for (auto& L : l.fwd) L.p.stats_T = T;
  l.stats_T = T;
  if (l.has_norm && T > l.stats_T_max) return SGGAN_E_WORKSPACE;
  return 0;
}
static int layer_prepare_fwd_impl(Layer& l, const float* bias, void* out, const FrameMap& omap, int out_f32, bool dry) {
  l.fwd.clear();
  const int k = l.k, P = l.P;
  ConvGemmParams p;
  conv_common(p, l);
  p.A = l.X; p.a_frame_pix = l.xmap.frame_pix; p.a_row_stride = l.xmap.C; p.B = l.nb;
  p.Wt = l.Wf; p.wt_taps = l.packf.T; p.CoutPad = l.CoutN; p.Cout = l.Cout; p.BN = bn_for(l.CoutN);
  p.P = P; p.out = out; p.out_f32 = out_f32; p.omap = omap; p.bias = bias;
  p.stats = l.has_norm ? l.stats_part : nullptr;
  if (!l.has_norm) p.act = l.act;
  auto push = [&](const ConvGemmParams& q) -> int {
    ConvGemmLaunch L;
    if (dry) { L.p = q; l.fwd.push_back(L); return 0; }
    int r = prepare_conv_gemm(q, &L);
    if (r) return r;
    l.fwd.push_back(L);
    return 0;
  };
  switch (l.type) {
    case LT_OUT7:
      // one tap per filter row; the kw taps live in N and are added across rows by the shift-sum epilogue
      p.Cin = l.Cin; p.ntaps = k; p.shift_kw = k;
      for (int kh = 0; kh < k; ++kh) { p.tap_off[kh] = kh * P; p.tap_w[kh] = uint8_t(kh); }
      p.M = l.Hout * P; p.Hv = l.Hout; p.Wv = l.Wout;
      return push(p);
    case LT_S1:
      p.Cin = l.Cin; p.ntaps = k * k;
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) { p.tap_off[kh * k + kw] = kh * P + kw; p.tap_w[kh * k + kw] = uint8_t(kh * k + kw); }
      p.M = l.Hout * P; p.Hv = l.Hout; p.Wv = l.Wout;
      return push(p);
    case LT_WIN_C1:
      p.Cin = 64; p.ntaps = k;
      for (int kh = 0; kh < k; ++kh) { p.tap_off[kh] = kh * P; p.tap_w[kh] = uint8_t(kh); }
      p.M = l.Hout * P; p.Hv = l.Hout; p.Wv = l.Wout;
      return push(p);
    case LT_S2:
      p.Cin = l.Cin; p.ntaps = 9;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_off[kh * 3 + kw] = ((kh & 1) * 2 + (kw & 1)) * l.xmap.plane_pix + (kh >> 1) * P + (kw >> 1);
          p.tap_w[kh * 3 + kw] = uint8_t(kh * 3 + kw);
        }
      p.M = l.Hout * P; p.Hv = l.Hout; p.Wv = l.Wout;
      return push(p);
    case LT_WIN_H0:
      p.Cin = 64; p.ntaps = 6;
      for (int kh = 0; kh < 3; ++kh)
        for (int bp = 0; bp < 2; ++bp) {
          p.tap_off[kh * 2 + bp] = ((kh & 1) * 2 + bp) * l.xmap.plane_pix + (kh >> 1) * P;
          p.tap_w[kh * 2 + bp] = uint8_t(kh * 2 + bp);
        }
      p.M = l.Hout * P; p.Hv = l.Hout; p.Wv = l.Wout;
      return push(p);
    case LT_DECONV:
      // out[2i+a, 2j+b] = sum_{kh in K(a), kw in K(b)} x[i - (kh==2), j - (kw==2)] * w[kh, kw]
      p.Cin = l.Cin; p.M = l.Hin * P; p.Hv = l.Hin; p.Wv = l.Win; p.o_scale = 2;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          ConvGemmParams q = p;
          q.o_a = a; q.o_b = b; q.ntaps = 0;
          for (int kh = a; kh < 3; kh += 2)
            for (int kw = b; kw < 3; kw += 2) {
              q.tap_off[q.ntaps] = (l.xmap.pt - (kh == 2)) * P - (kw == 2);
              q.tap_w[q.ntaps] = uint8_t(kh * 3 + kw);
              ++q.ntaps;
            }
          int r = push(q);
          if (r) return r;
        }
      return 0;
  }
  return SGGAN_E_INVALID;
}

// dgrad launches: dX (plain, l.dxH x l.dxW x Cin) from the dY frames of images [b0, b0 + nimg).
int layer_prepare_dgrad(Layer& l, int b0, int nimg, bool dry, const ConvGemmParams* nr = nullptr) {
  l.dgrad.clear();
  if (l.type == LT_WIN_C1) return 0;
  const int k = l.k, P = l.P;
  ConvGemmParams p;
  conv_common(p, l);
  if (nr != nullptr && l.type == LT_S1) {  // fold the norm-backward sums of the layer below into this launch's epilogue
    p.nr_Y = nr->nr_Y; p.nr_stats = nr->nr_stats; p.nr_gamma = nr->nr_gamma; p.nr_beta = nr->nr_beta; p.nr_part = nr->nr_part;
    p.nr_eps = nr->nr_eps; p.nr_gneg = nr->nr_gneg; p.nr_H = nr->nr_H; p.nr_W = nr->nr_W; p.nr_pad = nr->nr_pad;
  }
  const int Cdy = l.dymap.C;
  p.A = l.dY + int64_t(b0) * l.dymap.frame_pix * Cdy;
  p.a_frame_pix = l.dymap.frame_pix; p.a_row_stride = Cdy; p.B = nimg;
  p.Wt = l.Wd; p.wt_taps = l.packd.T; p.CoutPad = l.CinN; p.Cout = l.Cin; p.BN = bn_for(l.CinN);
  p.P = P; p.out = l.dX; p.out_f32 = l.dx_f32;
  p.omap = plain_map(l.dxH, l.dxW, l.Cin);
  auto push = [&](const ConvGemmParams& q) -> int {
    ConvGemmLaunch L;
    if (dry) { L.p = q; l.dgrad.push_back(L); return 0; }
    int r = prepare_conv_gemm(q, &L);
    if (r) return r;
    l.dgrad.push_back(L);
    return 0;
  };
  switch (l.type) {
    case LT_S1:
      // dx_p[u, v] = sum_{kh,kw} dy[u - kh, v - kw] * w[kh,kw]^T on the padded grid (pitch P, no slack)
      p.Cin = Cdy; p.ntaps = k * k;
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
          p.tap_off[kh * k + kw] = (k - 1 - kh) * P - kw;
          p.tap_w[kh * k + kw] = uint8_t(kh * k + kw);
        }
      p.M = l.dxH * P; p.Hv = l.dxH; p.Wv = l.dxW;
      return push(p);
    case LT_OUT7:
      // sliding-window K: row r of the dY map = pixels r..r+7 x 8 channels; window slot q <-> kw = 6 - q
      p.a_row_stride = 8; p.Cin = 64; p.ntaps = k;
      for (int kh = 0; kh < k; ++kh) { p.tap_off[kh] = (k - 1 - kh) * P - (k - 1); p.tap_w[kh] = uint8_t(kh); }
      p.M = l.dxH * P; p.Hv = l.dxH; p.Wv = l.dxW;
      return push(p);
    case LT_S2:
    case LT_WIN_H0:
      // input phase (a, b): dx[2i+a, 2j+b] = sum_{kh in K(a), kw in K(b)} dy[i - (kh==2), j - (kw==2)] w^T
      p.Cin = Cdy; p.o_scale = 2;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int Ha = (l.Hin - a + 1) / 2, Wb = (l.Win - b + 1) / 2;
          if (Ha < 1 || Wb < 1) continue;
          ConvGemmParams q = p;
          q.o_a = a; q.o_b = b; q.ntaps = 0; q.M = Ha * P; q.Hv = Ha; q.Wv = Wb;
          for (int kh = a; kh < 3; kh += 2)
            for (int kw = b; kw < 3; kw += 2) {
              q.tap_off[q.ntaps] = (1 - (kh == 2)) * P - (kw == 2);
              q.tap_w[q.ntaps] = uint8_t(kh * 3 + kw);
              ++q.ntaps;
            }
          int r = push(q);
          if (r) return r;
        }
      return 0;
    case LT_DECONV:
      // dx[i, j] = sum_{kh,kw} dOut[2i+kh, 2j+kw] * w[kh,kw]  (dOut phase-split)
      p.Cin = Cdy; p.ntaps = 9;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_off[kh * 3 + kw] = ((kh & 1) * 2 + (kw & 1)) * l.dymap.plane_pix + (kh >> 1) * P + (kw >> 1);
          p.tap_w[kh * 3 + kw] = uint8_t(kh * 3 + kw);
        }
      p.M = l.Hin * P; p.Hv = l.Hin; p.Wv = l.Win;
      return push(p);
    default:
      return 0;
  }
}

// wgrad launches over the first nimg images; dW = gradient of the layer's kernel (Keras layout).
int layer_prepare_wgrad(Layer& l, float* dW, int nimg, bool dry, float* part, size_t part_elems) {
  l.wgrad.clear();
  const int k = l.k, P = l.P;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = nimg;
  int xo[SGGAN_MAX_TAPS], yo[SGGAN_MAX_TAPS];  // per-tap offsets of the activation frame / the dY frame
  int ntaps = 0, Mpix = 0;
  switch (l.type) {
    case LT_S1:
      ntaps = k * k; Mpix = l.Hout * P;
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) { xo[kh * k + kw] = kh * P + kw; yo[kh * k + kw] = (k - 1) * P; }
      break;
    case LT_S2:
      ntaps = 9; Mpix = l.Hout * P;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          xo[kh * 3 + kw] = ((kh & 1) * 2 + (kw & 1)) * l.xmap.plane_pix + (kh >> 1) * P + (kw >> 1);
          yo[kh * 3 + kw] = P;
        }
      break;
    case LT_DECONV:
      ntaps = 9; Mpix = l.Hin * P;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          xo[kh * 3 + kw] = l.xmap.pt * P;
          yo[kh * 3 + kw] = ((kh & 1) * 2 + (kw & 1)) * l.dymap.plane_pix + (kh >> 1) * P + (kw >> 1);
        }
      break;
    default:
      break;
  }
  auto finish = [&](WgradParams& q) -> int {
    q.Mpix = Mpix;
    const int na = (!q.x_pair && q.Cx % 256 == 0) ? 2 : 1;  // as chosen by prepare_wgrad_gemm
    const int tiles = (q.x_pair ? 1 : q.Cx / (128 * na)) * q.ntaps * (q.Cy / q.BN);
    const int chunks = nimg * ((Mpix + 63) / 64);
    // one CTA per SM holds a whole accumulator: fill exactly one wave of 148 SMs (no tail round)
    int ks = tiles >= 148 ? 1 : 148 / tiles;
    const int pg = wgrad_pair_groups(q, nullptr, nullptr);
    if (pg > 0) {  // CTA-pair kernel: one cluster per TPC, 74 clusters in a wave
      const int clusters = pg * (q.Cx / 256) * (q.Cy / 256);
      ks = clusters >= 74 ? 1 : 74 / clusters;
    }
    if (ks > chunks / 8) ks = chunks / 8;
    if (ks < 1) ks = 1;
    q.ksplit = ks;
    if (part != nullptr) {  // (the sliding-window layers too: their few KB of partials instead of fp32 atomics)
      const int64_t numel = int64_t(q.ntaps) * q.dw_tap_stride;
      if (size_t(numel) * size_t(ks) <= part_elems && (numel & 3) == 0) { q.part = part; q.part_stride = numel; }
    }
    WgradLaunch L;
    if (dry) { L.p = q; l.wgrad.push_back(L); return 0; }
    int r = prepare_wgrad_gemm(q, &L);
    if (r) return r;
    l.wgrad.push_back(L);
    return 0;
  };
  if (l.type == LT_S1 || l.type == LT_S2 || l.type == LT_DECONV) {
    const bool deconv = l.type == LT_DECONV;
    const bool x_is_act = (l.Cin % 128 == 0);  // which operand takes the 128-row M role
    if (!x_is_act && l.CoutK % 128 != 0) return SGGAN_E_INVALID;
    p.ntaps = ntaps;
    p.dW = dW; p.dw_tap_stride = int64_t(l.Cin) * l.Cout;
    // element strides of dW for (input channel ci, output channel co)
    const int64_t s_ci = deconv ? 1 : l.Cout, s_co = deconv ? l.Cin : 1;
    if (x_is_act) {
      p.X = l.X; p.x_frame_pix = l.xmap.frame_pix; p.x_row_stride = l.xmap.C; p.Cx = l.Cin;
      p.Y = l.dY; p.y_frame_pix = l.dymap.frame_pix; p.y_row_stride = l.dymap.C; p.Cy = l.CoutK;
      for (int t = 0; t < ntaps; ++t) { p.x_off[t] = xo[t]; p.y_off[t] = yo[t]; }
      p.nx_valid = l.Cin; p.ny_valid = l.Cout; p.dw_sx = s_ci; p.dw_sy = s_co;
    } else {
      p.X = l.dY; p.x_frame_pix = l.dymap.frame_pix; p.x_row_stride = l.dymap.C; p.Cx = l.CoutK;
      p.Y = l.X; p.y_frame_pix = l.xmap.frame_pix; p.y_row_stride = l.xmap.C; p.Cy = l.Cin;
      for (int t = 0; t < ntaps; ++t) { p.x_off[t] = yo[t]; p.y_off[t] = xo[t]; }
      p.nx_valid = l.Cout; p.ny_valid = l.Cin; p.dw_sx = s_co; p.dw_sy = s_ci;
    }
    p.BN = bn_for(p.Cy);
    return finish(p);
  }
  // sliding-window layers: pairs of kernel rows / tap groups fill the 128-row M tile; result goes to
  // wscratch [pairs][128][64] and is unpacked by launch_unpack_wgrad.
  p.x_pair = 1; p.Cx = 64; p.Cy = 64; p.BN = 64; p.nx_valid = 128; p.ny_valid = 64;
  p.dW = l.wscratch; p.dw_tap_stride = 128 * 64; p.dw_sx = 64; p.dw_sy = 1;
  Mpix = l.Hout * P;
  if (l.type == LT_WIN_C1) {
    p.X = l.X; p.x_frame_pix = l.xmap.frame_pix; p.x_row_stride = 8;
    p.Y = l.dY; p.y_frame_pix = l.dymap.frame_pix; p.y_row_stride = l.dymap.C;
    p.ntaps = (k + 1) / 2;
    for (int t = 0; t < p.ntaps; ++t) {
      const int kh0 = 2 * t, kh1 = (2 * t + 1 < k) ? 2 * t + 1 : 2 * t;
      p.x_off[t] = kh0 * P; p.x_off2[t] = kh1 * P; p.y_off[t] = (k - 1) * P;
    }
    return finish(p);
  }
  if (l.type == LT_OUT7) {
    p.X = l.X; p.x_frame_pix = l.xmap.frame_pix; p.x_row_stride = l.xmap.C;
    p.Y = l.dY; p.y_frame_pix = l.dymap.frame_pix; p.y_row_stride = 8;
    p.ntaps = (k + 1) / 2;
    for (int t = 0; t < p.ntaps; ++t) {
      const int kh0 = 2 * t, kh1 = (2 * t + 1 < k) ? 2 * t + 1 : 2 * t;
      p.x_off[t] = kh0 * P; p.x_off2[t] = kh1 * P; p.y_off[t] = (k - 1) * P - (k - 1);
    }
    return finish(p);
  }
  if (l.type == LT_WIN_H0) {
    p.X = l.X; p.x_frame_pix = l.xmap.frame_pix; p.x_row_stride = 8;
    p.Y = l.dY; p.y_frame_pix = l.dymap.frame_pix; p.y_row_stride = l.dymap.C;
    p.ntaps = 3;
    auto goff = [&](int g) { const int kh = g >> 1, bp = g & 1; return ((kh & 1) * 2 + bp) * l.xmap.plane_pix + (kh >> 1) * P; };
    for (int t = 0; t < 3; ++t) { p.x_off[t] = goff(2 * t); p.x_off2[t] = goff(2 * t + 1); p.y_off[t] = P; }
    return finish(p);
  }
  return SGGAN_E_INVALID;
}

// ---------------------------------------------------------------------------------------------
// nets
static void add_tensor(Net& n, int rank, int64_t a, int64_t b = 1, int64_t c = 1, int64_t d = 1) {
  TensorInfo t;
  t.rank = rank;
  t.shape[0] = a; t.shape[1] = b; t.shape[2] = c; t.shape[3] = d;
  t.numel = a * b * c * d;
  t.offset = n.nparams;
  n.nparams += t.numel;
  n.T.push_back(t);
}
static Layer make_layer(Net& n, LayerType type, int k, PadMode pad, int Cin, int Cout, int Hin, int Win, bool norm, int act,
                        float alpha, int nb, int nbv) {
  Layer l;
  l.type = type; l.k = k; l.pad = pad; l.Cin = Cin; l.Cout = Cout; l.Hin = Hin; l.Win = Win;
  l.has_norm = norm; l.act = act; l.alpha = alpha; l.nb = nb; l.nbv = nbv;
  l.X = l.Y = l.dY = nullptr; l.dX = nullptr; l.Yf32 = nullptr; l.stats = l.bsums = l.stats_part = nullptr; l.bsync = nullptr; l.Wf = l.Wd = nullptr;
  l.wscratch = nullptr;
  l.ti_w = int(n.T.size());
  if (type == LT_DECONV) add_tensor(n, 4, k, k, Cout, Cin); else add_tensor(n, 4, k, k, Cin, Cout);
  l.ti_b = int(n.T.size());
  add_tensor(n, 1, Cout);
  l.ti_g = l.ti_be = -1;
  if (norm) {
    l.ti_g = int(n.T.size()); add_tensor(n, 1, Cout);
    l.ti_be = int(n.T.size()); add_tensor(n, 1, Cout);
  }
  return l;
}

int Engine::build_net_g() {
  const int B = cfg.batch, H = cfg.image_height, W = cfg.image_width, g = cfg.gf_dim;
  G.nparams = 0;
  auto add = [&](Layer l) -> int { int r = layer_geometry(l); if (r) return r; G.L.push_back(l); return 0; };
  int r;
  if ((r = add(make_layer(G, LT_WIN_C1, 7, PAD_REFLECT, 3, g, H, W, true, SG_ACT_RELU, 0.f, B, B)))) return r;
  if ((r = add(make_layer(G, LT_S2, 3, PAD_ZERO, g, 2 * g, H, W, true, SG_ACT_RELU, 0.f, B, B)))) return r;
  if ((r = add(make_layer(G, LT_S2, 3, PAD_ZERO, 2 * g, 4 * g, H / 2, W / 2, true, SG_ACT_RELU, 0.f, B, B)))) return r;
  for (int i = 0; i < cfg.n_blocks; ++i) {
    if ((r = add(make_layer(G, LT_S1, 3, PAD_REFLECT, 4 * g, 4 * g, H / 4, W / 4, true, SG_ACT_RELU, 0.f, B, B)))) return r;
    if ((r = add(make_layer(G, LT_S1, 3, PAD_REFLECT, 4 * g, 4 * g, H / 4, W / 4, true, SG_ACT_NONE, 0.f, B, B)))) return r;
  }
  if ((r = add(make_layer(G, LT_DECONV, 3, PAD_ZERO, 4 * g, 2 * g, H / 4, W / 4, true, SG_ACT_RELU, 0.f, B, B)))) return r;
  if ((r = add(make_layer(G, LT_DECONV, 3, PAD_ZERO, 2 * g, g, H / 2, W / 2, true, SG_ACT_RELU, 0.f, B, B)))) return r;
  if ((r = add(make_layer(G, LT_OUT7, 7, PAD_REFLECT, g, 3, H, W, false, SG_ACT_TANH, 0.f, B, B)))) return r;
  return 0;
}

int Engine::build_net_d() {
  const int B = cfg.batch, H = cfg.image_height, W = cfg.image_width, d = cfg.df_dim;
  D.nparams = 0;
  const float a = 0.3f;  // tf.keras.layers.LeakyReLU() default
  int h = H, w = W;
  auto add = [&](Layer l) -> int {
    int r = layer_geometry(l);
    if (r) return r;
    h = l.Hout; w = l.Wout;
    D.L.push_back(l);
    return 0;
  };
  int r;
  if ((r = add(make_layer(D, LT_WIN_H0, 3, PAD_ZERO, 3, d, h, w, false, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S2, 3, PAD_ZERO, d, 2 * d, h, w, true, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S2, 3, PAD_ZERO, 2 * d, 4 * d, h, w, true, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S1, 3, PAD_ZERO, 4 * d, 8 * d, h, w, true, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S2, 3, PAD_VALID, 8 * d, 8 * d, h, w, true, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S2, 3, PAD_VALID, 8 * d, 8 * d, h, w, true, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S1, 3, PAD_VALID, 8 * d, 8 * d, h, w, true, SG_ACT_LRELU, a, 2 * B, 3 * B)))) return r;
  if ((r = add(make_layer(D, LT_S1, 3, PAD_ZERO, 8 * d, cfg.segment_class, h, w, false, SG_ACT_NONE, 0.f, 2 * B, 3 * B)))) return r;
  Hd = h; Wd = w;
  return 0;
}

static const size_t kSlack = 4096;  // the sliding-window maps read up to 7 pixels past a frame

void Engine::alloc_and_prepare(Net& n, Arena& a, bool zero_part) {
  if (zero_part) {
    n.g = (float*)a.take(size_t(n.nparams) * 4);
    for (auto& l : n.L) {
      l.stats = l.has_norm ? (float*)a.take(size_t(l.nb) * l.Cout * 2 * 4) : nullptr;
      l.stats_part = l.has_norm ? (float*)a.take(size_t(l.nb) * l.stats_T_max * l.Cout * 2 * 4) : nullptr;
      l.bsums = l.has_norm ? (float*)a.take(size_t(l.nbv) * l.Cout * 2 * 4) : nullptr;
      l.bsync = l.has_norm ? (int*)a.take(size_t(l.nbv) * sizeof(int)) : nullptr;  // arrival counters of the fused norm backward
      l.wscratch = l.wscratch_elems ? (float*)a.take(size_t(l.wscratch_elems) * 4) : nullptr;
    }
    return;
  }
  n.p = (float*)a.take(size_t(n.nparams) * 4);
  n.m = (float*)a.take(size_t(n.nparams) * 4);
  n.v = (float*)a.take(size_t(n.nparams) * 4);
  for (size_t i = 0; i < n.L.size(); ++i) {
    Layer& l = n.L[i];
    l.Wf = (sg_bf16*)a.take(size_t(l.packf.T) * l.packf.N * l.packf.K * 2);
    l.Wd = l.packd.T ? (sg_bf16*)a.take(size_t(l.packd.T) * l.packd.N * l.packd.K * 2) : nullptr;
    l.X = (sg_bf16*)a.take(size_t(l.nb) * l.xmap.frame_pix * l.xmap.C * 2 + kSlack);
    l.dY = (sg_bf16*)a.take(size_t(l.nbv) * l.dymap.frame_pix * l.dymap.C * 2 + kSlack);
    const bool raw_y = l.has_norm;  // no-norm layers write straight into their consumer
    l.Y = raw_y ? (sg_bf16*)a.take(size_t(l.nb) * l.Hout * l.Wout * l.Cout * 2) : nullptr;
    const bool need_dx = !(l.type == LT_WIN_C1);
    const int ndx = l.type == LT_WIN_H0 ? (l.nbv - l.nb) : l.nbv;
    l.dX = need_dx ? a.take(size_t(ndx) * l.dxH * l.dxW * l.Cin * (l.dx_f32 ? 4 : 2)) : nullptr;
  }
}

int Engine::prepare_layer(Net& n, int li) {
  Layer& l = n.L[li];
  const bool isG = (&n == &G);
  const float* bias = n.p ? n.p + n.T[l.ti_b].offset : nullptr;
  int r;
  if (l.has_norm) {
    r = layer_prepare_fwd(l, bias, l.Y, plain_map(l.Hout, l.Wout, l.Cout), 0, dry);
  } else if (l.type == LT_OUT7) {
    r = layer_prepare_fwd(l, bias, fake, plain_map(l.Hout, l.Wout, 3), 1, dry);
  } else if (l.type == LT_WIN_H0) {
    r = layer_prepare_fwd(l, bias, n.L[li + 1].X, n.L[li + 1].xmap, 0, dry);
  } else {  // D h4
    r = layer_prepare_fwd(l, bias, h4, plain_map(l.Hout, l.Wout, l.Cout), 1, dry);
  }
  if (r) return r;
  if (l.type == LT_WIN_H0) r = layer_prepare_dgrad(l, l.nb, l.nbv - l.nb, dry);
  else if (isG && li == 0) r = 0;
  else {
    // second convolution of a residual block: its dgrad produces the gradient w.r.t. pad(relu(norm(Y of the first
    // convolution))) -- let that launch's epilogue accumulate the norm-backward sums of the first convolution's norm
    // (SGGAN_FOLD_INRED=0: separate reduce pass)
    static const bool fold_on = []() { const char* e = getenv("SGGAN_FOLD_INRED"); return !(e && e[0] == '0'); }();
    const bool conv_b = isG && li >= 4 && li <= 3 + 2 * cfg.n_blocks - 1 && ((li - 3) & 1) == 1;
    ConvGemmParams nr;
    memset(&nr, 0, sizeof(nr));
    Layer* below = conv_b ? &n.L[li - 1] : nullptr;
    if (conv_b && fold_on && !dry && below->nr_part != nullptr && l.pad == PAD_REFLECT) {
      nr.nr_Y = below->Y; nr.nr_stats = below->stats;
      nr.nr_gamma = n.p + n.T[below->ti_g].offset; nr.nr_beta = n.p + n.T[below->ti_be].offset;
      nr.nr_part = below->nr_part; nr.nr_eps = cfg.in_eps;
      nr.nr_gneg = below->act == SG_ACT_RELU ? 0.f : (below->act == SG_ACT_LRELU ? below->alpha : 1.f);
      nr.nr_H = below->Hout; nr.nr_W = below->Wout; nr.nr_pad = (l.k - 1) / 2;
    }
    r = layer_prepare_dgrad(l, 0, l.nbv, dry, nr.nr_Y ? &nr : nullptr);
    if (below != nullptr) below->nr_T = (!dry && r == 0 && l.dgrad.size() == 1 && l.dgrad[0].nr_ok) ? l.dgrad[0].T128 : 0;
  }
  if (r) return r;
  float* dW = n.g ? n.g + n.T[l.ti_w].offset : nullptr;
  return layer_prepare_wgrad(l, dW, l.nb, dry, wg_part, wg_part_elems);
}

int Engine::build(const sggan_config& c, void* ws, size_t ws_bytes, cudaStream_t stream, bool dry_run, size_t* need) {
  cfg = c; st = stream; dry = dry_run; step = 0; nlaunch = 0; weights_ready = false;
  if (!dry_run && st2 == nullptr) {
    const char* env = getenv("SGGAN_SIDE_STREAM");
    if (!(env && env[0] == '0')) {
      if (cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
        err = "side stream creation failed";
        return SGGAN_E_CUDA;
      }
    }
  }
  if (c.gf_dim != 64 || c.df_dim != 64) { err = "gf_dim and df_dim must be 64 (module.py:221,274)"; return SGGAN_E_INVALID; }
  if (c.batch < 1 || c.image_height % 4 || c.image_width % 4 || c.n_blocks < 1 || c.segment_class < 1 ||
      c.segment_class > 64) { err = "unsupported batch / image size / class count"; return SGGAN_E_INVALID; }
  int r;
  if ((r = build_net_g()) || (r = build_net_d())) { err = "image size too small for the layer stack"; return r; }
  Ho = Hd > c.mask_height ? Hd : c.mask_height;
  Wo = Wd > c.mask_width ? Wd : c.mask_width;
  if ((Hd != Ho && Hd != 1) || (c.mask_height != Ho && c.mask_height != 1) || (Wd != Wo && Wd != 1) ||
      (c.mask_width != Wo && c.mask_width != 1)) {
    err = "mask grid does not broadcast against the discriminator logit grid (SURVEY D4)";
    return SGGAN_E_INVALID;
  }
  Arena a;
  a.base = dry ? nullptr : (uint8_t*)ws;
  a.off = 0;
  const int B = c.batch, H = c.image_height, W = c.image_width;
  // ---- zeroed every step
  a.take(0);
  const size_t z0 = (a.off + 255) & ~size_t(255);
  alloc_and_prepare(G, a, true);
  alloc_and_prepare(D, a, true);
  loss = (float*)a.take(8 * 4);
  const size_t z1 = a.off;
  // ---- persistent
  alloc_and_prepare(G, a, false);
  alloc_and_prepare(D, a, false);
  fake = (float*)a.take(size_t(B) * H * W * 3 * 4);
  h4 = (float*)a.take(size_t(2 * B) * Hd * Wd * c.segment_class * 4);
  logits = (float*)a.take(size_t(2 * B) * Ho * Wo * 4);
  dD = (float*)D.L[0].dX;
  edge_w = (float*)a.take(size_t(B) * H * W * 4);
  dGl = (float*)a.take(size_t(B) * H * W * 3 * 4);
  const Layer& rb = G.L[3];
  for (int i = 0; i < 2; ++i) resG[i] = (sg_bf16*)a.take(size_t(B) * rb.Hin * rb.Win * rb.Cin * 2);
  wg_part_elems = size_t(160) * 256 * 256;  // one (2 x 128) x 256 fp32 tile per CTA of a single-wave split-K launch
  wg_part = (float*)a.take(wg_part_elems * 4);
  step_dev = (long long*)a.take(sizeof(long long));
  in_part = (float*)a.take(in_bwd_partials_bytes(512));  // per-block partial sums of the norm-backward reduce pass
  for (int li = 3; li + 1 <= 3 + 2 * c.n_blocks - 1; li += 2) {  // first convolution of every residual block (see prepare_layer)
    Layer& la = G.L[li];
    const Layer& lb = G.L[li + 1];
    const int T = (lb.dxH * lb.P + 127) / 128;
    la.nr_part = (float*)a.take(size_t(la.nbv) * T * la.Cout * 2 * 4);
  }
  {
    const Layer& h0 = D.L[0];
    red_scratch = (float*)a.take(ordered_sum_scratch_floats(B, H, W, c.segment_class, h0.Cout, h0.Hout, h0.Wout, h0.nbv) * 4);
    red_ticket = (unsigned int*)a.take(16 * sizeof(unsigned int));
  }
  for (int i = 0; i < 2; ++i) {
    pack_jobs[i] = (PackParams*)a.take(kMaxPackJobs * sizeof(PackParams));
    pack_starts[i] = (int*)a.take((kMaxPackJobs + 1) * sizeof(int));
  }
  if (need) *need = a.off + 256;
  if (dry) return 0;
  if (a.off > ws_bytes) { err = "workspace too small"; return SGGAN_E_WORKSPACE; }
  zero_begin = (uint8_t*)ws + z0;
  zero_end = (uint8_t*)ws + z1;
  if (cudaMemsetAsync(ws, 0, a.off, st) != cudaSuccess) { err = "cudaMemset of the workspace failed"; return SGGAN_E_CUDA; }
  for (size_t i = 0; i < G.L.size(); ++i)
    if ((r = prepare_layer(G, int(i)))) { err = "generator layer " + std::to_string(i) + " prepare failed: " + std::to_string(r); return SGGAN_E_CUDA; }
  if ((r = upload_pack_jobs(SGGAN_NET_G)) || (r = upload_pack_jobs(SGGAN_NET_D))) { err = "weight-pack job table upload failed"; return r; }
  for (size_t i = 0; i < D.L.size(); ++i)
    if ((r = prepare_layer(D, int(i)))) { err = "discriminator layer " + std::to_string(i) + " prepare failed: " + std::to_string(r); return SGGAN_E_CUDA; }
  return 0;
}

// ---------------------------------------------------------------------------------------------
int Engine::pack_weights(int net, cudaStream_t s) {
  // one launch per net: the job table (one PackParams per weight slab) was uploaded by build()
  launch_pack_weights_batch(pack_jobs[net], pack_starts[net], pack_njobs[net], pack_blocks[net], s ? s : st);
  ++nlaunch;
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

int Engine::upload_pack_jobs(int net) {
  Net& n = net == SGGAN_NET_G ? G : D;
  std::vector<PackParams> jobs;
  std::vector<int> starts;
  int blocks = 0;
  auto add = [&](PackParams q) {
    starts.push_back(blocks);
    blocks += int((((int64_t(q.T) * q.N * q.K) >> 3) + 255) / 256);  // 8 elements per thread
    jobs.push_back(q);
  };
  for (auto& l : n.L) {
    PackParams pf = l.packf;
    pf.src = n.p + n.T[l.ti_w].offset; pf.dst = l.Wf;
    add(pf);
    if (l.packd.T) {
      PackParams pd = l.packd;
      pd.src = n.p + n.T[l.ti_w].offset; pd.dst = l.Wd;
      add(pd);
    }
  }
  starts.push_back(blocks);
  if (jobs.size() > kMaxPackJobs) return SGGAN_E_INVALID;
  pack_njobs[net] = int(jobs.size());
  pack_blocks[net] = blocks;
  if (cudaMemcpyAsync(pack_jobs[net], jobs.data(), jobs.size() * sizeof(PackParams), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(pack_starts[net], starts.data(), starts.size() * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)  // the host vectors go out of scope
    return SGGAN_E_CUDA;
  return 0;
}

int Engine::run_conv_list(const std::vector<ConvGemmLaunch>& v) {
  for (const auto& L : v) {
    int r = run_conv_gemm(L, st);
    ++nlaunch;
    if (r) { err = "conv launch failed " + std::to_string(r); return SGGAN_E_CUDA; }
  }
  return 0;
}

int Engine::run_wgrad(Layer& l, Net& n) {
  cudaStream_t ws = st;
  if (st2 != nullptr) {  // fork: everything the weight gradient reads has been enqueued on `st` by now
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(st2, ev_fork, 0);
    ws = st2;
    side_used = true;
  }
  if (l.has_norm) {  // dgamma / dbeta from the sums the norm-backward apply pass of this layer just published
    launch_in_param_grad(l.bsums, l.nb, l.Cout, n.g + n.T[l.ti_g].offset, n.g + n.T[l.ti_be].offset, ws);
    ++nlaunch;
  }
  for (const auto& L : l.wgrad) {
    int r = run_wgrad_gemm(L, ws);
    ++nlaunch;
    if (r) { err = "wgrad launch failed " + std::to_string(r); return SGGAN_E_CUDA; }
    if (L.p.part != nullptr) { launch_wgrad_reduce(L, int64_t(L.p.ntaps) * L.p.dw_tap_stride, ws); ++nlaunch; }
  }
  if (l.unpack_mode >= 0) {
    launch_unpack_wgrad(l.wscratch, n.g + n.T[l.ti_w].offset, l.unpack_mode, l.k, l.k, l.Cin, l.Cout, 64, ws);
    ++nlaunch;
  }
  return 0;
}

void Engine::join_side() {
  if (!side_used) return;
  cudaEventRecord(ev_join, st2);
  cudaStreamWaitEvent(st, ev_join, 0);
  side_used = false;
}

void Engine::in_apply(Net& n, int li, sg_bf16* dst, const FrameMap& dmap, const sg_bf16* res, const FrameMap* rmap) {
  Layer& l = n.L[li];
  InApplyParams p;
  memset(&p, 0, sizeof(p));
  p.Y = l.Y; p.B = l.nb; p.H = l.Hout; p.W = l.Wout; p.C = l.Cout;
  p.stats = l.stats; p.gamma = n.p + n.T[l.ti_g].offset; p.beta = n.p + n.T[l.ti_be].offset;
  p.stats_part = l.stats_part; p.stats_T = l.stats_T; p.stats_out = l.stats;  // finalize fused into the apply pass
  p.eps = cfg.in_eps; p.act = l.act; p.act_alpha = l.alpha;
  p.res = res;
  if (rmap) p.rmap = *rmap;
  p.dst = dst; p.dmap = dmap;
  if (launch_in_apply(p, st) < 0 && glue_err == 0) glue_err = 1;
  ++nlaunch;
}

GradSrc Engine::dx_src(const Layer& l) const {
  GradSrc g;
  g.ptr = l.dX; g.f32 = l.dx_f32; g.Hs = l.dxH; g.Ws = l.dxW; g.oy = l.dx_oy; g.ox = l.dx_ox; g.fold = l.dx_fold;
  return g;
}
static GradSrc plain_src(const void* ptr, int H, int W) {
  GradSrc g;
  g.ptr = ptr; g.f32 = 0; g.Hs = H; g.Ws = W; g.oy = 0; g.ox = 0; g.fold = 0;
  return g;
}
static GradSrc no_src() {
  GradSrc g;
  memset(&g, 0, sizeof(g));
  return g;
}

// instance-norm backward of layer li: gradient sources g1 (+ g2) w.r.t. its post-activation output
void Engine::in_bwd(Net& n, int li, const GradSrc& g1, const GradSrc& g2, int nb_act, int act_wrap, int nb_param,
                    sg_bf16* gather_dst, int gH, int gW) {
  Layer& l = n.L[li];
  InBwdParams p;
  memset(&p, 0, sizeof(p));
  p.Y = l.Y; p.B = l.nbv; p.H = l.Hout; p.W = l.Wout; p.C = l.Cout;
  p.nb_act = nb_act; p.act_wrap = act_wrap;
  p.stats = l.stats; p.gamma = n.p + n.T[l.ti_g].offset; p.beta = n.p + n.T[l.ti_be].offset;
  p.eps = cfg.in_eps; p.act = l.act; p.act_alpha = l.alpha;
  p.g1 = g1; p.g2 = g2; p.sums = l.bsums; p.sums_part = in_part; p.dst = l.dY; p.dmap = l.dymap;
  (void)nb_param;  // dgamma / dbeta are taken from l.bsums by run_wgrad (side stream)
  p.gather_dst = gather_dst;  // reduce pass also materialises g1 + g2 (the residual-stream gradient) ...
  // Opt-in (SGGAN_FUSE_INBWD=1): both passes in ONE launch, the blocks of an image meeting on l.bsync.  Measured on
  // B200 it is correct but not faster than the two launches (55 us vs 52 us at the residual blocks): every block waits
  // for the slowest block of its image at the barrier, and pass 2 starts with an empty pipeline because all stages
  // still hold pass-1 data (DESIGN.md, row streams).
  static const bool fused = []() { const char* e = getenv("SGGAN_FUSE_INBWD"); return e && e[0] == '1'; }();
  if (fused) {
    p.sync_ctr = l.bsync;
    const int r = launch_in_bwd_fused(p, st);
    if (r < 0 && glue_err == 0) glue_err = 5;
    ++nlaunch;
    return;
  }
  if (l.nr_T > 0 && gather_dst == nullptr && g2.ptr == nullptr) {
    // the sums were accumulated per tile by the dgrad launch that produced g1 (its epilogue): apply pass only
    p.sums_part = l.nr_part;
    p.sums_nblk = l.nr_T;
    if (launch_in_bwd_apply(p, st) < 0 && glue_err == 0) glue_err = 3;
    ++nlaunch;
    return;
  }
  const int nblk = launch_in_bwd_reduce(p, st);
  if (nblk <= 0) { if (glue_err == 0) glue_err = 2; return; }
  if (gather_dst != nullptr) {  // ... which is then the single source of the apply pass
    p.g1.ptr = gather_dst; p.g1.f32 = 0; p.g1.Hs = gH; p.g1.Ws = gW; p.g1.oy = 0; p.g1.ox = 0; p.g1.fold = 0;
    memset(&p.g2, 0, sizeof(p.g2));
  }
  p.sums_nblk = nblk;
  p.gather_dst = nullptr;
  if (launch_in_bwd_apply(p, st) < 0 && glue_err == 0) glue_err = 3;  // also publishes the reduced sums in l.bsums
  nlaunch += 2;
}

// ---------------------------------------------------------------------------------------------
int Engine::gen_forward(const float* real_A, float* fake_out) {
  join_side();  // an optimizer update issued on the side stream (sggan_step_adam_async) must have landed
  if (!weights_ready) { err = "weights not set (call sggan_weights_changed)"; return SGGAN_E_STATE; }
  const int B = cfg.batch, H = cfg.image_height, W = cfg.image_width;
  int r;
  launch_prep_image3(real_A, B, H, W, G.L[0].X, G.L[0].xmap, 0, st); ++nlaunch;
  const int nl = int(G.L.size());
  for (int li = 0; li < nl; ++li) {
    Layer& l = G.L[li];
    const bool in_block = (li >= 3 && li < 3 + 2 * cfg.n_blocks);
    const bool block_b = in_block && ((li - 3) & 1) == 1;
    const bool timed = prof_on && prof_kind == 0 && in_block && prof_used + 2 <= prof_ev.size();
    if (timed) cudaEventRecord(prof_ev[prof_used++], st);
    if ((r = run_conv_list(l.fwd))) return r;
    if (timed) cudaEventRecord(prof_ev[prof_used++], st);
    if (!l.has_norm) continue;
    Layer& nx = G.L[li + 1];
    const bool timed_a = prof_on && prof_kind == 1 && in_block && !block_b && prof_used + 2 <= prof_ev.size();
    if (timed_a) cudaEventRecord(prof_ev[prof_used++], st);
    if (block_b) in_apply(G, li, nx.X, nx.xmap, G.L[li - 1].X, &G.L[li - 1].xmap);  // y + x (module.py:217)
    else in_apply(G, li, nx.X, nx.xmap, nullptr, nullptr);
    if (timed_a) cudaEventRecord(prof_ev[prof_used++], st);
  }
  if (fake_out != nullptr && fake_out != fake)
    if (cudaMemcpyAsync(fake_out, fake, size_t(B) * H * W * 3 * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return SGGAN_E_CUDA;
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

// D forward over [first | second] (each nimg_each = B images)
int Engine::disc_forward_2b(const float* first, const float* second, int nimg_each) {
  join_side();  // an optimizer update issued on the side stream (sggan_step_adam_async) must have landed
  const int H = cfg.image_height, W = cfg.image_width;
  int r;
  launch_prep_image3(first, nimg_each, H, W, D.L[0].X, D.L[0].xmap, 0, st);
  launch_prep_image3(second, nimg_each, H, W, D.L[0].X, D.L[0].xmap, nimg_each, st);
  nlaunch += 2;
  const int nl = int(D.L.size());
  for (int li = 0; li < nl; ++li) {
    Layer& l = D.L[li];
    if ((r = run_conv_list(l.fwd))) return r;
    if (l.has_norm) {
      in_apply(D, li, D.L[li + 1].X, D.L[li + 1].xmap, nullptr, nullptr);
    }
  }
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

int Engine::disc_forward_user(const float* x, const float* mask, float* logits_out) {
  if (!weights_ready) { err = "weights not set"; return SGGAN_E_STATE; }
  int r = disc_forward_2b(x, x, cfg.batch);
  if (r) return r;
  launch_mask_reduce(h4, mask, cfg.batch, Hd, Wd, cfg.mask_height, cfg.mask_width, cfg.segment_class, logits_out, st);
  ++nlaunch;
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

int Engine::step_fwd_bwd_d(const float* real_A, const float* seg_A, const float* mask, float* losses_out) {
  join_side();  // an optimizer update issued on the side stream (sggan_step_adam_async) must have landed
  if (!weights_ready) { err = "weights not set"; return SGGAN_E_STATE; }
  nlaunch = 0;
  real_A_ = real_A; seg_A_ = seg_A; mask_ = mask; losses_out_ = losses_out;
  const int B = cfg.batch;
  if (cudaMemsetAsync(zero_begin, 0, size_t(zero_end - zero_begin), st) != cudaSuccess) return SGGAN_E_CUDA;
  int r;
  // fake_A = G(real_A);  D([seg_A ; fake_A])   (model.py:176,186-188; the third D forward is a duplicate)
  if ((r = gen_forward(real_A, nullptr))) return r;
  if ((r = disc_forward_2b(seg_A, fake, B))) return r;
  // losses + d logits for the three backward seeds (real->1, fake->0 for D; fake->1 for G)
  Layer& l4 = D.L.back();
  DiscLossParams dl;
  memset(&dl, 0, sizeof(dl));
  dl.h4 = h4; dl.mask = mask; dl.B = B; dl.Hd = Hd; dl.Wd = Wd; dl.hm = cfg.mask_height; dl.wm = cfg.mask_width;
  dl.Cs = cfg.segment_class;
  dl.lsgan = (cfg.loss_mode == SGGAN_LOSS_SGGAN && cfg.use_lsgan) ? 1 : 0;
  dl.disc_scale = cfg.loss_mode == SGGAN_LOSS_SGGAN ? 0.5f : 1.f;
  dl.logits = logits; dl.loss = loss; dl.dst = l4.dY; dl.dmap = l4.dymap;
  dl.dbias = D.g + D.T[l4.ti_b].offset;
  dl.red = OrderedSum{red_scratch, red_ticket + 0};
  launch_disc_loss(dl, st); ++nlaunch;
  // D backward over 3B virtual images; weights see the first 2B (disc_tape), the last B carry the
  // generator's GAN gradient back to fake_A (gen_tape)   (model.py:196-197)
  const int nl = int(D.L.size());
  for (int li = nl - 1; li >= 0; --li) {
    Layer& l = D.L[li];
    if (li < nl - 1) {
      Layer& up = D.L[li + 1];
      if (l.has_norm) {
        in_bwd(D, li, dx_src(up), no_src(), 2 * B, B, 2 * B);
      } else {  // h0: LeakyReLU only, activation lives in h1's input frame
        ActBwdParams ab;
        memset(&ab, 0, sizeof(ab));
        ab.g = dx_src(up); ab.Z = up.X; ab.zmap = up.xmap; ab.B = l.nbv; ab.H = l.Hout; ab.W = l.Wout; ab.C = l.Cout;
        ab.nb_act = 2 * B; ab.act_wrap = B; ab.alpha = l.alpha; ab.dst = l.dY; ab.dmap = l.dymap;
        ab.dbias = D.g + D.T[l.ti_b].offset; ab.nb_bias = 2 * B;
        ab.red = OrderedSum{red_scratch, red_ticket + 1};
        launch_act_bwd(ab, st); ++nlaunch;
      }
    }
    if ((r = run_wgrad(l, D))) return r;
    if ((r = run_conv_list(l.dgrad))) return r;
  }
  join_side();
  if (glue_err) { err = "row-stream launch failed (code " + std::to_string(glue_err) + ")"; glue_err = 0; return SGGAN_E_CUDA; }
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

// The layer at which a data-parallel caller may split the generator backward (sggan_step_backward_g_part): the first
// convolution of the middle residual block.  Once the backward has passed it, the gradients of every tensor from that
// layer's kernel to the end of the flat buffer are final (and contiguous), so their all-reduce can run underneath the
// rest of the backward.
int Engine::bwd_split_layer() const { return 3 + 2 * (cfg.n_blocks / 2); }

// part < 0: the whole backward; part 0: loss seeds .. the split layer (inclusive); part 1: the rest.
int Engine::step_bwd_g(int part) {
  const int B = cfg.batch, H = cfg.image_height, W = cfg.image_width;
  int r;
  const int nl = int(G.L.size());
  const int li_split = bwd_split_layer();
  Layer& lo = G.L[nl - 1];
  static const bool fuse_gather = []() { const char* e = getenv("SGGAN_FUSE_GATHER"); return !(e && e[0] == '0'); }();
  const int first_blk = 3, last_b = 3 + 2 * cfg.n_blocks - 1;
  const Layer& rb = G.L[first_blk];
  if (part == 1) {
    if (!bwd_half_done) { err = "sggan_step_backward_g_part(1) without part 0"; return SGGAN_E_STATE; }
    bwd_half_done = false;
  } else {
  bwd_half_done = false;
  FakeGradParams fg;
  memset(&fg, 0, sizeof(fg));
  fg.fake = fake; fg.dD = dD; fg.B = B; fg.H = H; fg.W = W; fg.loss = loss; fg.dst = lo.dY; fg.dmap = lo.dymap;
  fg.dbias = G.g + G.T[lo.ti_b].offset;
  fg.red = OrderedSum{red_scratch, red_ticket + 2};
  float l1w, lgw = 0.f;
  // profile kind 3: the loss kernels of the generator side (seg-edge weights + gradient-sensitive loss in SG-GAN mode,
  // the L1 / GAN gradient seed, the loss finalize) timed as one group
  const bool timed_l = prof_on && prof_kind == 3 && prof_used + 2 <= prof_ev.size();
  if (timed_l) {
    join_side();  // D's Adam + re-pack run on the side stream right now: keep them out of the timed interval
    cudaEventRecord(prof_ev[prof_used++], st);
  }
  if (cfg.loss_mode == SGGAN_LOSS_P2P) {
    fg.target = seg_A_; l1w = cfg.p2p_lambda;
  } else {
    fg.target = real_A_; l1w = cfg.L1_lambda; lgw = cfg.Lg_lambda;
    if (lgw != 0.f) {
      launch_seg_edge_weight(seg_A_, B, H, W, edge_w, st);
      launch_gradloss(fake, real_A_, edge_w, B, H, W, lgw, loss + 3, dGl, st, OrderedSum{red_scratch, red_ticket + 3});
      nlaunch += 2;
      fg.dG = dGl;
    }
  }
  fg.l1_weight = l1w;
  launch_fake_grad(fg, st); ++nlaunch;
  launch_finalize_losses(loss, l1w, float(B) * H * W * 3.f, lgw, losses_out_, st); ++nlaunch;
  if (timed_l) cudaEventRecord(prof_ev[prof_used++], st);
  // output conv
  if ((r = run_wgrad(lo, G))) return r;
  if ((r = run_conv_list(lo.dgrad))) return r;
  bwd_cur = 0;
  bwd_gres = no_src();
  bwd_add = no_src();
  }
  int& cur = bwd_cur;         // resG ping-pong index holding the gradient w.r.t. the current block output
  GradSrc& gres = bwd_gres;
  GradSrc& add = bwd_add;     // pending G_{k-1} = G_k + fold(dX of conv_a): materialised by the NEXT norm-backward reduce
  const int li_hi = part == 1 ? li_split - 1 : nl - 2, li_lo = part == 0 ? li_split : 0;
  for (int li = li_hi; li >= li_lo; --li) {
    Layer& l = G.L[li];
    Layer& up = G.L[li + 1];
    const bool in_blocks = li >= first_blk && li <= last_b;
    const bool is_b = in_blocks && ((li - first_blk) & 1) == 1;
    if (li == last_b) gres = dx_src(up);                                // grad w.r.t. r_n = dX of the first deconv
    if (is_b || li == first_blk - 1) {  // IN after conv_b (no activation, dz = G_k) / c3 (dz = G_0)
      if (add.ptr != nullptr) {
        in_bwd(G, li, gres, add, l.nb, 0, l.nb, resG[cur], rb.Hin, rb.Win);
        gres = plain_src(resG[cur], rb.Hin, rb.Win);
        cur ^= 1;
        add = no_src();
      } else {
        in_bwd(G, li, gres, no_src(), l.nb, 0, l.nb);
      }
    } else {
      const bool timed_b = prof_on && prof_kind == 2 && in_blocks && prof_used + 2 <= prof_ev.size();
      if (timed_b) {  // time the pass alone: without this, the previous layer's weight gradient (side stream) shares the SMs
        join_side();
        cudaEventRecord(prof_ev[prof_used++], st);
      }
      in_bwd(G, li, dx_src(up), no_src(), l.nb, 0, l.nb);
      if (timed_b) cudaEventRecord(prof_ev[prof_used++], st);
    }
    if ((r = run_wgrad(l, G))) return r;
    if (li == 0) break;
    if ((r = run_conv_list(l.dgrad))) return r;
    if (in_blocks && !is_b) {
      if (fuse_gather) {
        add = dx_src(l);
      } else {  // separate gather launch: G_{k-1} = G_k + fold(dX of conv_a)
        if (launch_grad_gather(gres, dx_src(l), B, rb.Hin, rb.Win, rb.Cin, resG[cur], st) < 0 && glue_err == 0) glue_err = 4;
        ++nlaunch;
        gres = plain_src(resG[cur], rb.Hin, rb.Win);
        cur ^= 1;
      }
    }
  }
  join_side();  // the weight gradients launched so far are final on `st`
  if (part == 0) bwd_half_done = true;
  if (glue_err) { err = "row-stream launch failed (code " + std::to_string(glue_err) + ")"; glue_err = 0; return SGGAN_E_CUDA; }
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

int Engine::step_adam(int net, bool on_side_stream) {
  Net& n = net == SGGAN_NET_G ? G : D;
  cudaStream_t s = st;
  if (on_side_stream && st2 != nullptr) {  // the caller guarantees this net's gradients are final on `st`
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(st2, ev_fork, 0);
    s = st2;
    side_used = true;
  }
  // Keras Adam; the time step comes from the device-side counter (workspace is zero-initialised: step 0)
  launch_adam(n.p, n.g, n.m, n.v, n.nparams, 0.f, cfg.beta1, cfg.beta2, cfg.adam_eps,
              1.f / float(cfg.world_size > 0 ? cfg.world_size : 1), s, step_dev, cfg.lr);
  ++nlaunch;
  return pack_weights(net, s);
}

}  // namespace sggan
