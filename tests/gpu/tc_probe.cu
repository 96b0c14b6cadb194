// tc_probe.cu -- standalone B200 probe for the tcgen05 kernels (no Python, no torch).
// Usage: tc_probe <test>   with test in {conv_small, conv_res, conv_out7, wgrad_small, wgrad_res,
//                                         shift, tmap_overlap}
// Each test checks the kernel against a CPU loop of its own specification (kparams.h) and,
// for the *_res shapes, times it with CUDA events.  Exit code 0 = pass.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../sg-gan-tf2_b200/csrc/conv_gemm_tc.h"
#include "../../sg-gan-tf2_b200/csrc/tc_common.cuh"
#include "../../sg-gan-tf2_b200/csrc/tmap.h"

using namespace sggan;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  return uint16_t(r >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

struct ConvCase {
  int B, H, W, Cin, Cout, CoutPad, BN, k;  // k x k taps, frame padded by k/2, pitch W + k - 1
  int act, out_f32, use_stats;
  int MT;  // 0 = let prepare_conv_gemm choose
  int shift;  // 1: shift-sum form of a k x k convolution with Cout <= 4 (one tap per filter row, N = kw*4 + co)
};

static int run_conv_case(const ConvCase& c, int iters, bool full_check) {
  const int pad = c.k / 2, P = c.W + 2 * pad, rows = c.H + 2 * pad;
  const int64_t frame_pix = int64_t(rows) * P + 8;
  const int ntaps = c.shift ? c.k : c.k * c.k;
  std::vector<uint16_t> hA(size_t(c.B) * frame_pix * c.Cin), hW(size_t(ntaps) * c.CoutPad * c.Cin, 0);
  for (auto& v : hA) v = f2bf(frand());
  for (int t = 0; t < ntaps; ++t)
    for (int n = 0; n < (c.shift ? c.CoutPad : c.Cout); ++n)
      for (int ci = 0; ci < c.Cin; ++ci) hW[(size_t(t) * c.CoutPad + n) * c.Cin + ci] = f2bf(frand() * 0.1f);
  std::vector<float> hbias(c.Cout);
  for (auto& v : hbias) v = frand();

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  uint16_t *dA, *dW;
  void* dOut;
  float *dBias, *dStats;
  const size_t out_elems = size_t(c.B) * c.H * c.W * c.Cout;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dOut, out_elems * (c.out_f32 ? 4 : 2)));
  CK(cudaMalloc(&dBias, c.Cout * 4));
  CK(cudaMalloc(&dStats, size_t(c.B) * c.Cout * 2 * 4));
  const int Tmax = (c.H * (c.W + c.k - 1) + 127) / 128 + 1;
  float* dPart;
  CK(cudaMalloc(&dPart, size_t(c.B) * Tmax * c.Cout * 2 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBias, hbias.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0, out_elems * (c.out_f32 ? 4 : 2)));
  CK(cudaMemset(dStats, 0, size_t(c.B) * c.Cout * 2 * 4));

  p.A = dA;
  p.a_frame_pix = frame_pix;
  p.a_row_stride = c.Cin;
  p.Cin = c.Cin;
  p.B = c.B;
  p.Wt = dW;
  p.wt_taps = ntaps;
  p.ntaps = ntaps;
  p.CoutPad = c.CoutPad;
  p.Cout = c.Cout;
  p.BN = c.BN;
  if (c.shift) {
    p.shift_kw = c.k;
    for (int kh = 0; kh < c.k; ++kh) { p.tap_off[kh] = kh * P; p.tap_w[kh] = uint8_t(kh); }
  } else
  for (int kh = 0; kh < c.k; ++kh)
    for (int kw = 0; kw < c.k; ++kw) {
      p.tap_off[kh * c.k + kw] = kh * P + kw;
      p.tap_w[kh * c.k + kw] = uint8_t(kh * c.k + kw);
    }
  p.M = c.H * P;
  p.P = P;
  p.Hv = c.H;
  p.Wv = c.W;
  p.out = dOut;
  p.out_f32 = c.out_f32;
  p.omap.frame_pix = int64_t(c.H) * c.W;
  p.omap.C = c.Cout;
  p.omap.H = c.H;
  p.omap.W = c.W;
  p.omap.P = c.W;
  p.o_scale = 1;
  p.bias = dBias;
  p.stats = c.use_stats ? dPart : nullptr;
  p.act = c.act;
  p.act_alpha = 0.3f;
  p.MT = c.MT;
  // PROBE_FOLD=1: the launch also accumulates the norm-backward sums of a layer "below" whose raw output is nrY (identity
  // position map: nr_pad = 0), as the engine's dgrad launches do for the residual blocks (ConvGemmParams::nr_*)
  const bool probe_fold = getenv("PROBE_FOLD") && getenv("PROBE_FOLD")[0] == '1' && !c.use_stats && !c.out_f32;
  std::vector<uint16_t> hY;
  std::vector<float> hNs, hG, hBe;
  uint16_t* dNY = nullptr;
  float *dNs = nullptr, *dG = nullptr, *dBe = nullptr, *dNpart = nullptr;
  if (probe_fold) {
    hY.resize(out_elems);
    for (auto& v : hY) v = f2bf(frand() * 2.f);
    hNs.resize(size_t(c.B) * c.Cout * 2);
    hG.resize(c.Cout);
    hBe.resize(c.Cout);
    const float n = float(c.H) * c.W;
    for (int q = 0; q < c.B * c.Cout; ++q) { const float mu = 0.2f * frand(); hNs[2 * q] = mu * n; hNs[2 * q + 1] = (mu * mu + 0.5f + 0.3f * frand()) * n; }
    for (int q = 0; q < c.Cout; ++q) { hG[q] = 1.f + 0.3f * frand(); hBe[q] = 0.3f * frand(); }
    CK(cudaMalloc(&dNY, hY.size() * 2));
    CK(cudaMalloc(&dNs, hNs.size() * 4));
    CK(cudaMalloc(&dG, c.Cout * 4));
    CK(cudaMalloc(&dBe, c.Cout * 4));
    CK(cudaMalloc(&dNpart, size_t(c.B) * Tmax * c.Cout * 2 * 4));
    CK(cudaMemcpy(dNY, hY.data(), hY.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dNs, hNs.data(), hNs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dG, hG.data(), c.Cout * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBe, hBe.data(), c.Cout * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dNpart, 0, size_t(c.B) * Tmax * c.Cout * 2 * 4));
    p.nr_Y = reinterpret_cast<const sg_bf16*>(dNY); p.nr_stats = dNs; p.nr_gamma = dG; p.nr_beta = dBe; p.nr_part = dNpart;
    p.nr_eps = 1e-3f; p.nr_gneg = 0.f; p.nr_H = c.H; p.nr_W = c.W; p.nr_pad = 0;
  }

  ConvGemmLaunch L;
  int r = prepare_conv_gemm(p, &L);
  if (r) {
    printf("prepare_conv_gemm failed %d\n", r);
    return 1;
  }
  L.p.stats_T = L.stat_tiles;
  L.p.stats_t0 = 0;
  if (L.pair) printf("CTA-pair persistent kernel: %d tiles/image, %d pair tiles\n", L.T128, L.npairs);
  if (L.swap) printf("transposed persistent kernel: %d tiles/image, %d pixel + %d weight stages, smem %zu\n", L.T256, L.swap_pstages, L.swap_wstages, L.swap_smem);
  printf("grid %d x %d x %d, MT %d, runs %d, stages A %d B %d, smem %zu, tmem cols %u\n", L.grid_x, L.grid_y, L.grid_z,
         L.p.MT, L.p.nruns, L.sa_stages, L.sb_stages, L.smem, L.tmem_cols);
  r = run_conv_gemm(L, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (r || e != cudaSuccess) {
    printf("launch/sync failed r=%d err=%s watchdog=%d\n", r, cudaGetErrorString(e), 0);
    return 1;
  }
  std::vector<uint8_t> hOut(out_elems * (c.out_f32 ? 4 : 2));
  std::vector<float> hStats(size_t(c.B) * c.Cout * 2, 0.f);
  CK(cudaMemcpy(hOut.data(), dOut, hOut.size(), cudaMemcpyDeviceToHost));
  auto read_stats = [&](std::vector<float>& dst) {
    std::vector<float> part(size_t(c.B) * L.p.stats_T * c.Cout * 2);
    CK(cudaMemcpy(part.data(), dPart, part.size() * 4, cudaMemcpyDeviceToHost));
    std::fill(dst.begin(), dst.end(), 0.f);
    for (int b = 0; b < c.B; ++b)
      for (int t = 0; t < L.p.stats_T; ++t)
        for (int q = 0; q < c.Cout * 2; ++q) dst[size_t(b) * c.Cout * 2 + q] += part[(size_t(b) * L.p.stats_T + t) * c.Cout * 2 + q];
  };
  if (c.use_stats) read_stats(hStats);

  // CPU check (all positions if full_check, else a pseudo-random sample of 4096)
  double max_err = 0, max_ref = 0;
  int bad = 0;
  const int64_t npos = int64_t(c.B) * c.H * c.W;
  const int64_t nsample = full_check ? npos : 4096;
  std::vector<double> s1(size_t(c.B) * c.Cout, 0.0), s2(size_t(c.B) * c.Cout, 0.0);
  for (int64_t sidx = 0; sidx < nsample; ++sidx) {
    int64_t pos = full_check ? sidx : (int64_t)((uint64_t(sidx) * 2654435761ull) % uint64_t(npos));
    const int b = int(pos / (c.H * c.W)), i = int((pos / c.W) % c.H), j = int(pos % c.W);
    for (int n = 0; n < c.Cout; ++n) {
      double acc = hbias[n];
      for (int t = 0; t < c.k * c.k; ++t) {
        const int kh = t / c.k, kw = t % c.k;
        const int64_t pix = int64_t(i) * P + j + kh * P + kw;
        const uint16_t* a = &hA[(size_t(b) * frame_pix + pix) * c.Cin];
        const uint16_t* w = c.shift ? &hW[(size_t(kh) * c.CoutPad + kw * 4 + n) * c.Cin] : &hW[(size_t(t) * c.CoutPad + n) * c.Cin];
        for (int ci = 0; ci < c.Cin; ++ci) acc += double(bf2f(a[ci])) * double(bf2f(w[ci]));
      }
      float ref = float(acc);
      if (c.act == SG_ACT_TANH) ref = tanhf(ref);
      if (c.act == SG_ACT_LRELU) ref = ref > 0 ? ref : 0.3f * ref;
      if (c.act == SG_ACT_RELU) ref = ref > 0 ? ref : 0.f;
      const float rs = c.out_f32 ? ref : bf2f(f2bf(ref));  // statistics are taken over the stored (bf16) values
      s1[size_t(b) * c.Cout + n] += rs;
      s2[size_t(b) * c.Cout + n] += double(rs) * rs;
      const size_t oi = ((size_t(b) * c.H + i) * c.W + j) * c.Cout + n;
      const float got = c.out_f32 ? reinterpret_cast<float*>(hOut.data())[oi]
                                  : bf2f(reinterpret_cast<uint16_t*>(hOut.data())[oi]);
      const double err = fabs(double(got) - ref);
      const double tol = (c.out_f32 ? 2e-3 : 1e-2) * (fabs(ref) + 1.0);
      if (err > tol) {
        if (bad < 8) printf("  mismatch b%d i%d j%d n%d got %f ref %f\n", b, i, j, n, got, ref);
        ++bad;
      }
      if (err > max_err) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
  }
  printf("conv check: %lld positions, max_err %.4g (max |ref| %.3g), bad %d\n", (long long)nsample, max_err,
         max_ref, bad);
  if (probe_fold) {
    printf("norm-backward fold: %s\n", L.nr_ok ? "active" : "NOT taken by this launch");
    if (L.nr_ok) {
      // per (image, channel): sum over tiles of the partials against a CPU loop over the STORED output (bf16) and nrY
      std::vector<float> part(size_t(c.B) * L.T128 * c.Cout * 2);
      CK(cudaMemcpy(part.data(), dNpart, part.size() * 4, cudaMemcpyDeviceToHost));
      const float n = float(c.H) * c.W;
      double worst = 0;
      for (int b = 0; b < c.B; ++b)
        for (int ch = 0; ch < c.Cout; ch += 37) {
          const float mu = hNs[(size_t(b) * c.Cout + ch) * 2] / n;
          const float rs = 1.f / sqrtf(fmaxf(hNs[(size_t(b) * c.Cout + ch) * 2 + 1] / n - mu * mu, 0.f) + 1e-3f);
          double r1 = 0, r2 = 0, g1 = 0, g2 = 0;
          for (int i = 0; i < c.H; ++i)
            for (int j = 0; j < c.W; ++j) {
              const size_t oi = ((size_t(b) * c.H + i) * c.W + j) * c.Cout + ch;
              const float d = bf2f(reinterpret_cast<uint16_t*>(hOut.data())[oi]);
              const float yc = bf2f(hY[oi]) - mu;
              const float dz = (yc * (hG[ch] * rs) + hBe[ch]) > 0.f ? d : 0.f;
              r1 += dz;
              r2 += double(dz) * yc * rs;
            }
          for (int t = 0; t < L.T128; ++t) {
            g1 += part[((size_t(b) * L.T128 + t) * c.Cout + ch) * 2];
            g2 += part[((size_t(b) * L.T128 + t) * c.Cout + ch) * 2 + 1];
          }
          worst = fmax(worst, fabs(g1 - r1) / (fabs(r1) + 1.0));
          worst = fmax(worst, fabs(g2 - r2) / (fabs(r2) + 1.0));
        }
      printf("fold check: max rel err %.4g\n", worst);
      if (worst > 2e-3) ++bad;
    }
  }
  if (c.use_stats && full_check) {
    double se = 0;
    for (size_t q = 0; q < s1.size(); ++q) {
      se = fmax(se, fabs(hStats[q * 2] - s1[q]) / (fabs(s1[q]) + 1.0));
      se = fmax(se, fabs(hStats[q * 2 + 1] - s2[q]) / (fabs(s2[q]) + 1.0));
    }
    printf("stats check: max rel err %.4g\n", se);
    if (se > 5e-3) ++bad;
  }
  {  // determinism: a second launch must reproduce the output bit for bit (statistics up to atomic order)
    CK(cudaMemset(dStats, 0, size_t(c.B) * c.Cout * 2 * 4));
    run_conv_gemm(L, 0);
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> hOut2(hOut.size());
    std::vector<float> hStats2(hStats.size());
    CK(cudaMemcpy(hOut2.data(), dOut, hOut2.size(), cudaMemcpyDeviceToHost));
    if (c.use_stats) read_stats(hStats2);
    size_t ndiff = 0;
    for (size_t q = 0; q < hOut.size(); ++q) ndiff += hOut[q] != hOut2[q];
    double sd = 0;
    for (size_t q = 0; q < hStats.size(); ++q) sd = fmax(sd, fabs(hStats[q] - hStats2[q]) / (fabs(hStats[q]) + 1.0));
    printf("DETERMINISM: %zu output bytes differ between two launches; stats max rel diff %.3g\n", ndiff, sd);
    if (ndiff || sd != 0.0) ++bad;
  }
  if (iters > 0) {
    if (L.swap) {
    } else if (L.pair) {  // stamps of the persistent CTA-pair kernel
      const int nct = 148;
      long long* dDbg;
      CK(cudaMalloc(&dDbg, size_t(nct) * 16 * sizeof(long long)));
      CK(cudaMemset(dDbg, 0, size_t(nct) * 16 * sizeof(long long)));
      ConvGemmLaunch Ld = L;
      Ld.p.dbg = dDbg;
      run_conv_gemm(Ld, 0);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(size_t(nct) * 16);
      CK(cudaMemcpy(h.data(), dDbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      for (int c = 0; c < 4; ++c) {
        const long long* t = &h[size_t(c) * 16];
        printf("PAIR cta %d: setup %lld | mma-issued", c, t[1] - t[0]);
        for (int k = 0; k < 6; ++k) if (t[2 + k]) printf(" %lld", t[2 + k] - t[0]);
        printf(" | epi-done");
        for (int k = 0; k < 6; ++k) if (t[8 + k]) printf(" %lld", t[8 + k] - t[0]);
        printf(" | end %lld\n", t[15] - t[0]);
      }
      CK(cudaFree(dDbg));
    } else
    {  // per-phase clock64 stamps of one launch
      const int nct = L.grid_x * L.grid_y * L.grid_z;
      long long* dDbg;
      CK(cudaMalloc(&dDbg, size_t(nct) * 8 * sizeof(long long)));
      CK(cudaMemset(dDbg, 0, size_t(nct) * 8 * sizeof(long long)));
      ConvGemmLaunch Ld = L;
      Ld.p.dbg = dDbg;
      run_conv_gemm(Ld, 0);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(size_t(nct) * 8);
      CK(cudaMemcpy(h.data(), dDbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      double s[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < nct; ++c) {
        const long long* t = &h[size_t(c) * 8];
        s[0] += double(t[1] - t[0]);  // setup (barrier init, TMEM alloc, sync)
        s[1] += double(t[2] - t[1]);  // MMA issue loop (ends when the last MMA is issued)
        s[2] += double(t[3] - t[2]);  // issue end -> accumulators complete
        s[3] += double(t[4] - t[3]);  // epilogue accumulator 0
        s[4] += L.p.MT == 256 ? double(t[5] - t[4]) : 0.0;
        s[5] += double(t[6] - t[0]);  // whole CTA
      }
      double st_end = 0, cta_end = 0;
      for (int c = 0; c < nct; ++c) {
        const long long* t = &h[size_t(c) * 8];
        st_end += t[7] ? double(t[7] - t[3]) : 0.0;  // statistics warps done, relative to the epilogue start
        cta_end += double(t[6] - t[3]);
      }
      printf("PHASES (avg SM cycles per CTA over %d CTAs): setup %.0f | mma-issue %.0f | drain %.0f | epi0 %.0f | epi1 %.0f | total %.0f"
             " || from epilogue start: stat warps done %.0f, CTA done %.0f\n",
             nct, s[0] / nct, s[1] / nct, s[2] / nct, s[3] / nct, s[4] / nct, s[5] / nct, st_end / nct, cta_end / nct);
      CK(cudaFree(dDbg));
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) run_conv_gemm(L, 0);
    CK(cudaEventRecord(e0));
    for (int it = 0; it < iters; ++it) run_conv_gemm(L, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double flops = 2.0 * c.B * c.H * c.W * double(c.Cout) * c.Cin * ntaps;
    printf("TIMING conv %dx%d k%d %d->%d B%d: %.3f ms  %.1f TFLOP/s (algorithmic)\n", c.H, c.W, c.k, c.Cin, c.Cout,
           c.B, ms, flops / ms * 1e-9);
  }
  return bad ? 1 : 0;
}

static int run_wgrad_case(int B, int H, int W, int Cx, int Cy, int BN, int k, int ksplit, int iters, bool full) {
  const int pad = k / 2, P = W + 2 * pad, rows = H + 2 * pad;
  const int64_t xpix = int64_t(rows) * P + 8, ypix = int64_t(H) * P + 8;
  const int ntaps = k * k;
  std::vector<uint16_t> hX(size_t(B) * xpix * Cx), hY(size_t(B) * ypix * Cy, 0);
  for (auto& v : hX) v = f2bf(frand());
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j)  // slack columns stay zero
        for (int c = 0; c < Cy; ++c) hY[(size_t(b) * ypix + size_t(i) * P + j) * Cy + c] = f2bf(frand());
  WgradParams p;
  memset(&p, 0, sizeof(p));
  uint16_t *dX, *dY;
  float* dW;
  const size_t wel = size_t(ntaps) * Cx * Cy;
  CK(cudaMalloc(&dX, hX.size() * 2));
  CK(cudaMalloc(&dY, hY.size() * 2));
  CK(cudaMalloc(&dW, wel * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dY, hY.data(), hY.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dW, 0, wel * 4));
  p.X = dX;
  p.x_frame_pix = xpix;
  p.x_row_stride = Cx;
  p.Cx = Cx;
  p.Y = dY;
  p.y_frame_pix = ypix;
  p.y_row_stride = Cy;
  p.Cy = Cy;
  p.nx_valid = Cx;
  p.ny_valid = Cy;
  p.BN = BN;
  p.B = B;
  p.ntaps = ntaps;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      p.x_off[kh * k + kw] = kh * P + kw;
      p.y_off[kh * k + kw] = 0;
    }
  p.Mpix = H * P;
  p.dW = dW;
  p.dw_tap_stride = int64_t(Cx) * Cy;
  p.dw_sx = Cy;
  p.dw_sy = 1;
  p.ksplit = ksplit;
  WgradLaunch L;
  int r = prepare_wgrad_gemm(p, &L);
  if (r) {
    printf("prepare_wgrad_gemm failed %d\n", r);
    return 1;
  }
  printf("grid %d x %d x %d, stages %d, smem %zu%s\n", L.grid_x, L.grid_y, L.grid_z, L.stages, L.smem, L.pair_groups ? " (CTA-pair kernel)" : "");
  r = run_wgrad_gemm(L, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (r || e != cudaSuccess) {
    printf("launch/sync failed r=%d err=%s\n", r, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hW(wel);
  CK(cudaMemcpy(hW.data(), dW, wel * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double max_err = 0, max_ref = 0;
  const int64_t nsample = full ? int64_t(wel) : 2048;
  for (int64_t s = 0; s < nsample; ++s) {
    const size_t idx = full ? size_t(s) : size_t((uint64_t(s) * 2654435761ull) % wel);
    const int t = int(idx / (size_t(Cx) * Cy)), x = int((idx / Cy) % Cx), y = int(idx % Cy);
    double acc = 0;
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
          const int64_t m = int64_t(i) * P + j;
          acc += double(bf2f(hX[(size_t(b) * xpix + m + p.x_off[t]) * Cx + x])) *
                 double(bf2f(hY[(size_t(b) * ypix + m) * Cy + y]));
        }
    const double err = fabs(hW[idx] - acc);
    if (err > 2e-3 * (fabs(acc) + 1.0)) {
      if (bad < 8) printf("  mismatch t%d x%d y%d got %f ref %f\n", t, x, y, hW[idx], acc);
      ++bad;
    }
    max_err = fmax(max_err, err);
    max_ref = fmax(max_ref, fabs(acc));
  }
  printf("wgrad check: %lld entries, max_err %.4g (max |ref| %.3g), bad %d\n", (long long)nsample, max_err, max_ref,
         bad);
  if (iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) run_wgrad_gemm(L, 0);
    CK(cudaEventRecord(e0));
    for (int it = 0; it < iters; ++it) run_wgrad_gemm(L, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double flops = 2.0 * B * H * W * double(Cx) * Cy * ntaps;
    printf("TIMING wgrad %dx%d k%d %dx%d B%d ksplit %d: %.3f ms  %.1f TFLOP/s (algorithmic)\n", H, W, k, Cx, Cy, B,
           ksplit, ms, flops / ms * 1e-9);
  }
  return bad ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// Experiment: can a K-major SWIZZLE_128B descriptor start at a 128-byte ROW offset inside the
// 1024-byte swizzle atom (shifted view of one smem tile = the kw taps of a convolution)?
__global__ void shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                   float* out, int shift_rows, int base_off) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, accbar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sA = smem;          // 256 rows x 128 B = 32 KB
  uint8_t* sB = smem + 32768;  // 64 rows x 128 B = 8 KB
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&accbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tbase, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 32768 + 8192);
    tma_load_2d(&tmA, &bar, sA, 0, 0);
    tma_load_2d(&tmA, &bar, sA + 16384, 0, 128);
    tma_load_2d(&tmB, &bar, sB, 0, 0);
    mbar_wait(&bar, 0, 21);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16_f32(128, 64, 0, 0);
    uint64_t adesc = desc_kmajor_sw128(smem_u32(sA) + shift_rows * 128) | (uint64_t(base_off & 7) << 49);
    uint64_t bdesc = desc_kmajor_sw128(smem_u32(sB));
    for (int k = 0; k < 4; ++k) umma_bf16(tbase, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, k != 0);
    umma_commit(&accbar);
  }
  mbar_wait(&accbar, 0, 22);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tmem_ld32(tbase + (uint32_t(warp * 32) << 16) + c0, v);
    for (int e = 0; e < 32; ++e) out[row * 64 + c0 + e] = v[e];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 64);
  }
}

static int run_shift_probe() {
  std::vector<uint16_t> hA(256 * 64), hB(64 * 64, 0);
  for (int n = 0; n < 64; ++n) hB[n * 64 + n] = f2bf(1.0f);
  uint16_t *dA, *dB;
  float* dOut;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dOut, 128 * 64 * 4));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, dA, 64, 256, 128, 64, 128) || make_tmap_bf16_2d(&tmB, dB, 64, 64, 128, 64, 64)) {
    printf("tmap encode failed\n");
    return 1;
  }
  CK(cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 17};
  std::vector<float> rowv(128 * 64), colv(128 * 64);
  for (int si = 0; si < 9; ++si) {
    for (int mode = 0; mode < 2; ++mode) {
      const int sh = shifts[si];
      const int bo = mode ? (sh & 7) : 0;
      if (mode == 1 && bo == 0) continue;
      for (int pass = 0; pass < 2; ++pass) {
        for (int r = 0; r < 256; ++r)
          for (int c = 0; c < 64; ++c) hA[r * 64 + c] = f2bf(pass == 0 ? float(r) : float(c));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        shift_probe_kernel<<<1, 128, 41 * 1024>>>(tmA, tmB, dOut, sh, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("shift %d bo %d: CUDA error %s\n", sh, bo, cudaGetErrorString(e));
          return 1;
        }
        CK(cudaMemcpy(pass == 0 ? rowv.data() : colv.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost));
      }
      int ok = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n)
          if (int(rowv[m * 64 + n]) == m + sh && int(colv[m * 64 + n]) == n) ++ok;
      printf("SHIFT rows=%d base_offset=%d : %d / 8192 correct", sh, bo, ok);
      if (ok != 8192) {
        printf("  e.g. D[0][0..63 step 8] src(row,col):");
        for (int n = 0; n < 64; n += 8) printf(" (%d,%d)", int(rowv[n]), int(colv[n]));
        printf("  D[1][0],D[7][8],D[8][0]: (%d,%d) (%d,%d) (%d,%d)", int(rowv[64]), int(colv[64]),
               int(rowv[7 * 64 + 8]), int(colv[7 * 64 + 8]), int(rowv[8 * 64]), int(colv[8 * 64]));
      }
      printf("\n");
    }
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// Experiment: raw tcgen05.mma issue rate with both operands resident in shared memory (no TMA traffic).
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n_mma, int N, int nacc, long long* cycles_out, int elect, int per_commit = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    fence_barrier_init();
  }
  uint32_t cols = 32;
  while ((int)cols < nacc * N) cols <<= 1;
  if (warp == 0) tmem_alloc(&tbase, cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (elect == 0 && threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N, 0, 0);
    const uint64_t adesc = desc_kmajor_sw128(smem_u32(smem));
    const uint64_t bdesc = desc_kmajor_sw128(smem_u32(smem) + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i)
      umma_bf16(tbase + uint32_t((i % nacc) * N), adesc + uint64_t((i & 3) * 2), bdesc + uint64_t((i & 3) * 2), idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0, 31);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  if (elect != 0 && warp == 0) {
    // the same loop under elect.sync: ptxas emits the UTCHMMAs back to back (no ELECT / BRA.U.ANY wrapper per instruction)
    if (elect_one_sync()) {
      const uint32_t idesc = idesc_bf16_f32(128, N, 0, 0);
      const uint64_t adesc = desc_kmajor_sw128(smem_u32(smem));
      const uint64_t bdesc = desc_kmajor_sw128(smem_u32(smem) + 16384);
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tbase + uint32_t(((i + k) % nacc) * N), adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, 1);
        if (per_commit > 0 && ((i + 4) % per_commit) == 0) umma_commit(&bar2);  // a stage release every per_commit MMAs
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0, 31);
      const long long t1 = clock64();
      cycles_out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, cols);
  }
}

// Same for a CTA pair: tcgen05.mma.cta_group::2, M = 256 over two CTAs, each CTA holding 128 A rows and N/2 B rows.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_rate_pair_kernel(int n_mma, int N, int nacc, int per_commit, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (16384 + N * 64) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    fence_barrier_init();
  }
  uint32_t cols = 32;
  while ((int)cols < nacc * N) cols <<= 1;
  if (warp == 0) tmem_alloc_pair(&tbase, cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 0 && elect_one_sync()) {  // elect.sync: no per-instruction ELECT / BRA.U.ANY wrapper (see mma_rate)
    const long long t0 = clock64();
    if (rank == 0) {
      const uint32_t idesc = idesc_bf16_f32(256, N, 0, 0);
      const uint64_t adesc = desc_kmajor_sw128(smem_u32(smem));
      const uint64_t bdesc = desc_kmajor_sw128(smem_u32(smem) + 16384);
      for (int i = 0; i < n_mma; i += 4) {  // four MMAs per iteration hide the loop's integer work
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_pair(tbase + uint32_t(((i + k) & (nacc - 1)) * N), adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, 1);
        if (per_commit > 0 && ((i + 4) % per_commit) == 0) umma_commit_pair(&bar2);
      }
      umma_commit_pair(&bar);
    }
    mbar_wait(&bar, 0, 31);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tbase, cols);
  }
}

static int run_mma_rate_pair() {
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(mma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int grid : {2, 148})
    for (int N : {256, 128, 64})
      for (int nacc : {1, 2})
        for (int pc : {0, 4, 12, 24}) {
          if (nacc * N > 512) continue;
          const int n_mma = 4096;
          cudaEvent_t e0, e1;
          CK(cudaEventCreate(&e0));
          CK(cudaEventCreate(&e1));
          mma_rate_pair_kernel<<<grid, 128, 58 * 1024>>>(n_mma, N, nacc, pc, d);
          CK(cudaEventRecord(e0));
          mma_rate_pair_kernel<<<grid, 128, 58 * 1024>>>(n_mma, N, nacc, pc, d);
          CK(cudaEventRecord(e1));
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("mma_rate_pair: CUDA error %s\n", cudaGetErrorString(e));
            return 1;
          }
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          long long cyc;
          CK(cudaMemcpy(&cyc, d, sizeof(cyc), cudaMemcpyDeviceToHost));
          const double flops = 2.0 * 256 * N * 16 * double(n_mma) * (grid / 2);
          printf("MMA_RATE_PAIR grid %3d N %3d nacc %d commit/%d: %.1f cycles/MMA (ideal %d per SM), kernel %.3f ms, %.1f TFLOP/s\n",
                 grid, N, nacc, pc, double(cyc) / n_mma, N / 2, ms, flops / ms * 1e-9);
        }
  return 0;
}

static int run_mma_rate() {
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int Ns[] = {256, 128, 64, 32};
  for (int elect : {0, 1})
  for (int grid : {1, 148})
    for (int N : Ns)
      for (int nacc : {1, 2}) {
        if (nacc * N > 512) continue;
        const int n_mma = 4096;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        mma_rate_kernel<<<grid, 128, 58 * 1024>>>(n_mma, N, nacc, d, elect);
        CK(cudaEventRecord(e0));
        mma_rate_kernel<<<grid, 128, 58 * 1024>>>(n_mma, N, nacc, d, elect);
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mma_rate: CUDA error %s\n", cudaGetErrorString(e));
          return 1;
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        long long cyc;
        CK(cudaMemcpy(&cyc, d, sizeof(cyc), cudaMemcpyDeviceToHost));
        const double flops = 2.0 * 128 * N * 16 * double(n_mma) * grid;
        printf("MMA_RATE %s grid %3d N %3d nacc %d: %.1f cycles/MMA (ideal %d), kernel %.3f ms, %.1f TFLOP/s\n", elect ? "elect.sync" : "tid==0", grid, N, nacc,
               double(cyc) / n_mma, N / 2, ms, flops / ms * 1e-9);
      }
  for (int pc : {4, 8, 12, 24, 48}) {
    const int n_mma = 4080;
    mma_rate_kernel<<<148, 128, 58 * 1024>>>(n_mma, 256, 2, d, 1, pc);
    mma_rate_kernel<<<148, 128, 58 * 1024>>>(n_mma, 256, 2, d, 1, pc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma_rate commit sweep: CUDA error\n"); return 1; }
    long long cyc;
    CK(cudaMemcpy(&cyc, d, sizeof(cyc), cudaMemcpyDeviceToHost));
    printf("MMA_RATE elect.sync grid 148 N 256 nacc 2, tcgen05.commit every %2d MMAs: %.1f cycles/MMA\n", pc, double(cyc) / n_mma);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Experiment: L2 -> SM delivery rate of TMA tiles when the CTAs of a cluster all need the SAME tile
// (unicast: every CTA loads it; multicast: each CTA loads 1/csz of it for everybody).
__global__ void __launch_bounds__(64, 1) tma_share_kernel(const __grid_constant__ CUtensorMap tm,
                                                          const __grid_constant__ CUtensorMap tm_strided, int iters, int mode,
                                                          int ntiles, long long* cycles_out, const uint8_t* raw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[4], empty[4];
  uint32_t csz, rank, cid;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csz));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
  constexpr int kRows = 256, kTileBytes = kRows * 128;  // 32 KB tile of 256 rows x 64 bf16
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode == 1 ? csz : 1); }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {  // producer
    for (int it = 0; it < iters; ++it) {
      const int s = it & 3;
      mbar_wait(&empty[s], ((it >> 2) & 1) ^ 1, 71);
      int tile = int((uint64_t(cid) * 977u + uint64_t(it)) % uint64_t(ntiles));
      if (mode == 5 || mode == 6) tile = it % 36;  // every CTA walks the SAME 36 tiles (1.2 MB: one layer's weights)
      mbar_arrive_expect_tx(&full[s], kTileBytes);
      if (mode == 4 || mode == 6) {
        // rows of 128 B strided by 512 B (a 64-channel slice of a [pixel][256 channel] tensor): tile -> (row block, slice)
        const int rb = tile >> 2, cb = tile & 3;
        tma_load_2d(&tm_strided, &full[s], smem + s * kTileBytes, cb * 64, rb * kRows);
        tma_load_2d(&tm_strided, &full[s], smem + s * kTileBytes + kTileBytes / 2, cb * 64, rb * kRows + kRows / 2);
      } else if (mode == 0 || mode == 5) {
        tma_load_2d(&tm, &full[s], smem + s * kTileBytes, 0, tile * kRows);
        tma_load_2d(&tm, &full[s], smem + s * kTileBytes + kTileBytes / 2, 0, tile * kRows + kRows / 2);
      } else if (mode == 2) {  // the same bytes as ONE 1-D bulk copy (what the row-stream kernels issue)
        bulk_load_1d(smem + s * kTileBytes, raw + size_t(tile) * kTileBytes, kTileBytes, &full[s]);
      } else if (mode == 3) {  // ... or as two 16 KB 1-D bulk copies
        bulk_load_1d(smem + s * kTileBytes, raw + size_t(tile) * kTileBytes, kTileBytes / 2, &full[s]);
        bulk_load_1d(smem + s * kTileBytes + kTileBytes / 2, raw + size_t(tile) * kTileBytes + kTileBytes / 2, kTileBytes / 2,
                     &full[s]);
      } else {
        const int rows = kRows / int(csz);  // this CTA's slice, delivered to every CTA of the cluster
        // box height is fixed at 128 rows by the tensor map: csz == 2 -> one box each
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(smem_u32(smem + s * kTileBytes + int(rank) * rows * 128)), "l"(reinterpret_cast<uint64_t>(&tm)),
              "r"(smem_u32(&full[s])), "r"(0), "r"(tile * kRows + int(rank) * rows), "h"(uint16_t((1u << csz) - 1))
            : "memory");
      }
    }
  } else if (threadIdx.x == 32) {  // consumer: nothing to compute, just recycle the stage
    for (int it = 0; it < iters; ++it) {
      const int s = it & 3;
      mbar_wait(&full[s], (it >> 2) & 1, 72);
      if (mode != 1) mbar_arrive(&empty[s]);
      else
        for (uint32_t r = 0; r < csz; ++r) mbar_arrive_cluster(mapa_u32(smem_u32(&empty[s]), r));
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

static int run_tma_share() {
  const size_t rows = size_t(1) << 19;  // 64 MB of 128-byte rows: L2 resident
  uint16_t* d;
  CK(cudaMalloc(&d, rows * 128));
  CK(cudaMemset(d, 0, rows * 128));
  CUtensorMap tm, tm_strided;
  int r = make_tmap_bf16_2d(&tm, d, 64, rows, 128, 64, 128);
  if (r) { printf("tmap failed %d\n", r); return 1; }
  r = make_tmap_bf16_2d(&tm_strided, d, 256, rows / 4, 512, 64, 128);  // the same memory seen as [rows/4][256 channels]
  if (r) { printf("strided tmap failed %d\n", r); return 1; }
  long long* dc;
  CK(cudaMalloc(&dc, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(tma_share_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 1024));
  const int iters = 2000, ntiles = int(rows / 256);
  for (int csz : {1, 2}) {
    for (int mode : {0, 1, 2, 3, 4, 5, 6}) {
      if (csz == 1 && mode == 1) continue;
      if (csz == 2 && mode >= 2) continue;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(64);
      cfg.dynamicSmemBytes = 4 * 32768 + 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, tma_share_kernel, tm, tm_strided, iters, mode, ntiles, dc, (const uint8_t*)d);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sync failed: %s (watchdog %d)\n", cudaGetErrorString(e), read_tc_watchdog()); return 1; }
      }
      std::vector<long long> h(148);
      CK(cudaMemcpy(h.data(), dc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (auto v : h) avg += double(v);
      avg /= 148;
      printf("TMA_SHARE cluster %d %s: %.1f B/clk delivered per SM (%.0f cycles per 32 KB tile), chip %.0f B/clk\n", csz,
             mode == 1 ? "multicast" : (mode == 0 ? "unicast 2-D boxes" : (mode == 2 ? "1-D bulk 32 KB" : (mode == 3 ? "1-D bulk 2x16 KB" : (mode == 4 ? "2-D boxes, 128 B rows strided by 512 B" : (mode == 5 ? "dense boxes, all CTAs read the same 36 tiles" : "strided boxes, all CTAs read the same 36 tiles"))))), iters * 32768.0 / avg, avg / iters, 148 * iters * 32768.0 / avg);
    }
  }
  return 0;
}

static int run_tmap_overlap() {
  // Overlapping-window map: rows of 64 bf16 that start every 8 elements (16 B).  Used for the
  // Cin=3 (padded to 8) 7x7 convolution if the driver accepts it.
  uint16_t* d;
  CK(cudaMalloc(&d, 1 << 20));
  CUtensorMap tm;
  int r = make_tmap_bf16_2d(&tm, d, 64, 4096, 16, 64, 128);
  printf("TMAP_OVERLAP encode (dim0=64 elems, row stride 16 B) -> %d (0 = accepted)\n", r);
  return 0;
}

int main(int argc, char** argv) {
  const char* t = argc > 1 ? argv[1] : "conv_small";
  int rc = 1;
  if (!strcmp(t, "conv_small")) {
    ConvCase c = {2, 6, 20, 128, 128, 128, 128, 3, SG_ACT_NONE, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_small64")) {
    ConvCase c = {1, 5, 9, 64, 64, 64, 64, 3, SG_ACT_LRELU, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_res")) {
    ConvCase c = {8, 64, 128, 256, 256, 256, 256, 3, SG_ACT_NONE, 0, 1};
    rc = run_conv_case(c, 20, false);
  } else if (!strcmp(t, "conv_res_nostats")) {  // the dgrad-like form of conv_res (no statistics): PROBE_FOLD=1 adds the fold
    ConvCase c = {8, 64, 128, 256, 256, 256, 256, 3, SG_ACT_NONE, 0, 0};
    rc = run_conv_case(c, 20, false);
  } else if (!strcmp(t, "conv_fold_check")) {  // odd tile count, ragged last tile, every position checked
    ConvCase c = {3, 51, 126, 64, 256, 256, 256, 3, SG_ACT_NONE, 0, 0};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_pair_check")) {
    ConvCase c = {2, 80, 128, 64, 256, 256, 256, 3, SG_ACT_LRELU, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_pair_odd")) {  // odd tile count: the last pair has a dummy peer
    ConvCase c = {3, 51, 126, 64, 256, 256, 256, 3, SG_ACT_NONE, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_swap128")) {  // Cout 128: transposed kernel, full check incl. statistics
    ConvCase c = {2, 40, 126, 64, 128, 128, 128, 3, SG_ACT_LRELU, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_swap64")) {
    ConvCase c = {3, 37, 90, 128, 64, 64, 64, 3, SG_ACT_RELU, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_swap_big")) {  // the 128x256-resolution 3x3 layers of the generator
    ConvCase c = {8, 128, 256, 64, 128, 128, 128, 3, SG_ACT_NONE, 0, 1};
    rc = run_conv_case(c, 20, false);
  } else if (!strcmp(t, "conv_res128")) {
    ConvCase c = {8, 64, 128, 256, 256, 256, 128, 3, SG_ACT_NONE, 0, 1};
    rc = run_conv_case(c, 20, false);
  } else if (!strcmp(t, "conv_small256")) {
    ConvCase c = {2, 6, 50, 128, 128, 128, 128, 3, SG_ACT_NONE, 0, 1, 256};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_res_mt128")) {
    ConvCase c = {8, 64, 128, 256, 256, 256, 256, 3, SG_ACT_NONE, 0, 1, 128};
    rc = run_conv_case(c, 20, false);
  } else if (!strcmp(t, "conv_out7_big")) {
    ConvCase c = {8, 256, 512, 64, 3, 32, 32, 7, SG_ACT_TANH, 1, 0, 0};
    rc = run_conv_case(c, 10, false);
  } else if (!strcmp(t, "conv_out7_mid")) {  // shift-sum epilogue of the transposed persistent kernel, full check
    ConvCase c = {2, 40, 200, 64, 3, 32, 32, 7, SG_ACT_TANH, 1, 0, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_out7_shift_small")) {  // shift-sum epilogue of the single-CTA kernel (few tiles)
    ConvCase c = {1, 8, 40, 64, 3, 32, 32, 7, SG_ACT_TANH, 1, 0, 0, 1};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "conv_out7_shift_big")) {
    ConvCase c = {8, 256, 512, 64, 3, 32, 32, 7, SG_ACT_TANH, 1, 0, 0, 1};
    rc = run_conv_case(c, 10, false);
  } else if (!strcmp(t, "conv_out7")) {
    ConvCase c = {1, 8, 40, 64, 3, 32, 32, 7, SG_ACT_TANH, 1, 0};
    rc = run_conv_case(c, 0, true);
  } else if (!strcmp(t, "wgrad_small")) {
    rc = run_wgrad_case(2, 6, 20, 128, 64, 64, 3, 3, 0, true);
  } else if (!strcmp(t, "wgrad_pair_small")) {  // CTA-pair kernel (Cx, Cy multiples of 256), every entry checked
    rc = run_wgrad_case(2, 6, 20, 256, 256, 256, 3, 3, 0, true);
  } else if (!strcmp(t, "wgrad_pair_wide")) {  // two x blocks and two y blocks, ragged pixel count
    rc = run_wgrad_case(1, 5, 13, 512, 512, 256, 3, 2, 0, false);
  } else if (!strcmp(t, "wgrad_res14")) {  // the split the engine picks for the pair kernel (5 tap groups x 14 slices)
    rc = run_wgrad_case(8, 64, 128, 256, 256, 256, 3, 14, 10, false);
  } else if (!strcmp(t, "wgrad_res")) {
    rc = run_wgrad_case(8, 64, 128, 256, 256, 256, 3, 8, 10, false);
  } else if (!strcmp(t, "tma_share")) {
    rc = run_tma_share();
  } else if (!strcmp(t, "mma_rate_pair")) {
    rc = run_mma_rate_pair();
  } else if (!strcmp(t, "mma_rate")) {
    rc = run_mma_rate();
  } else if (!strcmp(t, "conv_overhead")) {
    ConvCase c = {8, 64, 128, 64, 256, 256, 256, 1, SG_ACT_NONE, 0, 1, 256};
    rc = run_conv_case(c, 20, false);
    ConvCase c2 = {8, 64, 128, 64, 256, 256, 256, 1, SG_ACT_NONE, 0, 0, 256};
    rc |= run_conv_case(c2, 20, false);
    ConvCase c3 = {8, 64, 128, 64, 256, 256, 256, 1, SG_ACT_NONE, 0, 0, 128};
    rc |= run_conv_case(c3, 20, false);
  } else if (!strcmp(t, "shift")) {
    rc = run_shift_probe();
  } else if (!strcmp(t, "tmap_overlap")) {
    rc = run_tmap_overlap();
  }
  printf("watchdog flag: %d\n", read_tc_watchdog());
  printf("RESULT %s %s\n", t, rc == 0 ? "PASS" : "FAIL");
  return rc;
}
