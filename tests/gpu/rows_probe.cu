// rows_probe.cu -- standalone timing probe for the row-stream (instance-norm / gather) passes at the
// residual-block shape of BASELINE config 3 (B=8, 64x128x256) and the 128x256x128 / 256x512x64 layers.
// Usage: rows_probe [B H W C]   -- prints CUDA-event time per launch over rotating buffer sets (cold L2)
// and per-block phase stamps.  Timing only; correctness is covered by tests/test_gpu_parity.py.
#define SG_ROWS_DEBUG 1
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../sg-gan-tf2_b200/csrc/glue_rows.cu"

using namespace sggan;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

struct Set {
  sg_bf16 *Y, *X, *R, *dX, *dY, *gat;
  float *part, *stats, *sums, *gamma, *beta, *bpart;
  int* ctr;
};

int main(int argc, char** argv) {
  int B = 8, H = 64, W = 128, C = 256;
  if (argc >= 5) { B = atoi(argv[1]); H = atoi(argv[2]); W = atoi(argv[3]); C = atoi(argv[4]); }
  const int NSETS = argc >= 6 ? atoi(argv[5]) : 4, ITERS = 24;  // 1 set: the whole working set stays L2-resident
  const int P = W + 2, T = (H * P + 127) / 128 + 1;
  FrameMap xm;
  memset(&xm, 0, sizeof(xm));
  xm.frame_pix = int64_t(H + 2) * P; xm.C = C; xm.H = H; xm.W = W; xm.kind = 0; xm.P = P; xm.pt = 1; xm.pl = 1; xm.reflect = 1;
  FrameMap dym;
  memset(&dym, 0, sizeof(dym));
  dym.frame_pix = int64_t(H + 4) * P; dym.C = C; dym.H = H; dym.W = W; dym.kind = 0; dym.P = P; dym.pt = 2; dym.pl = 0;
  const size_t nY = size_t(B) * H * W * C, nX = size_t(B) * xm.frame_pix * C, ndX = size_t(B) * (H + 2) * P * C,
               ndY = size_t(B) * dym.frame_pix * C;
  std::vector<Set> sets(NSETS);
  std::vector<uint16_t> h(nX + 4096);
  for (size_t i = 0; i < h.size(); ++i) h[i] = uint16_t(0x3c00 + (i * 2654435761u >> 22));  // small positive/neg mix is irrelevant for timing
  for (auto& s : sets) {
    CK(cudaMalloc(&s.Y, nY * 2)); CK(cudaMalloc(&s.X, nX * 2 + 4096)); CK(cudaMalloc(&s.R, nX * 2 + 4096));
    CK(cudaMalloc(&s.dX, ndX * 2)); CK(cudaMalloc(&s.dY, ndY * 2 + 4096)); CK(cudaMalloc(&s.gat, nY * 2));
    CK(cudaMalloc(&s.part, size_t(B) * T * C * 8)); CK(cudaMalloc(&s.stats, size_t(B) * C * 8));
    CK(cudaMalloc(&s.sums, size_t(B) * C * 8)); CK(cudaMalloc(&s.gamma, C * 4)); CK(cudaMalloc(&s.beta, C * 4));
    CK(cudaMalloc(&s.ctr, 64 * sizeof(int)));
    CK(cudaMalloc(&s.bpart, in_bwd_partials_bytes(C))); CK(cudaMemset(s.bpart, 0, in_bwd_partials_bytes(C)));
    CK(cudaMemcpy(s.Y, h.data(), nY * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.R, h.data(), nX * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.dX, h.data(), ndX * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(s.part, 0, size_t(B) * T * C * 8)); CK(cudaMemset(s.stats, 0, size_t(B) * C * 8));
    CK(cudaMemset(s.sums, 0, size_t(B) * C * 8)); CK(cudaMemset(s.gamma, 0, C * 4)); CK(cudaMemset(s.beta, 0, C * 4));
    CK(cudaMemset(s.X, 0, nX * 2)); CK(cudaMemset(s.dY, 0, ndY * 2));
  }
  long long* dbg;
  CK(cudaMalloc(&dbg, 8 * 256 * 16 * 8));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  auto apply = [&](Set& s, bool res, bool part) {
    InApplyParams p;
    memset(&p, 0, sizeof(p));
    p.Y = s.Y; p.B = B; p.H = H; p.W = W; p.C = C; p.stats = s.stats; p.gamma = s.gamma; p.beta = s.beta; p.eps = 1e-3f;
    p.act = res ? SG_ACT_NONE : SG_ACT_RELU;
    if (part) { p.stats_part = s.part; p.stats_T = T; p.stats_out = s.stats; }
    if (res) { p.res = s.R; p.rmap = xm; }
    p.dst = s.X; p.dmap = xm;
    launch_in_apply(p, st);
  };
  auto bwd = [&](Set& s, int which, bool gather) {
    InBwdParams p;
    memset(&p, 0, sizeof(p));
    p.Y = s.Y; p.B = B; p.H = H; p.W = W; p.C = C; p.nb_act = B; p.act_wrap = 0; p.stats = s.stats; p.gamma = s.gamma;
    p.beta = s.beta; p.eps = 1e-3f; p.act = SG_ACT_RELU;
    p.g1.ptr = s.dX; p.g1.f32 = 0; p.g1.Hs = H + 2; p.g1.Ws = P; p.g1.oy = 1; p.g1.ox = 1; p.g1.fold = 1;
    if (gather) { p.g2 = p.g1; p.g2.ptr = s.R; p.gather_dst = s.gat; }
    p.sums = s.sums; p.sums_part = s.bpart; p.dst = s.dY; p.dmap = dym;
    if (which == 2) {
      CK(cudaMemsetAsync(s.ctr, 0, 64 * sizeof(int), st));
      p.sync_ctr = s.ctr;
      const int r = launch_in_bwd_fused(p, st);
      if (r < 0) { printf("fused launch failed %d\n", r); exit(3); }
    } else if (which == 0) launch_in_bwd_reduce(p, st);
    else { p.gather_dst = nullptr; p.sums_nblk = 148 / B > 0 ? 148 / B : 1; launch_in_bwd_apply(p, st); }
  };
  struct Case { const char* name; int kind; double mb; };
  const double mY = nY * 2 / 1e6;
  Case cases[] = {{"apply (stats given)", 0, 2 * mY},       {"apply + finalize", 1, 2 * mY},
                  {"apply + residual + finalize", 2, 3 * mY}, {"bwd reduce", 3, 2 * mY},
                  {"bwd reduce + gather", 4, 4 * mY},        {"bwd apply", 5, 3 * mY},
                  {"bwd reduce -> apply (same set)", 6, 5 * mY}, {"bwd fused (reduce + apply)", 7, 3 * mY},
                  {"bwd fused + gather", 8, 5 * mY}};
  printf("rows_probe B %d H %d W %d C %d  (tensor %.1f MB)  stages %d consumers %d\n", B, H, W, C, mY, kStages, kConsumers);
  for (auto& c : cases) {
    auto run = [&](Set& s) {
      switch (c.kind) {
        case 0: apply(s, false, false); break;
        case 1: apply(s, false, true); break;
        case 2: apply(s, true, true); break;
        case 3: bwd(s, 0, false); break;
        case 4: bwd(s, 0, true); break;
        case 5: bwd(s, 1, false); break;
        case 6: bwd(s, 0, false); bwd(s, 1, false); break;
        case 7: bwd(s, 2, false); break;
        case 8: bwd(s, 2, true); break;
      }
    };
    for (int i = 0; i < NSETS; ++i) run(sets[i]);
    CK(cudaStreamSynchronize(st));
    // (a) back-to-back over rotating sets (PDL chaining as in the step)
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < ITERS; ++i) run(sets[i % NSETS]);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / ITERS;
    // (b) phase stamps of three back-to-back launches (cycles within a block; %globaltimer across blocks / launches)
    CK(cudaMemset(dbg, 0, 8 * 256 * 16 * 8));
    CK(cudaMemcpyToSymbol(g_rows_dbg, &dbg, sizeof(dbg)));
    g_rows_dbg_launch = 0;
    for (int i = 0; i < 3; ++i) run(sets[i % NSETS]);
    CK(cudaStreamSynchronize(st));
    long long* nullp = nullptr;
    CK(cudaMemcpyToSymbol(g_rows_dbg, &nullp, sizeof(nullp)));
    std::vector<long long> hd(8 * 256 * 16);
    CK(cudaMemcpy(hd.data(), dbg, hd.size() * 8, cudaMemcpyDeviceToHost));
    double acc[8] = {0};
    int nblk = 0;
    long long s0[3], s1[3], e0[3], e1[3];
    for (int l = 0; l < 3; ++l) {
      s0[l] = e0[l] = (1ll << 62); s1[l] = e1[l] = 0;
      for (int bk = 0; bk < 256; ++bk) {
        const long long* d = &hd[(l * 256 + bk) * 16];
        if (d[0] == 0) continue;
        s0[l] = std::min(s0[l], d[8]); s1[l] = std::max(s1[l], d[8]);
        e0[l] = std::min(e0[l], d[9]); e1[l] = std::max(e1[l], d[9]);
        if (l == 1) { for (int k = 1; k < 7; ++k) acc[k] += double(d[k] - d[0]); ++nblk; }
      }
    }
    // per block-x position (averaged over the images): lifetime after the dependency wait, in cycles
    {
      const int gx = nblk / B > 0 ? nblk / B : 1;
      printf("    block x -> cycles from dep-wait to end (avg over images):");
      for (int x = 0; x < gx; ++x) {
        double a = 0;
        for (int bb = 0; bb < B; ++bb) { const long long* d = &hd[(1 * 256 + bb * gx + x) * 16]; a += double(d[5] - d[1]); }
        printf(" %.0f", a / B);
      }
      printf("\n");
    }
    printf("%-30s %7.2f us/launch  %6.0f GB/s (algorithmic %.0f MB) | blocks %d, avg cycles from block start: dep-wait %.0f "
           "coeffs %.0f first-chunk %.0f last-load-issued %.0f last-consumed %.0f end %.0f\n",
           c.name, us, c.mb / us * 1e3, c.mb, nblk, acc[1] / nblk, acc[2] / nblk, acc[3] / nblk, acc[6] / nblk,
           acc[4] / nblk, acc[5] / nblk);
    printf("    launch 2 of 3 (ns): first block start 0, last block start %lld, first block end %lld, last block end %lld | "
           "previous launch's last block ended at %lld, next launch's first block started at %lld\n",
           s1[1] - s0[1], e0[1] - s0[1], e1[1] - s0[1], e1[0] - s0[1], s0[2] - s0[1]);
  }
  return 0;
}
