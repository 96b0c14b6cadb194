// conv_gemm_tc.h -- host interface of the tcgen05 implicit-GEMM kernels (see conv_gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "kparams.h"

namespace sggan {

// A prepared launch: parameters + TMA tensor maps (encoded once, reusable every step and
// capturable in a CUDA graph).
struct ConvGemmLaunch {
  ConvGemmParams p;
  CUtensorMap tmA, tmA8, tmB;  // A boxes of 128 rows / 8 halo rows, weight boxes of BN rows
  int grid_x, grid_y, grid_z;
  int sa_stages, sb_stages;  // depth of the A ring and of the B ring
  unsigned tmem_cols;
  size_t smem;
  int stat_tiles;          // 128-row statistics tiles per image this launch writes
  int pair, T128, npairs;  // CTA-pair persistent kernel (256-channel layers): tiles per image, pair tiles in all
  int nr_ok = 0;           // the launch also folds the norm-backward sums of the layer below (ConvGemmParams::nr_*)
  CUtensorMap tmBh;        // 128-row weight tile (pair kernel: half of the channels; swap kernel: the M operand)
  int swap, T256, swap_pstages, swap_wstages, swap_nblk;  // transposed persistent kernel (<= 128 output channels): 256-pixel tiles per image
  size_t swap_smem;
};
struct WgradLaunch {
  WgradParams p;
  CUtensorMap tmX, tmY;
  int grid_x, grid_y, grid_z;
  int stages;
  int na;  // accumulators (128 x-channels each) per CTA
  int pair_groups = 0;  // > 0: the CTA-pair kernel with this many tap groups (wgrad_pair_groups)
  unsigned tmem_cols;
  size_t smem;
};

// All return 0 on success, a negative SGGAN error code otherwise.  Launches are asynchronous.
int prepare_conv_gemm(const ConvGemmParams& p, ConvGemmLaunch* L);
int run_conv_gemm(const ConvGemmLaunch& L, cudaStream_t st);
int prepare_wgrad_gemm(const WgradParams& p, WgradLaunch* L);
// tap groups the CTA-pair weight-gradient kernel would use for this shape (0: not eligible); the arrays may be null
int wgrad_pair_groups(const WgradParams& p, uint8_t* pair_a, uint8_t* pair_b);
int run_wgrad_gemm(const WgradLaunch& L, cudaStream_t st);
// dW[i] += sum over split-K slices (fixed order) of the partial tiles, i < numel
void launch_wgrad_reduce(const WgradLaunch& L, int64_t numel, cudaStream_t st);
int read_tc_watchdog();

}  // namespace sggan
