// conv_gemm_tc.cu -- the two tensor-core kernels of the SG-GAN step, hand-written for sm_100a:
//
//   conv_gemm_tc_kernel   implicit-GEMM convolution (forward conv, dgrad, deconv phases) over
//                         pitch-linearised frames: TMA boxes -> smem (SWIZZLE_128B) ->
//                         tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> epilogue (bias, act,
//                         per-(image,channel) sum / sum^2 for instance norm, bf16 store).
//   wgrad_gemm_tc_kernel  weight gradient: MN-major operands (pixels are K), split-K over
//                         images x pixel chunks, fp32 red.global accumulation.
//
// Replaces the cuDNN/Eigen calls behind tf.keras.layers.Conv2D / Conv2DTranspose and their
// gradients on the reference path (module.py:211-216,232-265,284-311; model.py:196-197).
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA
// issuer (one lane), warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31).
#include "conv_gemm_tc.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace sggan {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                          // bf16 elements per 128-byte swizzled row
constexpr int kABytes = kTileM * kChunkK * 2;        // 16 KB
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;              // pipeline bytes (leaves room for alignment slack)
constexpr int kThreads = 192;

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

static inline uint32_t tmem_cols_for(int bn) {
  uint32_t c = 32;
  while ((int)c < bn) c <<= 1;
  return c;
}

// =============================================================================================
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const ConvGemmParams p, const int stages, const uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int stage_bytes = kABytes + BN * 128;
  const int m0 = blockIdx.x * kTileM, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int cchunks = p.Cin / kChunkK;
  const int ksteps = p.ntaps * cchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&tmem_base_sh, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        const int tap = ks / cchunks, cc = ks - tap * cchunks;
        uint8_t* sa = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        tma_load_3d(&tmA, &full_bar[s], sa, cc * kChunkK, m0 + p.tap_off[tap], b);
        tma_load_2d(&tmB, &full_bar[s], sa + kABytes, cc * kChunkK, int(p.tap_w[tap]) * p.CoutPad + n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(kTileM, BN, 0, 0);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&full_bar[s], ph, 2);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint64_t adesc = desc_kmajor_sw128(sa);
        const uint64_t bdesc = desc_kmajor_sw128(sa + kABytes);
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes inside the swizzle row
          umma_bf16(tmem_acc, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (ks | k) != 0);
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(&acc_bar);  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue (128 threads)
    mbar_wait(&acc_bar, 0, 3);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const int i = m / p.P, j = m - i * p.P;
    const int oi = i * p.o_scale + p.o_a, oj = j * p.o_scale + p.o_b;
    const bool valid = (m < p.M) && (i < p.Hv) && (j < p.Wv) && (oi < p.omap.H) && (oj < p.omap.W);
    float* tsm = reinterpret_cast<float*>(smem);  // [BN][129] transposed fp32 tile for the statistics
    const int64_t obase = (int64_t(b) * p.omap.frame_pix + (valid ? frame_pixel(p.omap, oi, oj) : 0)) * p.omap.C;
    const bool vec_ok = (!p.out_f32) && ((p.omap.C & 7) == 0);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
      const int nb = n0 + c0;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int n = nb + e;
        float x = v[e];
        if (p.bias != nullptr && n < p.Cout) x += __ldg(p.bias + n);
        x = apply_act(x, p.act, p.act_alpha);
        // statistics are taken over the values as stored (bf16), so that the normalisation that
        // follows is exact for what it reads (H*W == 1 must give exactly beta)
        v[e] = p.out_f32 ? x : __bfloat162float(__float2bfloat16_rn(x));
      }
      if (p.stats != nullptr) {
#pragma unroll
        for (int e = 0; e < 32; ++e) tsm[(c0 + e) * 129 + row] = (valid && nb + e < p.Cout) ? v[e] : 0.f;
      }
      if (valid) {
        if (vec_ok && nb + 32 <= p.Cout) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + obase + nb);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
            w.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
            w.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
            w.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
            dst[g] = w;
          }
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
    if (p.stats != nullptr) {
      named_bar_sync(1, 128);
      const int t = threadIdx.x - 64;  // 0..127
      for (int c = t; c < BN; c += 128) {
        const int n = n0 + c;
        if (n >= p.Cout) continue;
        const float* col = tsm + c * 129;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 128; ++r) {
          const float x = col[r];
          s1 += x;
          s2 += x * x;
        }
        float* dst = p.stats + (int64_t(b) * p.Cout + n) * 2;
        atomicAdd(dst, s1);
        atomicAdd(dst + 1, s2);
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
__global__ void __launch_bounds__(kThreads, 1)
wgrad_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                     const WgradParams p, const int stages, const uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int stage_bytes = kABytes + BN * 128;
  const int mtiles = p.x_pair ? 1 : p.Cx / kTileM;
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {
      const int xoff = p.x_off[tap], yoff = p.y_off[tap];
      const int xoff2 = p.x_pair ? p.x_off2[tap] : xoff;
      const int xc0 = p.x_pair ? 0 : mt * kTileM, xc1 = p.x_pair ? 0 : mt * kTileM + 64;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 11);
        const int c = kbeg + ks;
        const int b = c / nchunk, mc = (c - b * nchunk) * 64;
        uint8_t* sa = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        tma_load_3d(&tmX, &full_bar[s], sa, xc0, mc + xoff, b);
        tma_load_3d(&tmX, &full_bar[s], sa + 8192, xc1, mc + xoff2, b);
        for (int h = 0; h < BN / 64; ++h)
          tma_load_3d(&tmY, &full_bar[s], sa + kABytes + h * 8192, n0 + h * 64, mc + yoff, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(kTileM, BN, 1, 1);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&full_bar[s], ph, 12);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint64_t adesc = desc_mnmajor_sw128(sa, 8192);
        const uint64_t bdesc = desc_mnmajor_sw128(sa + kABytes, 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 pixels (= 16 rows of 128 B) per MMA
          umma_bf16(tmem_acc, adesc + uint64_t(k * (2048 >> 4)), bdesc + uint64_t(k * (2048 >> 4)), idesc,
                    (ks | k) != 0);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    mbar_wait(&acc_bar, 0, 13);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int xg = mt * kTileM + row;
    float* dst = p.dW + int64_t(tap) * p.dw_tap_stride + int64_t(xg) * p.dw_sx;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c0), v);  // warp-collective: no divergence here
      if (xg < p.nx_valid) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (n0 + c0 + e < p.ny_valid) atomicAdd(dst + int64_t(n0 + c0 + e) * p.dw_sy, v[e]);
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
// Host side
static int stages_for(int bn) {
  int st = kSmemBudget / (kABytes + bn * 128);
  return st > kMaxStages ? kMaxStages : st;
}

int prepare_conv_gemm(const ConvGemmParams& p, ConvGemmLaunch* L) {
  if (p.Cin % 64 != 0 || p.Cin <= 0) return -10;
  if (!(p.BN == 32 || p.BN == 64 || p.BN == 128 || p.BN == 256)) return -11;
  if (p.CoutPad % p.BN != 0 || p.ntaps < 1 || p.ntaps > SGGAN_MAX_TAPS) return -12;
  L->p = p;
  L->stages = stages_for(p.BN);
  L->tmem_cols = tmem_cols_for(p.BN);
  L->smem = size_t(L->stages) * (kABytes + p.BN * 128) + 1024;
  L->grid_x = (p.M + kTileM - 1) / kTileM;
  L->grid_y = p.CoutPad / p.BN;
  L->grid_z = p.B;
  int r = make_tmap_bf16_3d(&L->tmA, p.A, p.Cin, p.a_frame_pix, p.B, uint64_t(p.a_row_stride) * 2,
                            uint64_t(p.a_frame_pix) * p.a_row_stride * 2, 64, 128);
  if (r) return -1000 - r;
  r = make_tmap_bf16_2d(&L->tmB, p.Wt, p.Cin, uint64_t(p.wt_taps) * p.CoutPad, uint64_t(p.Cin) * 2, 64, p.BN);
  if (r) return -2000 - r;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget + 1024);
    if (e != cudaSuccess) return -3000 - int(e);
    attr_set = true;
  }
  return 0;
}

int run_conv_gemm(const ConvGemmLaunch& L, cudaStream_t st) {
  dim3 grid(L.grid_x, L.grid_y, L.grid_z);
  conv_gemm_tc_kernel<<<grid, kThreads, L.smem, st>>>(L.tmA, L.tmB, L.p, L.stages, L.tmem_cols);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -4000 - int(e);
}

int prepare_wgrad_gemm(const WgradParams& p, WgradLaunch* L) {
  if (p.Cx <= 0 || (p.x_pair ? p.Cx != 64 : p.Cx % 128 != 0)) return -20;
  if (!(p.BN == 64 || p.BN == 128 || p.BN == 256) || p.Cy % p.BN != 0) return -21;
  if (p.ntaps < 1 || p.ntaps > SGGAN_MAX_TAPS || p.ksplit < 1) return -22;
  L->p = p;
  L->stages = stages_for(p.BN);
  L->tmem_cols = tmem_cols_for(p.BN);
  L->smem = size_t(L->stages) * (kABytes + p.BN * 128) + 1024;
  L->grid_x = (p.x_pair ? 1 : p.Cx / 128) * p.ntaps;
  L->grid_y = p.Cy / p.BN;
  L->grid_z = p.ksplit;
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
  wgrad_gemm_tc_kernel<<<grid, kThreads, L.smem, st>>>(L.tmX, L.tmY, L.p, L.stages, L.tmem_cols);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -4000 - int(e);
}

int read_tc_watchdog() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_watchdog_flag, sizeof(int));
  return v;
}

}  // namespace sggan
