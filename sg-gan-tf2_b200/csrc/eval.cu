// eval.cu -- the per-epoch evaluation that follows the training path in the reference (metric.py:18-47,71-77;
// model.py:307-378), as two small integer kernels: the argmax-over-RGB label adapter and the confusion matrix.
// HBM-bound byte / integer work: coalesced reads, a shared-memory histogram per block, 64-bit global counters.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/sggan.h"

namespace sggan {

// metric.py:71-77 scores_seg_fake: labels = argmax over the 3 channels of (255 * x).astype(uint8), first maximum wins,
// with the reference's transpose(0, 3, 2, 1): image [B,H,W,3] -> labels [B,W,H].
__global__ void __launch_bounds__(256) rgb_argmax_labels_kernel(const float* __restrict__ img, int B, int H, int W,
                                                                int32_t* __restrict__ labels) {
  const int64_t tot = int64_t(B) * H * W;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < tot; idx += int64_t(gridDim.x) * blockDim.x) {
    const int b = int(idx / (int64_t(H) * W));
    const int r = int(idx - int64_t(b) * H * W), i = r / W, j = r - i * W;
    const float* p = img + idx * 3;
    // float -> uint8 like numpy's astype: truncate toward zero, wrap modulo 256
    const uint8_t c0 = uint8_t(int(255.f * p[0])), c1 = uint8_t(int(255.f * p[1])), c2 = uint8_t(int(255.f * p[2]));
    int best = 0;
    uint8_t bv = c0;
    if (c1 > bv) { best = 1; bv = c1; }
    if (c2 > bv) best = 2;
    labels[(int64_t(b) * W + j) * H + i] = best;
  }
}

// metric.py:18-24 _fast_hist: hist[t * n_class + p] += 1 for every element with 0 <= t < n_class.
__global__ void __launch_bounds__(256) fast_hist_kernel(const int32_t* __restrict__ lt, const int32_t* __restrict__ lp, int64_t n,
                                                        int n_class, unsigned long long* hist) {
  extern __shared__ unsigned int sh[];
  const int bins = n_class * n_class;
  for (int t = threadIdx.x; t < bins; t += blockDim.x) sh[t] = 0u;
  __syncthreads();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int t = lt[i], q = lp[i];
    if (t >= 0 && t < n_class && q >= 0 && q < n_class) atomicAdd(&sh[t * n_class + q], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < bins; t += blockDim.x)
    if (sh[t]) atomicAdd(hist + t, (unsigned long long)sh[t]);
}

}  // namespace sggan

extern "C" int sggan_rgb_argmax_labels(const float* img, int32_t* labels, int B, int H, int W, void* stream) {
  const int64_t tot = int64_t(B) * H * W;
  int blocks = int((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  sggan::rgb_argmax_labels_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img, B, H, W, labels);
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}
extern "C" int sggan_fast_hist(const int32_t* label_true, const int32_t* label_pred, int64_t n, int n_class, int64_t* hist,
                               void* stream) {
  if (n_class < 1 || n_class > 96) return SGGAN_E_INVALID;  // n_class^2 counters in shared memory
  int blocks = int((n + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  sggan::fast_hist_kernel<<<blocks, 256, size_t(n_class) * n_class * sizeof(unsigned int), (cudaStream_t)stream>>>(
      label_true, label_pred, n, n_class, reinterpret_cast<unsigned long long*>(hist));
  return cudaGetLastError() == cudaSuccess ? 0 : SGGAN_E_CUDA;
}

// Host helper of the checkpoint reader / writer (tf_checkpoint.py): CRC-32C (Castagnoli, reflected 0x82F63B78) as the TF
// tensor-bundle format uses it for every table block and every tensor.  crc = value to extend (0 to start).
extern "C" uint32_t sggan_crc32c(const void* data, size_t n, uint32_t crc) {
  static uint32_t table[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xFF];
    ready = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  while (n >= 8) {  // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xFF] ^ table[6][(lo >> 8) & 0xFF] ^ table[5][(lo >> 16) & 0xFF] ^ table[4][lo >> 24] ^
        table[3][hi & 0xFF] ^ table[2][(hi >> 8) & 0xFF] ^ table[1][(hi >> 16) & 0xFF] ^ table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = table[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return ~c;
}
