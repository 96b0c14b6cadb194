// tmap.h -- host-side creation of TMA tensor maps (cuTensorMapEncodeTiled), resolved through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sggan {

// bf16 tensor [d2][d1][d0] (d0 innermost, contiguous), byte strides s1 (between d1 rows) and s2
// (between d2 slices).  Box = box0 x box1 x 1, SWIZZLE_128B, out-of-bounds reads return zeros.
// Returns 0 on success, a CUresult / cudaError code otherwise.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                      uint64_t s2, uint32_t box0, uint32_t box1);

// The same with a selectable element type (f32 != 0: fp32 elements, for the tf32 convolution; box0 then counts fp32
// elements, 32 per 128-byte swizzled row).
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                 uint32_t box0, uint32_t box1, int f32);
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t s1, uint32_t box0, uint32_t box1,
                 int f32);

// 2-D variant [d1][d0].
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t s1, uint32_t box0,
                      uint32_t box1);

}  // namespace sggan
