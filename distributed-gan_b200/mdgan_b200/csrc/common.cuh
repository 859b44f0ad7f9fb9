// Shared host-side helpers for the C-ABI launchers (error codes, TMA tensor-map cache).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

// C-ABI error codes (include/mdgan_b200.h): 0 = ok, > 0 = cudaError_t, < 0 = own code.
#define MDGAN_ERR_BAD_ARG (-1)
#define MDGAN_ERR_UNSUPPORTED (-2)
#define MDGAN_ERR_DRIVER (-3)

#define MDGAN_CHECK_LAUNCH()                       \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

#define MDGAN_CUDA(call)                           \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

namespace mdgan {

// 2D fp32 row-major matrix [rows, cols] -> tensor map with box [box_rows, 32 cols] and 128B swizzle.
// Returns 0 on success.  Maps are cached per (ptr, rows, cols, box_rows).
int get_tmap_2d_f32(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMap* out);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mdgan
