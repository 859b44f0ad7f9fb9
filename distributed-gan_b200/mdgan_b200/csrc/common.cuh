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

#define MDGAN_LAUNCH(...)                          \
  do {                                             \
    cudaError_t e__ = mdgan::launch(__VA_ARGS__);  \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

#define MDGAN_CUDA(call)                           \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

namespace mdgan {

// Programmatic dependent launch (PDL): every kernel of the library calls pdl_prologue_done() once its CTA-local set-up
// is finished (so the next kernel in the stream may start scheduling CTAs into SMs that drain) and pdl_wait() before
// its first global-memory access (blocks until the preceding grids have completed and flushed).  Launches carry the
// programmatic-stream-serialization attribute when MDGAN_PDL=1 (off by default: on B200 the captured step measured
// 0.699 ms with it and 0.693 ms without, profiles/); without the attribute both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }
#endif

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 2D fp32 row-major matrix [rows, cols] -> tensor map with box [box_rows, 32 cols] and 128B swizzle.
// Returns 0 on success.  Maps are cached per (ptr, rows, cols, box_rows).
int get_tmap_2d_f32(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMap* out);

// 4-D im2col-tile map over an NHWC fp32 tensor (see common.cu).  Cached per argument tuple.
int get_tmap_im2col_f32(const void* ptr, int n_img, int Hs, int Ws, int C, int bn, int bh, int bw, int si,
                        CUtensorMap* out);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, so a process
// that drives several GPUs must set it on each of them.  Keyed by the kernel's ADDRESS (instantiations of one template
// share a function type).
cudaError_t configure_smem_once_impl(const void* kernel, int bytes);
template <typename K>
inline cudaError_t configure_smem_once(K kernel, int bytes) {
  return configure_smem_once_impl(reinterpret_cast<const void*>(kernel), bytes);
}

}  // namespace mdgan
