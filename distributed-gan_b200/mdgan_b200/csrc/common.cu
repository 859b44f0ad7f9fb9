// Host-side helpers shared by the launchers: TMA tensor-map creation (driver entry point fetched at run time so
// the library does not link libcuda) with a small cache, plus library-level C-ABI utilities.
#include "common.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace mdgan {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int get_tmap_2d_f32(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMap* out) {
  using Key = std::tuple<const void*, uint64_t, uint64_t, uint32_t>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key key{ptr, rows, cols, box_rows};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MDGAN_ERR_DRIVER;
  if (cols % 32 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || box_rows == 0 || box_rows > 256)
    return MDGAN_ERR_BAD_ARG;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return MDGAN_ERR_DRIVER;
  cache[key] = m;
  *out = m;
  return 0;
}

// NHWC fp32 activation tensor [n_img][Hs][Ws][C] -> 4-D tensor map whose box is one 128-row im2col tile of a single
// tap and 32-channel chunk: {32 channels, bw positions (step si), bh positions (step si), bn images}, 128B swizzle,
// out-of-bounds (padding) elements read as zero.  boxDim = count * elementStride as cuTensorMapEncodeTiled specifies.
int get_tmap_im2col_f32(const void* ptr, int n_img, int Hs, int Ws, int C, int bn, int bh, int bw, int si,
                        CUtensorMap* out) {
  using Key = std::tuple<const void*, int, int, int, int, int, int, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key key{ptr, n_img, Hs, Ws, C, bn, bh, bw, si};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MDGAN_ERR_DRIVER;
  if (C % 32 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || bw * si > 256 || bh * si > 256 || bn > 256 ||
      bn * bh * bw > 128 || bn * bh * bw < 1)
    return MDGAN_ERR_BAD_ARG;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)n_img};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 4, (cuuint64_t)Ws * C * 4, (cuuint64_t)Hs * Ws * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)(bw * si), (cuuint32_t)(bh * si), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)si, (cuuint32_t)si, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return MDGAN_ERR_DRIVER;
  cache[key] = m;
  *out = m;
  return 0;
}

cudaError_t configure_smem_once_impl(const void* kernel, int bytes) {
  static std::map<std::pair<const void*, int>, int> done;
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find({kernel, dev});
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[{kernel, dev}] = bytes;
  return e;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MDGAN_PDL");  // opt-in: measured on B200 (round 1) it does not shorten the captured step
    return e && e[0] == '1';
  }();
  return on;
}

}  // namespace mdgan

extern "C" int mdgan_abi_version(void) { return 2; }

// Which SM architecture the loaded device code targets and whether the current device can run it.
// Returns 0 when the current CUDA device is compute capability 10.x, MDGAN_ERR_UNSUPPORTED otherwise.
extern "C" int mdgan_check_device(void) {
  int dev = 0;
  MDGAN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MDGAN_CUDA(cudaGetDeviceProperties(&prop, dev));
  return prop.major == 10 ? 0 : MDGAN_ERR_UNSUPPORTED;
}
