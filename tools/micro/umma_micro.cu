// Micro-benchmark + layout check of tcgen05.mma on sm_100a: cycles per MMA for SS mode (A, B in shared memory) vs
// TS mode (A in tensor memory) at several N, same / rotating accumulators; and a numerical check that the TS-mode A
// layout is lane = row m, column = k (32-bit per tf32 element).  Build: make -C tools/micro ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../distributed-gan_b200/mdgan_b200/csrc/ptx.cuh"

using namespace mdgan;

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

// MODE 0: SS tf32, 1: TS tf32, 2: SS bf16.  The issue loop is unrolled x8 with loop-invariant descriptors so that it
// measures the tensor core, not the scalar instructions of the issuing thread (the first version of this benchmark
// measured ~190 clk per MMA at every N: a single thread needs that long for a modulo + two descriptor builds).
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int nacc, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
    // MODE 3: both operands MN-major (SWIZZLE_128B_BASE32B, the weight-gradient layout); MODE 4: A from TMEM, B MN-major
    const uint32_t idesc = MODE == 2 ? idesc_bf16(128, N) : MODE == 3 ? make_idesc_tf32(128, N, 1, 1)
                           : MODE == 4 ? make_idesc_tf32(128, N, 0, 1) : make_idesc_tf32(128, N, 0, 0);
    const uint32_t a_tmem = tb + 448;
    uint64_t da[4], db[4];
    uint32_t d[2] = {tb, tb + (nacc > 1 ? (uint32_t)N : 0u)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (MODE >= 3) {  // 8 pixels = two 512-byte k-atoms; LBO = 4096 (32-channel groups), SBO = 512
        da[k] = make_smem_desc_sw128(a_addr + k * 1024, 4096, 512, 1);
        db[k] = make_smem_desc_sw128(b_addr + k * 1024, 4096, 512, 1);
      } else {
        da[k] = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
        db[k] = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
      }
    }
    long long t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (MODE == 0 || MODE == 3) umma_tf32(d[u & 1], da[u & 3], db[u & 3], idesc, 1u);
        else if (MODE == 1 || MODE == 4) umma_tf32_ts(d[u & 1], a_tmem + (u & 3) * 8, db[u & 3], idesc, 1u);
        else umma_f16_ss(d[u & 1], da[u & 3], db[u & 3], idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tb); }
}

// TS-mode layout check: D[128 x 64] = A[128 x 8] * B[64 x 8]^T with A written to TMEM by tcgen05.st (lane = m, col = k)
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const float* A, const float* B, float* D, int ts) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  float* a_s = reinterpret_cast<float*>(smem);            // [128 rows][32 floats] SW128 K-major (only k < 8 used)
  float* b_s = reinterpret_cast<float*>(smem + 16384);    // [64 rows][32 floats]
  for (int i = threadIdx.x; i < (16384 + 8192) / 4; i += blockDim.x) a_s[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i / 8, k = i % 8;
    a_s[r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3))] = A[r * 8 + k];
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i / 8, k = i % 8;
    b_s[r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3))] = B[r * 8 + k];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<128>(&tptr);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    float v[8];
    for (int k = 0; k < 8; ++k) v[k] = A[(warp * 32 + lane) * 8 + k];
    tmem_st_x8(tb + (static_cast<uint32_t>(warp * 32) << 16) + 64, v);  // A at columns 64..71
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after_sync();
    const uint64_t da = make_smem_desc_sw128(smem_u32(a_s), 16, 1024);
    const uint64_t db = make_smem_desc_sw128(smem_u32(b_s), 16, 1024);
    const uint32_t idt = make_idesc_tf32(128, 64, 0, 0);
    if (ts) umma_tf32_ts(tb, tb + 64, db, idt, 0u);
    else umma_tf32(tb, da, db, idt, 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < 64; c += 16) {
    float v[16];
    tmem_ld_x16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 64 + c + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<128>(tb); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(ts_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // ---- numerical check
  std::vector<float> A(128 * 8), B(64 * 8), D(128 * 64);
  for (int i = 0; i < 128 * 8; ++i) A[i] = (float)((i * 7) % 13 - 6);
  for (int i = 0; i < 64 * 8; ++i) B[i] = (float)((i * 5) % 11 - 5);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int ts = 0; ts < 2; ++ts) {
    cudaMemset(dD, 0, D.size() * 4);
    ts_check_kernel<<<1, 128, smem>>>(dA, dB, dD, ts);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        for (int k = 0; k < 8; ++k) ref += (double)A[m * 8 + k] * B[n * 8 + k];
        worst = fmax(worst, fabs(ref - D[m * 64 + n]));
      }
    printf("%s-mode check: cuda=%s max abs err %.3g  (D[0][0]=%g D[5][7]=%g)\n", ts ? "TS" : "SS", cudaGetErrorString(e), worst, D[0], D[5 * 64 + 7]);
  }
  // ---- rates
  const int reps = 4096;
  const char* names[] = {"SS tf32", "TS tf32", "SS bf16", "SS tf32 MN-major A+B", "TS tf32 MN-major B"};
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 5; ++mode)
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * N > 448) continue;
        if (mode == 0) rate_kernel<0><<<148, 128, smem>>>(N, nacc, reps, d_out);
        else if (mode == 1) rate_kernel<1><<<148, 128, smem>>>(N, nacc, reps, d_out);
        else if (mode == 2) rate_kernel<2><<<148, 128, smem>>>(N, nacc, reps, d_out);
        else if (mode == 3) rate_kernel<3><<<148, 128, smem>>>(N, nacc, reps, d_out);
        else rate_kernel<4><<<148, 128, smem>>>(N, nacc, reps, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
        printf("%s N=%3d nacc=%d: %7.1f clk/MMA (floor %d)  %s\n", names[mode], N, nacc, (double)c / reps, N / 2,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
