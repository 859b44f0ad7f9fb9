// CUDA-core fp32 kernels for the reference's MLP plugin (datasets/MNIST.py:74-120: four Linear layers per net,
// LeakyReLU(0.2), always-active dropout(0.3) in the discriminator, tanh / sigmoid outputs).  One iteration of that model
// is ~2.4 GFLOP at b = 64 (F_D = 1.29, F_G = 1.13 GFLOP) in GEMMs of 64..128 rows: a 128-row tensor-core
// tile would be half padding and tf32x3 would buy nothing over plain fp32 FMA, so the Linear layers run as a
// register-blocked fp32 SGEMM with the layer's elementwise tail fused into the epilogue:
//
//   forward   y = x W^T + b -> LeakyReLU / tanh -> dropout (keep mask drawn on the HOST from the worker's torch RNG in
//             the reference's order, so the masks are the reference's masks bit for bit: F.dropout on the CPU is
//             noise = empty_like(x).bernoulli_(1 - p).div_(1 - p); x * noise)
//   data grad dx = dy W, times the previous layer's dropout mask and LeakyReLU gate (torch's order: dropout backward,
//             then leaky_relu backward)
//   weight grad dW = dy^T x (PyTorch [out, in] layout, written straight into the flat gradient buffer), db = column sums
//   head      Linear(L -> 1) + sigmoid + BCELoss(mean) and its backward (worker.py:197-206,220-227), bias included.
//
// All sums run in a fixed order (k ascending per output element), so results are bitwise repeatable.
#include <cstdlib>

#include "common.cuh"

namespace mdgan {

constexpr int kSgBK = 16;

// C[m][n] (row-major, ld = N) = epilogue(sum_k A(m,k) * B(k,n)),  A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs].
// BM x BN output tile per CTA, (BM/4) x (BN/4) threads, a 4 x 4 register tile per thread (rows 4 ty .. 4 ty + 3, columns
// 4 tx .. 4 tx + 3), K stepped by 16 through shared memory ([k][m] / [k][n], rows padded to a multiple of 16 bytes so
// that a thread fetches its four A and four B values of a k with ONE 128-bit load each: 2 LDS per 16 FMA).  The next
// K tile's global loads are issued into registers before the current tile is consumed, so their latency overlaps the
// FMAs.  The layers of this model family have 64 .. 128 rows (b, 2b): tiles of 32 x 32 (64 threads) keep ~130 CTAs in
// flight on such shapes where 64 x 64 tiles would occupy 32 of the 148 SMs; 64 x 64 is used once it fills the machine
// (the weight gradients).  The sum over k is sequential per output element in both configurations (k ascending, one
// fmaf chain), so the result does not depend on the tile size.
// Epilogue order: + bias[n] -> act (2 LeakyReLU, 3 tanh) -> keep mask (mask ? v * mask_scale : 0) -> gate
// (gate > 0 ? v : v * gate_slope) -> (+ C if accumulate).
template <int BM, int BN>
__global__ void __launch_bounds__((BM / 4) * (BN / 4))
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
             long long a_rs, long long a_cs, long long b_rs, long long b_cs, const float* __restrict__ bias, int act,
             float slope, const unsigned char* __restrict__ mask, float mask_scale, const float* __restrict__ gate,
             float gate_slope, int accumulate) {
  constexpr int THREADS = (BM / 4) * (BN / 4), TXN = BN / 4;
  constexpr int LA = BM * kSgBK / THREADS, LB = BN * kSgBK / THREADS;   // global loads per thread and K tile
  static_assert(BM % 4 == 0 && BN % 4 == 0 && (BM * kSgBK) % THREADS == 0 && (BN * kSgBK) % THREADS == 0, "tile shape");
  __shared__ __align__(16) float As[kSgBK][BM + 4];
  __shared__ __align__(16) float Bs[kSgBK][BN + 4];
  pdl_enter();
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // the index that is contiguous in memory varies fastest over the threads
  const bool a_k_fast = a_cs == 1, b_k_fast = b_rs == 1 && b_cs != 1;
  float ra[LA], rb[LB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int idx = tid + THREADS * u;
      const int am = a_k_fast ? idx / kSgBK : idx % BM, ak = a_k_fast ? idx % kSgBK : idx / BM;
      const int gm = m0 + am, gk = k0 + ak;
      ra[u] = (gm < M && gk < K) ? __ldg(A + gm * a_rs + gk * a_cs) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int idx = tid + THREADS * u;
      const int bn = b_k_fast ? idx / kSgBK : idx % BN, bk = b_k_fast ? idx % kSgBK : idx / BN;
      const int gn = n0 + bn, gk = k0 + bk;
      rb[u] = (gn < N && gk < K) ? __ldg(B + gk * b_rs + gn * b_cs) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int idx = tid + THREADS * u;
      const int am = a_k_fast ? idx / kSgBK : idx % BM, ak = a_k_fast ? idx % kSgBK : idx / BM;
      As[ak][am] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int idx = tid + THREADS * u;
      const int bn = b_k_fast ? idx / kSgBK : idx % BN, bk = b_k_fast ? idx % kSgBK : idx / BN;
      Bs[bk][bn] = rb[u];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kSgBK) {
    stash();
    __syncthreads();
    if (k0 + kSgBK < K) fetch(k0 + kSgBK);
#pragma unroll
    for (int kk = 0; kk < kSgBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const long long o = (long long)m * N + n;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (act == 2) v = v > 0.f ? v : v * slope;
      else if (act == 3) v = tanhf(v);
      if (mask) v = mask[o] ? v * mask_scale : 0.f;
      if (gate) v = gate[o] > 0.f ? v : v * gate_slope;
      if (accumulate) v += C[o];
      C[o] = v;
    }
  }
}

// The 64 .. 128-row layers again, with K split over the warps of the CTA: 32 x 32 output tile, 256 threads = four groups
// of 64 (8 x 8 threads, 4 x 4 register tile each); a K tile of 64 is loaded by all 256 threads (8 + 8 values each,
// prefetched into registers one tile ahead) and group g multiplies its 16 k of the tile.  Against sgemm_kernel<32,32>
// that is four times the warps per SM and four times the bytes in flight per exposed global-memory latency, and a
// quarter as many of those exposures (784 / 64 = 13 tiles instead of 49).  The four partial tiles are combined through
// shared memory in a fixed order ((g0 + g1) + g2) + g3 by group 0, which also runs the epilogue: bitwise repeatable; the
// last bits differ from the single-chain sum of the other configurations.  MDGAN_SGEMM_SPLITK=0 selects
// sgemm_kernel<32,32> instead.
constexpr int kSkBM = 32, kSkBN = 32, kSkBK = 64, kSkGroups = 4, kSkThreads = 256;

__global__ void __launch_bounds__(kSkThreads)
sgemm_splitk_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                    long long a_rs, long long a_cs, long long b_rs, long long b_cs, const float* __restrict__ bias, int act,
                    float slope, const unsigned char* __restrict__ mask, float mask_scale, const float* __restrict__ gate,
                    float gate_slope, int accumulate) {
  constexpr int LA = kSkBM * kSkBK / kSkThreads, LB = kSkBN * kSkBK / kSkThreads, KG = kSkBK / kSkGroups;
  __shared__ __align__(16) float As[kSkBK][kSkBM + 4];
  __shared__ __align__(16) float Bs[kSkBK][kSkBN + 4];
  __shared__ float red[kSkGroups - 1][64][17];
  pdl_enter();
  const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
  const int tx = t & 7, ty = t >> 3;
  const int m0 = blockIdx.y * kSkBM, n0 = blockIdx.x * kSkBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_k_fast = a_cs == 1, b_k_fast = b_rs == 1 && b_cs != 1;
  float ra[LA], rb[LB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int idx = tid + kSkThreads * u;
      const int am = a_k_fast ? idx / kSkBK : idx % kSkBM, ak = a_k_fast ? idx % kSkBK : idx / kSkBM;
      const int gm = m0 + am, gk = k0 + ak;
      ra[u] = (gm < M && gk < K) ? __ldg(A + gm * a_rs + gk * a_cs) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int idx = tid + kSkThreads * u;
      const int bn = b_k_fast ? idx / kSkBK : idx % kSkBN, bk = b_k_fast ? idx % kSkBK : idx / kSkBN;
      const int gn = n0 + bn, gk = k0 + bk;
      rb[u] = (gn < N && gk < K) ? __ldg(B + gk * b_rs + gn * b_cs) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int idx = tid + kSkThreads * u;
      const int am = a_k_fast ? idx / kSkBK : idx % kSkBM, ak = a_k_fast ? idx % kSkBK : idx / kSkBM;
      As[ak][am] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int idx = tid + kSkThreads * u;
      const int bn = b_k_fast ? idx / kSkBK : idx % kSkBN, bk = b_k_fast ? idx % kSkBK : idx / kSkBN;
      Bs[bk][bn] = rb[u];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kSkBK) {
    stash();
    __syncthreads();
    if (k0 + kSkBK < K) fetch(k0 + kSkBK);
#pragma unroll
    for (int kk = 0; kk < KG; ++kk) {
      const int k = g * KG + kk;
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (g > 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) red[g - 1][t][e] = acc[e >> 2][e & 3];
  }
  __syncthreads();
  if (g > 0) return;
#pragma unroll
  for (int gg = 0; gg < kSkGroups - 1; ++gg)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e >> 2][e & 3] += red[gg][t][e];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const long long o = (long long)m * N + n;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (act == 2) v = v > 0.f ? v : v * slope;
      else if (act == 3) v = tanhf(v);
      if (mask) v = mask[o] ? v * mask_scale : 0.f;
      if (gate) v = gate[o] > 0.f ? v : v * gate_slope;
      if (accumulate) v += C[o];
      C[o] = v;
    }
  }
}

// out[n] = sum_m x[m][n] (m ascending): bias gradients.
__global__ void col_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N) {
  pdl_enter();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
#pragma unroll 8
  for (int m = 0; m < M; ++m) acc += x[(long long)m * N + n];   // loads independent, adds in order
  out[n] = acc;
}

// logit[n] = <a[n, :], w> + bias; p = sigmoid(logit); per-sample BCE term (log clamp at -100); dlogit = dBCE/dlogit
// including the 1/b of the mean.  One 256-thread block per sample; the last block to finish reduces the terms in a
// fixed order: loss[g] = mean over the b samples of pass g (label[g]), loss[G] = their sum.  Same arithmetic as
// head_fwd_kernel (elementwise.cu) with the bias of nn.Linear added.
__global__ void __launch_bounds__(256)
linear_head_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ label, float* __restrict__ prob, float* __restrict__ loss_terms,
                       float* __restrict__ dlogit, float* __restrict__ loss, unsigned int* __restrict__ counter, int b,
                       int G, int L) {
  __shared__ float red[8];
  __shared__ bool last;
  pdl_enter();
  const int n = blockIdx.x;
  const float* row = a + (long long)n * L;
  float acc = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) acc = fmaf(row[i], __ldg(w + i), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float logit = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) logit += red[i];
    if (bias) logit += bias[0];
    const float y = label[n / b];
    const float p = 1.f / (1.f + expf(-logit));
    const float lp = fmaxf(logf(p), -100.f);
    const float l1p = fmaxf(log1pf(-p), -100.f);
    prob[n] = p;
    loss_terms[n] = (y - 1.f) * l1p - y * lp;
    const float pq = (1.f - p) * p;
    dlogit[n] = ((p - y) / fmaxf(pq, 1e-12f)) * (1.f / (float)b) * pq;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) *counter = 0;
  float total = 0.f;
  for (int g = 0; g < G; ++g) {
    float t = 0.f;
    for (int i = threadIdx.x; i < b; i += 256) t += __ldcg(loss_terms + g * b + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float sgl = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) sgl += red[wv];
      sgl /= (float)b;
      loss[g] = sgl;
      total += sgl;
    }
  }
  if (threadIdx.x == 0) loss[G] = total;
}

// da[n][l] = dlogit[n] * w[l], then the dropout mask and the LeakyReLU gate of the layer that produced a (a is its
// post-dropout output: where the mask kept the element, sign(a) is the sign of the pre-activation);
// dw[l] = sum_n dlogit[n] * a[n][l], dbias = sum_n dlogit[n] (both optional, n ascending).
__global__ void linear_head_bwd_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                       const float* __restrict__ dlogit, const unsigned char* __restrict__ mask,
                                       float mask_scale, float gate_slope, float* __restrict__ da, float* __restrict__ dw,
                                       float* __restrict__ dbias, int n_total, int L) {
  pdl_enter();
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const float wv = w[l];
  float acc = 0.f, accb = 0.f;
#pragma unroll 4
  for (int n = 0; n < n_total; ++n) {
    const long long o = (long long)n * L + l;
    const float d = dlogit[n], av = a[o];
    float v = d * wv;
    if (mask) v = mask[o] ? v * mask_scale : 0.f;
    v = av > 0.f ? v : v * gate_slope;
    da[o] = v;
    acc = fmaf(d, av, acc);
    accb += d;
  }
  if (dw) dw[l] = acc;
  if (dbias && l == 0) dbias[0] = accb;
}

}  // namespace mdgan

using namespace mdgan;

extern "C" int mdgan_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int a_rs, int a_cs, int b_rs,
                           int b_cs, const float* bias, int act, float slope, const unsigned char* mask, float mask_scale,
                           const float* gate, float gate_slope, int accumulate, void* stream) {
  if (!A || !B || !C || M < 1 || N < 1 || K < 1) return MDGAN_ERR_BAD_ARG;
  if (act != 0 && act != 2 && act != 3) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const long long big_tiles = (long long)ceil_div(N, 64) * ceil_div(M, 64);
  if (big_tiles >= 148) {   // 64 x 64 tiles fill the machine
    const dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
    if (grid.y > 65535u) return MDGAN_ERR_UNSUPPORTED;
    MDGAN_LAUNCH((sgemm_kernel<64, 64>), grid, dim3(256), 0, st, A, B, C, M, N, K, (long long)a_rs, (long long)a_cs,
                 (long long)b_rs, (long long)b_cs, bias, act, slope, mask, mask_scale, gate, gate_slope, accumulate);
  } else {
    static const bool split_k = [] {
      const char* e = getenv("MDGAN_SGEMM_SPLITK");
      return !(e && e[0] == '0');
    }();
    const dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32));
    if (grid.y > 65535u) return MDGAN_ERR_UNSUPPORTED;
    if (split_k)
      MDGAN_LAUNCH(sgemm_splitk_kernel, grid, dim3(kSkThreads), 0, st, A, B, C, M, N, K, (long long)a_rs, (long long)a_cs,
                   (long long)b_rs, (long long)b_cs, bias, act, slope, mask, mask_scale, gate, gate_slope, accumulate);
    else
      MDGAN_LAUNCH((sgemm_kernel<32, 32>), grid, dim3(64), 0, st, A, B, C, M, N, K, (long long)a_rs, (long long)a_cs,
                   (long long)b_rs, (long long)b_cs, bias, act, slope, mask, mask_scale, gate, gate_slope, accumulate);
  }
  return 0;
}

extern "C" int mdgan_col_sum(const float* x, float* out, int M, int N, void* stream) {
  if (!x || !out || M < 1 || N < 1) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(col_sum_kernel, dim3((unsigned)ceil_div(N, 128)), dim3(128), 0, (cudaStream_t)stream, x, out, M, N);
  return 0;
}

extern "C" int mdgan_linear_head_forward(const float* a, const float* w, const float* bias, const float* label,
                                         float* prob, float* loss_terms, float* dlogit, float* loss, void* counter, int G,
                                         int b, int L, void* stream) {
  if (!a || !w || !label || !prob || !loss_terms || !dlogit || !loss || !counter || G < 1 || b < 1 || L < 1)
    return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(linear_head_fwd_kernel, dim3((unsigned)(G * b)), dim3(256), 0, (cudaStream_t)stream, a, w, bias, label, prob,
               loss_terms, dlogit, loss, (unsigned int*)counter, b, G, L);
  return 0;
}

extern "C" int mdgan_linear_head_backward(const float* a, const float* w, const float* dlogit, const unsigned char* mask,
                                          float mask_scale, float gate_slope, float* da, float* dw, float* dbias,
                                          int n_total, int L, void* stream) {
  if (!a || !w || !dlogit || !da || n_total < 1 || L < 1) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(linear_head_bwd_kernel, dim3((unsigned)ceil_div(L, 128)), dim3(128), 0, (cudaStream_t)stream, a, w, dlogit,
               mask, mask_scale, gate_slope, da, dw, dbias, n_total, L);
  return 0;
}
