// CUDA-core kernels for the image-side "thin" layers (1 or 3 image channels; < 5 % of the step's FLOPs,
// bandwidth-bound).  The image side is NCHW fp32 (the layout of every tensor that crosses the
// reference's actor boundary: real batches, generated batches, feedback), the feature side NHWC.
//
//   thin_down  : out[n,i,j,co] = sum_{c,kh,kw} img[n,c,2i-1+kh,2j-1+kw] * W[co][c][kh][kw]  (+ LeakyReLU)
//                = first discriminator Conv2d(3->64,k4,s2,p1) forward (CIFAR10.py:85, CelebA.py:78) and the
//                data-gradient of the last generator ConvTranspose2d(->3) (W is then [ci][co][kh][kw]).
//   thin_wgrad : dW[c1][c2][kh][kw] = sum_p feat[p,c1] * img[n,c2,2i-1+kh,2j-1+kw]
//                = weight gradient of both of those layers.
#include "common.cuh"

namespace mdgan {

__device__ __forceinline__ float thin_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// block: 256 threads = 64 output pixels x 4 channel groups; N in {64, 128} (N/4 channels per thread).
template <int CI, int NPT>
__global__ void __launch_bounds__(256) thin_down_kernel(const float* __restrict__ img, const float* __restrict__ W,
                                                        float* __restrict__ out, int n_img, int Hi, int Wi, int N,
                                                        int act, float slope, int round_tf32) {
  extern __shared__ float w_s[];  // [CI*16][N]
  const int Ho = Hi >> 1, Wo = Wi >> 1;
  for (int i = threadIdx.x; i < CI * 16 * N; i += blockDim.x) {
    const int co = i % N, k = i / N;
    w_s[i] = W[co * CI * 16 + k];
  }
  __syncthreads();
  const int cg = threadIdx.x & 3;
  const long long pix = blockIdx.x * 64LL + (threadIdx.x >> 2);
  const long long P = (long long)n_img * Ho * Wo;
  if (pix >= P) return;
  const int n = pix / (Ho * Wo);
  const int rem = pix - (long long)n * Ho * Wo;
  const int oi = rem / Wo, oj = rem - oi * Wo;
  float acc[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) acc[j] = 0.f;
#pragma unroll
  for (int c = 0; c < CI; ++c) {
    const float* plane = img + ((long long)n * CI + c) * Hi * Wi;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int ih = 2 * oi - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int iw = 2 * oj - 1 + kw;
        float v = 0.f;
        if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) v = __ldg(plane + ih * Wi + iw);
        const float* wr = w_s + (c * 16 + kh * 4 + kw) * N + cg * NPT;
#pragma unroll
        for (int j = 0; j < NPT; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
      }
    }
  }
  float* o = out + pix * N + cg * NPT;
#pragma unroll
  for (int j = 0; j < NPT; j += 4) {
    float4 r;
    float* rp = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float x = acc[j + t];
      if (act == 2) x = x > 0.f ? x : x * slope;
      if (round_tf32) x = thin_to_tf32(x);
      rp[t] = x;
    }
    *reinterpret_cast<float4*>(o + j) = r;
  }
}

// Each block walks pixel tiles of 32; thread owns channel c1 = tid % C1 and one slice of the CI*16 taps.
template <int CI, int C1>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ feat, const float* __restrict__ img,
                                                         float* __restrict__ partial, int n_img, int Hl, int Wl) {
  constexpr int K = CI * 16;
  constexpr int SLICES = 256 / C1;  // 4 (C1=64) or 2 (C1=128)
  constexpr int KPT = K / SLICES;   // taps per thread
  static_assert(K % SLICES == 0, "tap slices");
  __shared__ float f_s[32][C1];
  __shared__ float p_s[32][K];
  const int Hi = 2 * Hl, Wi = 2 * Wl;
  const long long P = (long long)n_img * Hl * Wl;
  const int c1 = threadIdx.x % C1, slice = threadIdx.x / C1;
  float acc[KPT];
#pragma unroll
  for (int j = 0; j < KPT; ++j) acc[j] = 0.f;
  for (long long p0 = blockIdx.x * 32LL; p0 < P; p0 += gridDim.x * 32LL) {
    for (int i = threadIdx.x; i < 32 * C1; i += 256) {
      const int r = i / C1, c = i % C1;
      f_s[r][c] = (p0 + r < P) ? feat[(p0 + r) * C1 + c] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * K; i += 256) {
      const int r = i / K, k = i % K;
      float v = 0.f;
      if (p0 + r < P) {
        const long long pix = p0 + r;
        const int n = pix / (Hl * Wl);
        const int rem = pix - (long long)n * Hl * Wl;
        const int oi = rem / Wl, oj = rem - oi * Wl;
        const int c = k >> 4, kh = (k >> 2) & 3, kw = k & 3;
        const int ih = 2 * oi - 1 + kh, iw = 2 * oj - 1 + kw;
        if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) v = img[(((long long)n * CI + c) * Hi + ih) * Wi + iw];
      }
      p_s[r][k] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float f = f_s[r][c1];
#pragma unroll
      for (int j = 0; j < KPT; ++j) acc[j] = fmaf(f, p_s[r][slice * KPT + j], acc[j]);
    }
    __syncthreads();
  }
  float* o = partial + (long long)blockIdx.x * C1 * K + (long long)c1 * K + slice * KPT;
#pragma unroll
  for (int j = 0; j < KPT; ++j) o[j] = acc[j];
}

// thin_up: out[n, o, 2i+ph, 2j+pw] (NCHW, o < N in {1, 3}) = sum_{c, a, b} src[n, i+ph-a, j+pw-b, c] * W[c][o][kh][kw],
// kh = (1-ph)+2a, kw = (1-pw)+2b: the 4-phase form of ConvTranspose2d(C -> N, k4, s2, p1) on an NHWC source, which is
// also the data gradient of Conv2d(N -> C, k4, s2, p1) (W is then the conv weight [C][N][4][4]).  One thread per
// low-resolution position computes its 2x2 output block for all N channels from the 3x3 source neighbourhood
// (fp32 FMAs, weights broadcast from shared memory); optional tanh and accumulate-into-out epilogues.
template <int N>
__global__ void __launch_bounds__(128) thin_up_kernel(const float* __restrict__ src, const float* __restrict__ W,
                                                      float* __restrict__ out, int n_img, int H, int Wd, int C,
                                                      int act_tanh, int accumulate) {
  extern __shared__ float w_s[];  // [C][N][16]
  for (int i = threadIdx.x; i < C * N * 16; i += blockDim.x) w_s[i] = W[i];
  __syncthreads();
  const long long P = (long long)n_img * H * Wd;
  const long long pix = blockIdx.x * 128LL + threadIdx.x;
  if (pix >= P) return;
  const int n = pix / (H * Wd);
  const int rem = pix - (long long)n * H * Wd;
  const int i = rem / Wd, j = rem - i * Wd;
  float acc[2][2][N];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int o = 0; o < N; ++o) acc[a][b][o] = 0.f;
  const float* base = src + ((long long)n * H * Wd) * C;
  bool ok[3][3];
  int off[3][3];
#pragma unroll
  for (int di = 0; di < 3; ++di)
#pragma unroll
    for (int dj = 0; dj < 3; ++dj) {
      const int y = i + di - 1, x = j + dj - 1;
      ok[di][dj] = y >= 0 && y < H && x >= 0 && x < Wd;
      off[di][dj] = ok[di][dj] ? (y * Wd + x) * C : 0;
    }
  for (int c0 = 0; c0 < C; c0 += 4) {
    float xs[3][3][4];
#pragma unroll
    for (int di = 0; di < 3; ++di)
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) {
        const float4 v = ok[di][dj] ? __ldg(reinterpret_cast<const float4*>(base + off[di][dj] + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        xs[di][dj][0] = v.x; xs[di][dj][1] = v.y; xs[di][dj][2] = v.z; xs[di][dj][3] = v.w;
      }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
      for (int o = 0; o < N; ++o) {
        float w[16];
        const float4* wp = reinterpret_cast<const float4*>(w_s + ((c0 + cc) * N + o) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = wp[q];
          w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
        // phase ph uses source rows di-1 in {ph-1, ph}: (ph=0: di=1 -> kh=1, di=0 -> kh=3; ph=1: di=2 -> kh=0, di=1 -> kh=2)
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw)
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const int di = 1 + ph - a, dj = 1 + pw - b;
                const int kh = (1 - ph) + 2 * a, kw = (1 - pw) + 2 * b;
                acc[ph][pw][o] = fmaf(xs[di][dj][cc], w[kh * 4 + kw], acc[ph][pw][o]);
              }
      }
    }
  }
  const int Ho = 2 * H, Wo = 2 * Wd;
#pragma unroll
  for (int o = 0; o < N; ++o)
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      float2* dst = reinterpret_cast<float2*>(out + (((long long)n * N + o) * Ho + 2 * i + ph) * Wo + 2 * j);
      float2 r = make_float2(acc[ph][0][o], acc[ph][1][o]);
      if (act_tanh) { r.x = tanhf(r.x); r.y = tanhf(r.y); }
      if (accumulate) { const float2 prev = *dst; r.x += prev.x; r.y += prev.y; }
      *dst = r;
    }
}

}  // namespace mdgan

using namespace mdgan;

extern "C" int mdgan_thin_up(const float* src, const float* W, float* out, int n_img, int H, int Wd, int C, int N,
                             int act_tanh, int accumulate, void* stream) {
  if (!src || !W || !out) return MDGAN_ERR_BAD_ARG;
  if ((N != 1 && N != 3) || C % 4 != 0 || C <= 0 || C > 256) return MDGAN_ERR_UNSUPPORTED;
  const long long P = (long long)n_img * H * Wd;
  const unsigned blocks = (unsigned)((P + 127) / 128);
  const size_t smem = (size_t)C * N * 16 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 3) {
    static bool configured = false;
    if (!configured) {
      MDGAN_CUDA(cudaFuncSetAttribute(thin_up_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 3 * 16 * 4));
      configured = true;
    }
    thin_up_kernel<3><<<blocks, 128, smem, st>>>(src, W, out, n_img, H, Wd, C, act_tanh, accumulate);
  } else {
    thin_up_kernel<1><<<blocks, 128, smem, st>>>(src, W, out, n_img, H, Wd, C, act_tanh, accumulate);
  }
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_thin_down(const float* img, const float* W, float* out, int n_img, int CI, int Hi, int Wi, int N,
                               int act, float slope, int round_tf32, void* stream) {
  if (!img || !W || !out) return MDGAN_ERR_BAD_ARG;
  if ((CI != 1 && CI != 3) || (N != 64 && N != 128) || (Hi & 1) || (Wi & 1)) return MDGAN_ERR_UNSUPPORTED;
  const long long P = (long long)n_img * (Hi / 2) * (Wi / 2);
  const unsigned blocks = (unsigned)((P + 63) / 64);
  const size_t smem = (size_t)CI * 16 * N * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (CI == 3 && N == 64) thin_down_kernel<3, 16><<<blocks, 256, smem, st>>>(img, W, out, n_img, Hi, Wi, N, act, slope, round_tf32);
  else if (CI == 3 && N == 128) thin_down_kernel<3, 32><<<blocks, 256, smem, st>>>(img, W, out, n_img, Hi, Wi, N, act, slope, round_tf32);
  else if (CI == 1 && N == 64) thin_down_kernel<1, 16><<<blocks, 256, smem, st>>>(img, W, out, n_img, Hi, Wi, N, act, slope, round_tf32);
  else thin_down_kernel<1, 32><<<blocks, 256, smem, st>>>(img, W, out, n_img, Hi, Wi, N, act, slope, round_tf32);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

// Number of per-block partial slices thin_wgrad writes ([slices][C1][CI*16] floats); reduce with
// mdgan_reduce_slices into the PyTorch-layout gradient [C1][CI][4][4].
extern "C" int mdgan_thin_wgrad_slices(int n_img, int Hl, int Wl) {
  const long long P = (long long)n_img * Hl * Wl;
  long long tiles = (P + 31) / 32;
  return (int)(tiles < 296 ? tiles : 296);
}

extern "C" int mdgan_thin_wgrad(const float* feat, const float* img, float* partial, int n_img, int CI, int Hl, int Wl,
                                int C1, void* stream) {
  if (!feat || !img || !partial) return MDGAN_ERR_BAD_ARG;
  if ((CI != 1 && CI != 3) || (C1 != 64 && C1 != 128)) return MDGAN_ERR_UNSUPPORTED;
  const int blocks = mdgan_thin_wgrad_slices(n_img, Hl, Wl);
  cudaStream_t st = (cudaStream_t)stream;
  if (CI == 3 && C1 == 64) thin_wgrad_kernel<3, 64><<<blocks, 256, 0, st>>>(feat, img, partial, n_img, Hl, Wl);
  else if (CI == 3 && C1 == 128) thin_wgrad_kernel<3, 128><<<blocks, 256, 0, st>>>(feat, img, partial, n_img, Hl, Wl);
  else if (CI == 1 && C1 == 64) thin_wgrad_kernel<1, 64><<<blocks, 256, 0, st>>>(feat, img, partial, n_img, Hl, Wl);
  else thin_wgrad_kernel<1, 128><<<blocks, 256, 0, st>>>(feat, img, partial, n_img, Hl, Wl);
  MDGAN_CHECK_LAUNCH();
  return 0;
}
