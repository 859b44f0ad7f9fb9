// CUDA-core kernels for the image-side "thin" layers (1 or 3 image channels; < 5 % of the step's FLOPs):
// with K = 16 or 48 (or N = 4 or 12 outputs) a 128-wide tensor-core tile would be > 60 % padding, so these run as
// register-blocked fp32 FMA kernels whose floor is the HBM/L2 traffic of the feature-side tensor.  The image side is
// NCHW fp32 (the layout of every tensor that crosses the reference's actor boundary: real batches, generated batches,
// feedback), the feature side NHWC.
//
//   thin_down  : out[n,i,j,co] = sum_{c,kh,kw} img[n,c,2i-1+kh,2j-1+kw] * W[co][c][kh][kw]  (+ LeakyReLU)
//                = first discriminator Conv2d(3->64,k4,s2,p1) forward (CIFAR10.py:85, CelebA.py:78) and the
//                data-gradient of the last generator ConvTranspose2d(->3) (W is then [ci][co][kh][kw]).
//   thin_wgrad : dW[c1][c2][kh][kw] = sum_p feat[p,c1] * img[n,c2,2i-1+kh,2j-1+kw]
//                = weight gradient of both of those layers.
//   thin_up    : out[n,o,2i+ph,2j+pw] = sum_{c,a,b} src[n,i+ph-a,j+pw-b,c] * W[c][o][(1-ph)+2a][(1-pw)+2b]
//                = last generator ConvTranspose2d(->3)+tanh forward and the data gradient of the first
//                discriminator conv (the error feedback, accumulated into its slot).
#include "common.cuh"

namespace mdgan {

__device__ __forceinline__ float thin_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------------ thin_down
// Block = 256 threads = PG pixel groups x CG channel groups (CG = N/16); a thread owns PPT pixels x 16 channels
// (channels {4*(j*CG + cg) + t}: per j the CG lanes of a pixel read/write one contiguous 16*CG-byte run, so the
// shared-memory weight reads are conflict-free broadcasts and the NHWC stores are full sectors).
template <int CI, int N, int PPT>
__global__ void __launch_bounds__(256) thin_down_kernel(const float* __restrict__ img, const float* __restrict__ W,
                                                        float* __restrict__ out, int n_img, int Hi, int Wi, int act,
                                                        float slope, int round_tf32) {
  constexpr int CG = N / 16, PG = 256 / CG, K = CI * 16, WS = N + 4;  // WS: padded row stride (transposed staging)
  extern __shared__ float w_s[];  // [K][WS]
  pdl_enter();
  for (int i = threadIdx.x; i < K * N; i += 256) {
    const int co = i / K, k = i - co * K;
    w_s[k * WS + co] = W[i];
  }
  __syncthreads();
  const int Ho = Hi >> 1, Wo = Wi >> 1;
  const long long P = (long long)n_img * Ho * Wo;
  const int cg = threadIdx.x % CG, pg = threadIdx.x / CG;
  const long long tile0 = (long long)blockIdx.x * (PG * PPT);
  const float* base[PPT];
  int ih0[PPT], iw0[PPT];
  bool valid[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const long long pix = tile0 + q * PG + pg;
    valid[q] = pix < P;
    const long long pp = valid[q] ? pix : 0;
    const int n = pp / (Ho * Wo);
    const int rem = pp - (long long)n * Ho * Wo;
    const int oi = rem / Wo, oj = rem - oi * Wo;
    base[q] = img + (long long)n * CI * Hi * Wi;
    ih0[q] = 2 * oi - 1;
    iw0[q] = 2 * oj - 1;
  }
  float acc[PPT][16];
#pragma unroll
  for (int q = 0; q < PPT; ++q)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[q][j] = 0.f;
#pragma unroll
  for (int c = 0; c < CI; ++c) {
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      float v[PPT][4];
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        const int ih = ih0[q] + kh;
        const bool rok = valid[q] && ih >= 0 && ih < Hi;
        const float* row = base[q] + ((long long)c * Hi + ih) * Wi;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const int iw = iw0[q] + kw;
          v[q][kw] = (rok && iw >= 0 && iw < Wi) ? __ldg(row + iw) : 0.f;
        }
      }
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const float4* wr = reinterpret_cast<const float4*>(w_s + (c * 16 + kh * 4 + kw) * WS) + cg;
        float w[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = wr[j * CG];
          w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[q][j] = fmaf(v[q][kw], w[j], acc[q][j]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    if (!valid[q]) continue;
    float4* o = reinterpret_cast<float4*>(out + (tile0 + q * PG + pg) * N) + cg;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float x = acc[q][4 * j + t];
        if (act == 2) x = x > 0.f ? x : x * slope;
        if (round_tf32) x = thin_to_tf32(x);
        r[t] = x;
      }
      o[j * CG] = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------ thin_wgrad
// A [C1] x [K = CI*16] outer-product accumulation over all pixels.  Block = TILES register tiles ([4 channels] x
// [KPT taps]) x PL pixel lanes; a warp holds 32 tiles of ONE pixel lane, so per staged pixel its shared-memory reads
// are one 16-channel-group run (2 wavefronts) + KPT/4 broadcast float4 of the patch for 4*KPT FMAs per thread.
// Blocks walk pixel tiles of 32 (grid-stride) with the next tile's global loads in flight in registers while the
// current one is consumed from shared memory, then combine their pixel lanes through shared memory (conflict-free
// [lane][value][tile] layout, fixed summation order) and write one partial slice [C1][K] per block (reduced by
// mdgan_reduce_slices).
template <int CI, int C1>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ feat, const float* __restrict__ img,
                                                         float* __restrict__ partial, int n_img, int Hl, int Wl) {
  constexpr int K = CI * 16;
  constexpr int KPT = CI == 3 ? 12 : 16;   // taps per thread
  constexpr int TS = K / KPT;              // tap slices
  constexpr int CGS = C1 / 4;              // channel groups of 4
  constexpr int TILES = CGS * TS;          // register tiles per pixel lane
  constexpr int PL = 256 / TILES;          // pixel lanes
  constexpr int TP = 32;                   // pixels staged per step
  constexpr int FV = TP * C1 / 4 / 256;    // float4 of the feature tile per thread
  constexpr int PV = K / 8;                // patch values per thread (thread = pixel tid/8, taps tid%8 + 8u)
  static_assert(TILES * PL == 256 && TP % PL == 0, "thread layout");
  __shared__ __align__(16) float f_s[TP * C1];
  __shared__ __align__(16) float p_s[TP * K];
  __shared__ __align__(16) float red[256 * 16];
  pdl_enter();
  const int Hi = 2 * Hl, Wi = 2 * Wl, HW = Hl * Wl;
  const long long P = (long long)n_img * HW;
  const int tile = threadIdx.x % TILES, pl = threadIdx.x / TILES;
  const int g = tile % CGS, ts = tile / CGS;
  const int pr = threadIdx.x >> 3, pk = threadIdx.x & 7;
  float acc[4][KPT];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int j = 0; j < KPT; ++j) acc[a][j] = 0.f;
  float4 fv[FV];
  float pv[PV];
  auto prefetch = [&](long long p0) {
#pragma unroll
    for (int u = 0; u < FV; ++u) {
      const int i = threadIdx.x + 256 * u;
      const int r = i / (C1 / 4);
      fv[u] = (p0 + r < P) ? __ldg(reinterpret_cast<const float4*>(feat + p0 * C1) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long pix = p0 + pr;
    const bool ok = pix < P;
    const unsigned pp = ok ? (unsigned)pix : 0u;
    const unsigned n = pp / (unsigned)HW;
    const unsigned rem = pp - n * (unsigned)HW;
    const int oi = rem / (unsigned)Wl, oj = rem - (rem / (unsigned)Wl) * (unsigned)Wl;
    const float* base = img + (long long)n * CI * Hi * Wi;
#pragma unroll
    for (int u = 0; u < PV; ++u) {
      const int k = pk + 8 * u;
      const int c = k >> 4, kh = (k >> 2) & 3, kw = k & 3;
      const int ih = 2 * oi - 1 + kh, iw = 2 * oj - 1 + kw;
      pv[u] = (ok && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) ? __ldg(base + ((long long)c * Hi + ih) * Wi + iw) : 0.f;
    }
  };
  const long long stride = (long long)gridDim.x * TP;
  long long p0 = blockIdx.x * (long long)TP;
  if (p0 < P) prefetch(p0);
  for (; p0 < P; p0 += stride) {
#pragma unroll
    for (int u = 0; u < FV; ++u) reinterpret_cast<float4*>(f_s)[threadIdx.x + 256 * u] = fv[u];
#pragma unroll
    for (int u = 0; u < PV; ++u) p_s[pr * K + pk + 8 * u] = pv[u];
    __syncthreads();
    if (p0 + stride < P) prefetch(p0 + stride);
#pragma unroll 2
    for (int r = pl; r < TP; r += PL) {
      const float4 f4 = *reinterpret_cast<const float4*>(f_s + r * C1 + g * 4);
      const float f[4] = {f4.x, f4.y, f4.z, f4.w};
      float p[KPT];
#pragma unroll
      for (int j = 0; j < KPT / 4; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(p_s + r * K + ts * KPT + 4 * j);
        p[4 * j] = t.x; p[4 * j + 1] = t.y; p[4 * j + 2] = t.z; p[4 * j + 3] = t.w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < KPT; ++j) acc[a][j] = fmaf(f[a], p[j], acc[a][j]);
    }
    __syncthreads();
  }
  // combine the PL pixel lanes, 16 of the 4*KPT register values at a time: red[lane][value][tile]
  float* o = partial + (long long)blockIdx.x * C1 * K;
#pragma unroll
  for (int ep = 0; ep < 4 * KPT; ep += 16) {
#pragma unroll
    for (int e = 0; e < 16; ++e) red[(pl * 16 + e) * TILES + tile] = acc[(ep + e) / KPT][(ep + e) % KPT];
    __syncthreads();
    for (int q = threadIdx.x; q < 16 * TILES; q += 256) {
      const int el = q / TILES, to = q - el * TILES;
      float sum = 0.f;
#pragma unroll
      for (int l = 0; l < PL; ++l) sum += red[(l * 16 + el) * TILES + to];
      const int e = ep + el;
      o[((to % CGS) * 4 + e / KPT) * K + (to / CGS) * KPT + e % KPT] = sum;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------ thin_up
// The 4-phase form of ConvTranspose2d(C -> N, k4, s2, p1) on an NHWC source, N in {1, 3}.  Block = 32 pairs of
// horizontally adjacent low-resolution positions x 8 channel lanes; a lane owns channels {4*(l + 8t) + cc}, reads
// the 3x4 neighbourhood of its pair as float4 (the 8 lanes of a pair read one contiguous 128 B run), keeps the
// 2 x 4 x N partial outputs in registers, and the 8 lanes are combined by a butterfly.  Weights sit in shared memory
// as [cc][C/4][N*16 (+4 pad)] so that the lanes' float4 reads fall into distinct banks.
template <int N>
__global__ void __launch_bounds__(256, 2) thin_up_kernel(const float* __restrict__ src, const float* __restrict__ W,
                                                      float* __restrict__ out, int n_img, int H, int Wd, int C,
                                                      int act_tanh, int accumulate) {
  constexpr int S = N * 16 + 4;
  extern __shared__ float w_s[];  // [4][C/4][S]
  pdl_enter();
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < C * N * 16; i += 256) {
    const int c = i / (N * 16), r = i - c * (N * 16);
    w_s[((c & 3) * C4 + (c >> 2)) * S + r] = W[i];
  }
  __syncthreads();
  const int lane8 = threadIdx.x & 7, pr = threadIdx.x >> 3;
  const int Wp = (Wd + 1) >> 1;  // pairs per row
  const long long PP = (long long)n_img * H * Wp;
  const int Ho = 2 * H, Wo = 2 * Wd;
  for (long long t0 = blockIdx.x * 32LL; t0 < PP; t0 += gridDim.x * 32LL) {
    const long long pair = t0 + pr;
    const bool pv = pair < PP;
    const long long pq = pv ? pair : 0;
    const int n = pq / (H * Wp);
    const int rem = pq - (long long)n * H * Wp;
    const int i = rem / Wp, j0 = 2 * (rem - i * Wp);
    const float* base = src + (long long)n * H * Wd * C;
    float acc[2][2][2][N];  // [pos][ph][pw][o]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
          for (int o = 0; o < N; ++o) acc[a][b][d][o] = 0.f;
    bool ok[3][4];
    int off[3][4];
#pragma unroll
    for (int di = 0; di < 3; ++di)
#pragma unroll
      for (int dj = 0; dj < 4; ++dj) {
        const int y = i + di - 1, x = j0 + dj - 1;
        ok[di][dj] = pv && y >= 0 && y < H && x >= 0 && x < Wd;
        off[di][dj] = ok[di][dj] ? (y * Wd + x) * C : 0;
      }
    for (int q = lane8; q < C4; q += 8) {
      float xs[3][4][4];
#pragma unroll
      for (int di = 0; di < 3; ++di)
#pragma unroll
        for (int dj = 0; dj < 4; ++dj) {
          const float4 v = ok[di][dj] ? __ldg(reinterpret_cast<const float4*>(base + off[di][dj]) + q)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          xs[di][dj][0] = v.x; xs[di][dj][1] = v.y; xs[di][dj][2] = v.z; xs[di][dj][3] = v.w;
        }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float4* wp = reinterpret_cast<const float4*>(w_s + (cc * C4 + q) * S);
#pragma unroll
        for (int o = 0; o < N; ++o) {
          float w[16];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4 v = wp[o * 4 + t];
            w[4 * t] = v.x; w[4 * t + 1] = v.y; w[4 * t + 2] = v.z; w[4 * t + 3] = v.w;
          }
          // phase ph of position `pos` uses source rows di - 1 in {ph - 1, ph} and columns (pos + dj) - 1 likewise:
          // di = 1 + ph - a <-> kh = (1 - ph) + 2a,  dj = pos + 1 + pw - b <-> kw = (1 - pw) + 2b
#pragma unroll
          for (int pos = 0; pos < 2; ++pos)
#pragma unroll
            for (int ph = 0; ph < 2; ++ph)
#pragma unroll
              for (int pw = 0; pw < 2; ++pw)
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                  for (int b = 0; b < 2; ++b) {
                    const int di = 1 + ph - a, dj = pos + 1 + pw - b;
                    const int kh = (1 - ph) + 2 * a, kw = (1 - pw) + 2 * b;
                    acc[pos][ph][pw][o] = fmaf(xs[di][dj][cc], w[kh * 4 + kw], acc[pos][ph][pw][o]);
                  }
        }
      }
    }
    // butterfly over the 8 channel lanes (lanes 8k..8k+7 of a warp share a pair)
#pragma unroll
    for (int pos = 0; pos < 2; ++pos)
#pragma unroll
      for (int ph = 0; ph < 2; ++ph)
#pragma unroll
        for (int pw = 0; pw < 2; ++pw)
#pragma unroll
          for (int o = 0; o < N; ++o) {
            float v = acc[pos][ph][pw][o];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            acc[pos][ph][pw][o] = v;
          }
    // lane (o*2 + ph) writes the 4 outputs of its row: columns 2*j0 .. 2*j0+3
    if (pv) {
#pragma unroll
      for (int o = 0; o < N; ++o)
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          if (lane8 != o * 2 + ph) continue;
          float r[4] = {acc[0][ph][0][o], acc[0][ph][1][o], acc[1][ph][0][o], acc[1][ph][1][o]};
          float* dst = out + (((long long)n * N + o) * Ho + 2 * i + ph) * Wo + 2 * j0;
          const int cnt = (j0 + 1 < Wd) ? 4 : 2;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t >= cnt) break;
            float x = r[t];
            if (act_tanh) x = tanhf(x);
            if (accumulate) x += dst[t];
            dst[t] = x;
          }
        }
    }
  }
}

}  // namespace mdgan

using namespace mdgan;

static inline unsigned grid_cap(long long tiles, unsigned cap) { return (unsigned)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap); }

extern "C" int mdgan_thin_up(const float* src, const float* W, float* out, int n_img, int H, int Wd, int C, int N,
                             int act_tanh, int accumulate, void* stream) {
  if (!src || !W || !out) return MDGAN_ERR_BAD_ARG;
  if ((N != 1 && N != 3) || C % 4 != 0 || C <= 0 || C > 256) return MDGAN_ERR_UNSUPPORTED;
  const long long pairs = (long long)n_img * H * ((Wd + 1) / 2);
  const unsigned blocks = grid_cap((pairs + 31) / 32, 148 * 8);
  const size_t smem = (size_t)C * (N * 16 + 4) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 3) {
    MDGAN_CUDA(configure_smem_once(thin_up_kernel<3>, 256 * 52 * 4));
    MDGAN_LAUNCH(thin_up_kernel<3>, dim3(blocks), dim3(256), smem, st, src, W, out, n_img, H, Wd, C, act_tanh, accumulate);
  } else {
    MDGAN_LAUNCH(thin_up_kernel<1>, dim3(blocks), dim3(256), smem, st, src, W, out, n_img, H, Wd, C, act_tanh, accumulate);
  }
  return 0;
}

template <int CI, int N>
static int launch_thin_down(const float* img, const float* W, float* out, int n_img, int Hi, int Wi, int act, float slope,
                            int round_tf32, cudaStream_t st) {
  constexpr int PG = 256 / (N / 16);
  const long long P = (long long)n_img * (Hi / 2) * (Wi / 2);
  const size_t smem = (size_t)CI * 16 * (N + 4) * sizeof(float);
  // two pixels per thread (115 registers, two blocks per SM) when that still gives every SM two blocks
  if ((P + PG * 2 - 1) / (PG * 2) >= 296)
    MDGAN_LAUNCH((thin_down_kernel<CI, N, 2>), dim3((unsigned)((P + PG * 2 - 1) / (PG * 2))), dim3(256), smem, st, img, W,
                 out, n_img, Hi, Wi, act, slope, round_tf32);
  else
    MDGAN_LAUNCH((thin_down_kernel<CI, N, 1>), dim3((unsigned)((P + PG - 1) / PG)), dim3(256), smem, st, img, W, out, n_img,
                 Hi, Wi, act, slope, round_tf32);
  return 0;
}

extern "C" int mdgan_thin_down(const float* img, const float* W, float* out, int n_img, int CI, int Hi, int Wi, int N,
                               int act, float slope, int round_tf32, void* stream) {
  if (!img || !W || !out) return MDGAN_ERR_BAD_ARG;
  if ((CI != 1 && CI != 3) || (N != 64 && N != 128) || (Hi & 1) || (Wi & 1)) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (CI == 3 && N == 64) return launch_thin_down<3, 64>(img, W, out, n_img, Hi, Wi, act, slope, round_tf32, st);
  if (CI == 3 && N == 128) return launch_thin_down<3, 128>(img, W, out, n_img, Hi, Wi, act, slope, round_tf32, st);
  if (CI == 1 && N == 64) return launch_thin_down<1, 64>(img, W, out, n_img, Hi, Wi, act, slope, round_tf32, st);
  return launch_thin_down<1, 128>(img, W, out, n_img, Hi, Wi, act, slope, round_tf32, st);
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
  if (CI == 3 && C1 == 64) MDGAN_LAUNCH((thin_wgrad_kernel<3, 64>), dim3(blocks), dim3(256), 0, st, feat, img, partial, n_img, Hl, Wl);
  else if (CI == 3 && C1 == 128) MDGAN_LAUNCH((thin_wgrad_kernel<3, 128>), dim3(blocks), dim3(256), 0, st, feat, img, partial, n_img, Hl, Wl);
  else if (CI == 1 && C1 == 64) MDGAN_LAUNCH((thin_wgrad_kernel<1, 64>), dim3(blocks), dim3(256), 0, st, feat, img, partial, n_img, Hl, Wl);
  else MDGAN_LAUNCH((thin_wgrad_kernel<1, 128>), dim3(blocks), dim3(256), 0, st, feat, img, partial, n_img, Hl, Wl);
  return 0;
}
