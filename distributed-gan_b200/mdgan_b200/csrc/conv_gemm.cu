// Implicit-GEMM kernel for the DCGAN "k4 s2 p1" layer family, tcgen05 (kind::tf32) + TMEM + TMA.
//
// One kernel covers (reference call sites in /root/reference/src/datasets/{CIFAR10,CelebA}.py):
//   DOWN : Conv2d(k4,s2,p1) forward            (CIFAR10.py:88,92; CelebA.py:81,85,88)
//          ConvTranspose2d(k4,s2,p1) data-grad  (backward of CIFAR10.py:122,126; CelebA.py:119-131)
//   UP   : ConvTranspose2d(k4,s2,p1) forward    (CIFAR10.py:122,126,130; CelebA.py:119-131), as 4 output-parity
//          phases, each a 2x2 stride-1 conv;    Conv2d(k4,s2,p1) data-grad (the worker's feedback path,
//          /root/reference/src/actors/worker.py:227)
//   DENSE: ConvTranspose2d(k,s1,p0) on a 1x1 input = plain GEMM (CIFAR10.py:118; CelebA.py:113)
//
// GEMM view: rows = positions of the low-resolution grid (n, i, j); A[row, (tap, c)] is gathered from the
// NHWC fp32 source on the fly (never materialised), B = packed weights [N_pad(*4 phases), taps*C] (K-major),
// D[row, n] accumulates in TMEM.  Per CTA: 128 rows x BN columns, K stepped 32 fp32 (=128 B, one swizzle row)
// at a time through a STAGES-deep smem ring.
//   warps 0-7 : A producers (cp.async 16 B gathers with zero-fill for padding, manual 128B swizzle), then epilogue
//   warp  8   : B producer (TMA 2D tiled load, SWIZZLE_128B, mbarrier complete_tx)
//   warp  9   : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem stages
#include "common.cuh"
#include "ptx.cuh"

namespace mdgan {

struct ConvGemmParams {
  const float* src;   // NHWC [n_img, Hs, Ws, C]
  float* dst;         // NHWC [.., N] or NCHW
  const float* bias;  // optional [N]
  int n_img, Hg, Wg;  // GEMM row grid, M = n_img*Hg*Wg
  int Hs, Ws, C;      // source dims (C % 32 == 0)
  int mode;           // 0 DOWN, 1 UP, 2 DENSE
  int N, N_pad;       // real / padded output channels
  int out_nchw;       // 0: NHWC, 1: NCHW
  int act;            // 0: none, 1: tanh
  int M;
  int round_tf32;     // 1: store outputs rounded to TF32 (they feed another tensor-core operand)
};

constexpr int kBM = 128;
constexpr int kBK = 32;  // fp32 elements per K step = 128 bytes
constexpr int kNumProducerWarps = 8;
constexpr int kThreads = (kNumProducerWarps + 2) * 32;

template <int BN, int STAGES>
struct ConvGemmSmem {
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024 B alignment
};

__device__ __forceinline__ float round_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const ConvGemmParams p) {
  using S = ConvGemmSmem<BN, STAGES>;
  constexpr int LAG = STAGES - 2;  // cp.async groups kept in flight per producer thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int phase = blockIdx.z;  // UP: output parity phase (ph*2 + pw)
  const int ph = phase >> 1, pw = phase & 1;

  const int taps = (p.mode == 0) ? 16 : (p.mode == 1 ? 4 : 1);
  const int cchunks = p.C / kBK;
  const int ksteps = taps * cchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kNumProducerWarps * 32 + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 9) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < kNumProducerWarps) {
    // ------------------------------------------------------------------ A producers
    const int chunk = threadIdx.x & 7;   // 16-byte chunk within the 128-byte row
    const int row_in = threadIdx.x >> 3; // 0..31
    const int SI = (p.mode == 0) ? 2 : 1;
    int base_off[4], sh0[4], sw0[4];
    bool row_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + row_in + 32 * i;
      row_ok[i] = m < p.M;
      const int mm = row_ok[i] ? m : 0;
      const int img = mm / (p.Hg * p.Wg);
      const int rem = mm - img * (p.Hg * p.Wg);
      const int gi = rem / p.Wg, gj = rem - gi * p.Wg;
      sh0[i] = gi * SI;
      sw0[i] = gj * SI;
      base_off[i] = ((img * p.Hs + sh0[i]) * p.Ws + sw0[i]) * p.C + chunk * 4;
    }
    const uint32_t a_smem0 = smem_u32(smem);
    int tap = 0, cc = 0;
    for (int it = 0; it < ksteps; ++it) {
      const int s = it % STAGES;
      const uint32_t par = (it / STAGES) & 1;
      mbar_wait(&empty_bar[s], par ^ 1);
      int dh, dw;
      if (p.mode == 0) { dh = (tap >> 2) - 1; dw = (tap & 3) - 1; }
      else if (p.mode == 1) { dh = ph - (tap >> 1); dw = pw - (tap & 1); }
      else { dh = 0; dw = 0; }
      const int tap_off = (dh * p.Ws + dw) * p.C + cc * kBK;
      const uint32_t a_stage = a_smem0 + s * S::kStageBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row_in + 32 * i;
        const int sh = sh0[i] + dh, sw = sw0[i] + dw;
        const bool ok = row_ok[i] && sh >= 0 && sh < p.Hs && sw >= 0 && sw < p.Ws;
        const float* g = p.src + (ok ? (base_off[i] + tap_off) : 0);
        const uint32_t d = a_stage + r * 128 + ((chunk ^ (r & 7)) << 4);
        cp_async_16(d, g, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (it >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[(it - LAG) % STAGES]);
      }
      if (++cc == cchunks) { cc = 0; ++tap; }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (int it = (ksteps > LAG ? ksteps - LAG : 0); it < ksteps; ++it) mbar_arrive(&full_bar[it % STAGES]);

    // ------------------------------------------------------------------ epilogue
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const int q = warp & 3;
    constexpr int kColsPerHalf = (BN >= 32) ? BN / 2 : BN;
    const int half = warp >> 2;
    if (BN >= 32 || half == 0) {
      const int row = q * 32 + lane;
      const int m = m0 + row;
      const bool ok = m < p.M;
      const int mm = ok ? m : 0;
      const int img = mm / (p.Hg * p.Wg);
      const int rem = mm - img * (p.Hg * p.Wg);
      const int gi = rem / p.Wg, gj = rem - gi * p.Wg;
      int Ho, Wo, oh, ow;
      if (p.mode == 1) { Ho = 2 * p.Hg; Wo = 2 * p.Wg; oh = 2 * gi + ph; ow = 2 * gj + pw; }
      else { Ho = p.Hg; Wo = p.Wg; oh = gi; ow = gj; }
#pragma unroll 1
      for (int c16 = 0; c16 < kColsPerHalf; c16 += 16) {
        const int col = half * kColsPerHalf + c16;
        float v[16];
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col, v);
        if (!ok) continue;
        const int nbase = n0 + col;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = v[j];
          if (p.bias != nullptr && nbase + j < p.N) x += __ldg(p.bias + nbase + j);
          if (p.act == 1) x = tanhf(x);
          if (p.round_tf32) x = round_to_tf32(x);
          v[j] = x;
        }
        if (!p.out_nchw) {
          float* o = p.dst + (static_cast<size_t>((img * Ho + oh) * Wo + ow)) * p.N + nbase;
          if (nbase + 16 <= p.N && (p.N & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nbase + j < p.N) o[j] = v[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nbase + j < p.N)
              p.dst[(static_cast<size_t>(img * p.N + nbase + j) * Ho + oh) * Wo + ow] = v[j];
        }
      }
    }
    tc_fence_before_sync();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ B producer (TMA)
    if (lane == 0) {
      const int row0 = (p.mode == 1 ? phase * p.N_pad : 0) + n0;
      for (int it = 0; it < ksteps; ++it) {
        const int s = it % STAGES;
        const uint32_t par = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], par ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], S::kBBytes);
        tma_load_2d(smem_u32(smem + s * S::kStageBytes + S::kABytes), &tmap_w, &full_bar[s], it * kBK, row0);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 0, 0);
      for (int it = 0; it < ksteps; ++it) {
        const int s = it % STAGES;
        const uint32_t par = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], par);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
        const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_tf32(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
  }
}

template <int BN, int STAGES>
static int launch_conv_gemm(const CUtensorMap& tmap, const ConvGemmParams& p, dim3 grid, cudaStream_t st) {
  using S = ConvGemmSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    MDGAN_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    S::kDynamic));
    configured = true;
  }
  conv_gemm_kernel<BN, STAGES><<<grid, kThreads, S::kDynamic, st>>>(tmap, p);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace mdgan

using namespace mdgan;

// See include/mdgan_b200.h for the contract.
extern "C" int mdgan_conv_gemm(const float* src, const float* wpacked, float* dst, const float* bias, int n_img,
                               int Hg, int Wg, int Hs, int Ws, int C, int mode, int N, int N_pad, int out_nchw,
                               int act, int round_tf32, int force_bn, void* stream) {
  if (!src || !wpacked || !dst) return MDGAN_ERR_BAD_ARG;
  if (C <= 0 || C % kBK != 0 || mode < 0 || mode > 2) return MDGAN_ERR_UNSUPPORTED;
  if (N_pad % 16 != 0 || N > N_pad || N <= 0) return MDGAN_ERR_UNSUPPORTED;
  ConvGemmParams p{};
  p.src = src; p.dst = dst; p.bias = bias;
  p.n_img = n_img; p.Hg = Hg; p.Wg = Wg; p.Hs = Hs; p.Ws = Ws; p.C = C;
  p.mode = mode; p.N = N; p.N_pad = N_pad; p.out_nchw = out_nchw; p.act = act;
  p.M = n_img * Hg * Wg;
  p.round_tf32 = round_tf32;
  if (p.M <= 0) return MDGAN_ERR_BAD_ARG;
  const int taps = mode == 0 ? 16 : (mode == 1 ? 4 : 1);
  const int phases = mode == 1 ? 4 : 1;
  const int row_tiles = ceil_div(p.M, kBM);
  // Tile width: widest BN that divides N_pad, narrowed while the grid would leave most of the 148 SMs idle.
  int bn = 16;
  for (int c : {128, 64, 32})
    if (N_pad % c == 0) { bn = c; break; }
  while (bn > 32 && row_tiles * phases * (N_pad / bn) < 148 && N_pad % (bn / 2) == 0) bn /= 2;
  if (force_bn > 0) {
    if (N_pad % force_bn != 0) return MDGAN_ERR_BAD_ARG;
    bn = force_bn;
  }
  CUtensorMap tmap;
  int rc = get_tmap_2d_f32(wpacked, static_cast<uint64_t>(N_pad) * phases, static_cast<uint64_t>(taps) * C, bn, &tmap);
  if (rc != 0) return rc;
  dim3 grid(row_tiles, N_pad / bn, phases);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 128: return launch_conv_gemm<128, 6>(tmap, p, grid, st);
    case 64: return launch_conv_gemm<64, 4>(tmap, p, grid, st);
    case 32: return launch_conv_gemm<32, 4>(tmap, p, grid, st);
    case 16: return launch_conv_gemm<16, 4>(tmap, p, grid, st);
    default: return MDGAN_ERR_UNSUPPORTED;
  }
}
