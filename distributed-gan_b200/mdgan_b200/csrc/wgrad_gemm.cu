// Weight-gradient implicit GEMM for the "k4 s2 p1" layer family, tcgen05 (kind::tf32, both operands MN-major).
//
//   dW_tap[c1, c2] = sum over low-res positions p=(n,i,j) of  Lo[p, c1] * Hi[(n, 2i-1+kh, 2j-1+kw), c2]
//
// which is the weight gradient of both Conv2d(k4,s2,p1) (Lo = grad of the conv output, Hi = conv input;
// result [c1=co][c2=ci]) and ConvTranspose2d(k4,s2,p1) (Lo = convT input, Hi = grad of the convT output;
// result [c1=ci][c2=co]) -- the autograd work behind /root/reference/src/actors/worker.py:204 and
// /root/reference/src/actors/server.py:286-292.  mode 2 (identity gather, 1 tap) is the weight gradient of
// the generator's first layer (ConvTranspose2d on a 1x1 input = plain GEMM).
//
// The reduction dimension is the pixel index, which is the slow dimension of both NHWC operands, so both smem
// operands are MN-major: a stage holds 32 pixels; each 32-channel group is a [32 pixel rows x 128 B] block in the
// SWIZZLE_128B_BASE32B layout (4-row atoms, 32-byte swizzle granularity -- the only MN-major layout tcgen05
// accepts for 32-bit operands).  One tcgen05.mma (K = 8 pixels) consumes two 4-row atoms of every group.
// Split-K over pixels fills the machine; partial sums go to [split][tap][C1][C2] and are reduced (in a fixed
// order, deterministically) by the unpack kernel in elementwise.cu.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace mdgan {

struct WgradParams {
  const float* lo;  // NHWC [n_img, Hl, Wl, C1]
  const float* hi;  // NHWC [n_img, Hh, Wh, C2]  (Hh = 2*Hl for mode 0; = Hl for mode 2)
  float* partial;   // [splits][taps][C1][C2]
  int n_img, Hl, Wl, Hh, Wh, C1, C2;
  int mode;         // 0: k4 s2 p1 gather (16 taps), 2: identity (1 tap)
  int P;            // n_img*Hl*Wl
  int splits, pix_per_split;
  int lbo_a, sbo_a, lbo_b, sbo_b;  // descriptor byte offsets (see file header)
  unsigned long long magic_hw, magic_w;  // ceil(2^40 / (Hl*Wl)), ceil(2^40 / Wl): q = (n * magic) >> 40 (exact, n < 2^28)
};

constexpr int kWM = 128;
constexpr int kWK = 32;  // pixels per stage
constexpr int kWProducerWarps = 8;
constexpr int kWThreads = (kWProducerWarps + 1) * 32;
constexpr int kWPrefetch = 3;  // K steps of operand loads kept in flight in registers per producer thread
// 2 groups: 62-74 us per CelebA b=64 layer; 4 groups: 57-66 us.  A group may run at most ONE barrier phase ahead of the
// MMAs on any stage, because an mbarrier parity wait cannot tell "the phase I need" from "two phases earlier": group g's
// consecutive steps are kHiGroups apart, so it can be kHiGroups / kStages phases ahead on a stage -- the number of groups
// must not exceed the number of stages (static_assert in the kernel).  Round 2 first ran 4 groups over 3 stages: one
// benchmark process in five dead-locked (a group passed the `empty` wait of a stage whose previous contents had not
// been consumed yet); 2 groups over 3 stages and 4 over 4 are safe by this argument and never stalled.
#ifndef MDGAN_WGRAD_HI_GROUPS
#define MDGAN_WGRAD_HI_GROUPS 4
#endif
constexpr int kHiGroups = MDGAN_WGRAD_HI_GROUPS;  // producer-warp groups of the TMEM-operand kernel (see its Hi producers)

// Byte offset of 16-byte chunk `c16` (0..7) of pixel row `r` inside its 128-byte row under SWIZZLE_128B_BASE32B:
// the 32-byte chunk index is XORed with (row & 3).
__device__ __forceinline__ uint32_t swz32(int c16, int r) {
  return static_cast<uint32_t>(((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4)));
}

template <int BN, int STAGES, bool X3>
struct WgradSmem {
  static constexpr int kABytes = (kWM / 32) * kWK * 128;
  static constexpr int kBBytes = (BN / 32) * kWK * 128;
  static constexpr int kHalfBytes = kABytes + kBBytes;  // [A_hi | B_hi] then (tf32x3) [A_lo | B_lo]
  static constexpr int kStageBytes = (X3 ? 2 : 1) * kHalfBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
  // tf32x3: several TMEM accumulators keep the truncating in-tensor-core accumulation chains short (see
  // conv_gemm.cu): K-slice k of a stage -> main accumulator k % kMain, plus one for the correction products.
  static constexpr int kMain = X3 ? (BN > 64 ? 2 : 4) : 1;
  static constexpr int kAccs = X3 ? kMain + 1 : 1;
  static constexpr uint32_t kTmemCols = kAccs * BN <= 64 ? 64 : kAccs * BN <= 128 ? 128 : kAccs * BN <= 256 ? 256 : 512;
  static_assert(kAccs * BN <= 512, "TMEM holds 512 columns");
};

__device__ __forceinline__ void wsts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// store one 16-byte chunk rounded to TF32 (hi) and, for tf32x3, its exact remainder (lo) lo_delta bytes further
template <bool X3>
__device__ __forceinline__ void store_split(uint32_t addr, uint32_t lo_delta, const float4& v) {
  const float hx = tf32_round_fast(v.x), hy = tf32_round_fast(v.y), hz = tf32_round_fast(v.z), hw = tf32_round_fast(v.w);
  wsts128(addr, hx, hy, hz, hw);
  if (X3) wsts128(addr + lo_delta, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
}

template <int BN, int STAGES, bool X3>
__global__ void __launch_bounds__(kWThreads, 1) wgrad_gemm_kernel(const WgradParams p) {
  using S = WgradSmem<BN, STAGES, X3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int c1_0 = blockIdx.x * kWM;
  const int c2_0 = blockIdx.y * BN;
  const int tap = blockIdx.z / p.splits;
  const int split = blockIdx.z - tap * p.splits;
  const int pix0 = split * p.pix_per_split;
  const int pix1 = min(p.P, pix0 + p.pix_per_split);
  const int ksteps = (pix1 > pix0) ? (pix1 - pix0 + kWK - 1) / kWK : 0;
  const int dh = (p.mode == 0) ? (tap >> 2) - 1 : 0;
  const int dw = (p.mode == 0) ? (tap & 3) - 1 : 0;
  const int SI = (p.mode == 0) ? 2 : 1;

  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kWProducerWarps);  // one arrive per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == kWProducerWarps) tmem_alloc<S::kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp < kWProducerWarps) {
    // ---------------------------------------------------------------- producers (both operands)
    // global -> registers (kWPrefetch K steps ahead) -> TF32 hi/lo split in registers -> one swizzled store each.
    const uint32_t smem0 = smem_u32(smem);
    // A (Lo): 32 chunks of 16 B per pixel row (4 groups x 8 chunks); thread -> (cidx = tid%32, rows tid/32 + 8i)
    const int a_cidx = threadIdx.x & 31;
    const int a_row0 = threadIdx.x >> 5;
    // B (Hi): BN/4 chunks per pixel row
    constexpr int kBChunksPerRow = BN / 4;
    constexpr int kBRowsPerPass = 256 / kBChunksPerRow;
    constexpr int kBPasses = kWK / kBRowsPerPass;
    const int b_cidx = threadIdx.x % kBChunksPerRow;
    const int b_row0 = threadIdx.x / kBChunksPerRow;
    const int hw_l = p.Hl * p.Wl;
    uint32_t a_soff[4], b_soff[kBPasses];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = a_row0 + 8 * i;
      a_soff[i] = (a_cidx >> 3) * (kWK * 128) + r * 128 + swz32(a_cidx & 7, r);
    }
#pragma unroll
    for (int i = 0; i < kBPasses; ++i) {
      const int r = b_row0 + kBRowsPerPass * i;
      b_soff[i] = S::kABytes + (b_cidx >> 3) * (kWK * 128) + r * 128 + swz32(b_cidx & 7, r);
    }
    int pbase_l = pix0;  // first pixel of the next K step to load
    auto issue_loads = [&](float4(&abuf)[4], float4(&bbuf)[kBPasses]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pix = pbase_l + a_row0 + 8 * i;
        abuf[i] = pix < pix1 ? __ldg(reinterpret_cast<const float4*>(p.lo + static_cast<size_t>(pix) * p.C1 + c1_0 + a_cidx * 4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < kBPasses; ++i) {
        const int pix = pbase_l + b_row0 + kBRowsPerPass * i;
        bool ok = pix < pix1;
        size_t off = 0;
        if (ok) {
          const int img = static_cast<int>((static_cast<unsigned long long>(pix) * p.magic_hw) >> 40);
          const int rem = pix - img * hw_l;
          const int gi = static_cast<int>((static_cast<unsigned long long>(rem) * p.magic_w) >> 40);
          const int gj = rem - gi * p.Wl;
          const int sh = gi * SI + dh, sw = gj * SI + dw;
          ok = sh >= 0 && sh < p.Hh && sw >= 0 && sw < p.Wh;
          off = (static_cast<size_t>((img * p.Hh + sh) * p.Wh + sw)) * p.C2 + c2_0 + b_cidx * 4;
        }
        bbuf[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.hi + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      pbase_l += kWK;
    };
    float4 abuf[kWPrefetch][4], bbuf[kWPrefetch][kBPasses];
#pragma unroll
    for (int u = 0; u < kWPrefetch; ++u)
      if (u < ksteps) issue_loads(abuf[u], bbuf[u]);
    int s = 0;
    uint32_t par = 0;
    for (int it0 = 0; it0 < ksteps; it0 += kWPrefetch) {
#pragma unroll
      for (int u = 0; u < kWPrefetch; ++u) {
        const int it = it0 + u;
        if (it < ksteps) {
          mbar_wait(&empty_bar[s], par ^ 1);
          const uint32_t stage = smem0 + s * S::kStageBytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) store_split<X3>(stage + a_soff[i], S::kHalfBytes, abuf[u][i]);
#pragma unroll
          for (int i = 0; i < kBPasses; ++i) store_split<X3>(stage + b_soff[i], S::kHalfBytes, bbuf[u][i]);
          fence_proxy_async_smem();  // every thread publishes its own stores to the async proxy ...
          __syncwarp();              // ... the warp agrees they are all done, and one lane arrives for the 32
          if (lane == 0) mbar_arrive(&full_bar[s]);
          if (it + kWPrefetch < ksteps) issue_loads(abuf[u], bbuf[u]);
          if (++s == STAGES) { s = 0; par ^= 1; }
        }
      }
    }

    // ---------------------------------------------------------------- epilogue: TMEM -> partial[split][tap]
    if (ksteps > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
    }
    const int q = warp & 3, half = warp >> 2;
    constexpr int kColsPerHalf = BN / 2;
    const int c1 = c1_0 + q * 32 + lane;
    const int taps = gridDim.z / p.splits;
    float* orow = p.partial + (static_cast<size_t>(split * taps + tap) * p.C1 + c1) * p.C2 + c2_0;
#pragma unroll 1
    for (int c16 = 0; c16 < kColsPerHalf; c16 += 16) {
      const int col = half * kColsPerHalf + c16;
      float v[16];
      if (ksteps > 0) {
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col, v);
        if (X3) {
          float t[16];
#pragma unroll
          for (int a = 1; a < S::kAccs; ++a) {
            tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + col, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += t[j];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (c1 < p.C1) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(orow + col + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before_sync();
  } else {
    // ---------------------------------------------------------------- MMA issuer (lean loop, see conv_gemm.cu)
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      constexpr uint32_t idesc = make_idesc_tf32(kWM, BN, 1, 1);
      const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem), p.lbo_a, p.sbo_a, 1);
      constexpr uint64_t kStageStep = S::kStageBytes >> 4, kBStep = S::kABytes >> 4, kLoStep = S::kHalfBytes >> 4;
      int s = 0;
      uint32_t par = 0;
      uint64_t da0 = desc0;
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&full_bar[s], par);
        tc_fence_after_sync();
        const uint32_t acc = it != 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < kWK / 8; ++k) {
          const uint64_t da = da0 + 64 * k, db = da + kBStep;  // 8 pixels = two 512-byte k-atoms
          if (X3) {
            const uint32_t acc_corr = tmem_base + S::kMain * BN;
            umma_tf32(acc_corr, da + kLoStep, db, idesc, k == 0 ? acc : 1u);
            umma_tf32(acc_corr, da, db + kLoStep, idesc, 1u);
            umma_tf32(tmem_base + (k % S::kMain) * BN, da, db, idesc, k < S::kMain ? acc : 1u);
          } else {
            umma_tf32(tmem_base, da, db, idesc, k == 0 ? acc : 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        da0 += kStageStep;
        if (++s == STAGES) { s = 0; par ^= 1; da0 = desc0; }
      }
      if (ksteps > 0) umma_commit(tmem_full_bar);
    }
  }
  __syncthreads();
  if (warp == kWProducerWarps) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}

// ================================================================================================================
// tf32x3 variant with the Lo operand (the M side: 128 channels x 32 pixels per K step) in tensor memory.  The raw fp32
// tile [32 pixels][128 channels] arrives by TMA (four 2-D boxes of 32 pixels x 32 channels, 128B swizzle, pixels past
// the end read as zero); transposer warp q owns channel group q = TMEM lane quarter q: a thread reads ITS channel of
// the 32 pixel rows (a warp reads 32 consecutive words of a row: conflict-free under the swizzle), splits hi / lo in
// registers and writes its TMEM lane (tcgen05.st).  The MMAs read A from TMEM; only the Hi operand (N side) is still
// produced through shared memory by the eight producer warps.  Shared-memory/L1 traffic per K step: 32 KB for Lo
// (TMA write + one read) instead of 112 KB (L1 fill + read, hi/lo stores, three MMA reads).
//   warps 0..4*TG-1 : TG groups of four transposer warps (group g takes K steps g, g+TG, ...; one group's chain of
//                     wait raw -> 32 LDS -> split -> wait TMEM stage -> 2 tcgen05.st -> wait::st -> arrive is longer
//                     than the MMA time of a K step, so consecutive K steps are transposed concurrently)
//   next 8 warps    : Hi producers, then the epilogue
//   next warp       : Lo TMA issuer (lane 0);  last warp: TMEM allocator + single-thread MMA issuer (lane 0)
template <int TG>
constexpr int wta_threads() { return (4 * TG + kWProducerWarps + 2) * 32; }

template <int BN>
struct WgradTaSmem {
  static constexpr int kBBytes = (BN / 32) * kWK * 128;   // one of hi / lo
  static constexpr int kStageBytes = 2 * kBBytes;         // [B_hi | B_lo]
  static constexpr int kStages = 4;                       // >= kHiGroups (see MDGAN_WGRAD_HI_GROUPS)
  static constexpr int kRawBytes = (kWM / 32) * kWK * 128;  // 16 KB raw Lo tile
  static constexpr int kRawStages = 4;
  static constexpr int kAStages = BN > 64 ? 2 : 3;        // TMEM stages of [A_hi (32 columns) | A_lo (32 columns)]
  static constexpr int kRawOff = kStages * kStageBytes;
  static constexpr int kBarOffset = kRawOff + kRawStages * kRawBytes;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 2 * kRawStages + 2 * kAStages + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static constexpr int kMain = BN > 64 ? 2 : 4;
  static constexpr int kAccs = kMain + 1;
  static constexpr int kACol0 = kAccs * BN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kACol0 + kAStages * 64 <= 512, "TMEM holds 512 columns");
};

template <int BN, int TG>
__global__ void __launch_bounds__(wta_threads<TG>(), 1)
wgrad_gemm_ta_kernel(const __grid_constant__ CUtensorMap tmap_lo, const WgradParams p) {
  using S = WgradTaSmem<BN>;
  constexpr int kProd0 = 4 * TG;                       // first Hi-producer warp
  constexpr int kTmaWarp = kProd0 + kWProducerWarps;   // Lo TMA issuer
  constexpr int kMmaWarp = kTmaWarp + 1;               // TMEM allocator + MMA issuer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* b_empty = b_full + S::kStages;
  uint64_t* raw_full = b_empty + S::kStages;
  uint64_t* raw_empty = raw_full + S::kRawStages;
  uint64_t* a_full = raw_empty + S::kRawStages;
  uint64_t* a_empty = a_full + S::kAStages;
  uint64_t* tmem_full_bar = a_empty + S::kAStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int c1_0 = blockIdx.x * kWM;
  const int c2_0 = blockIdx.y * BN;
  const int tap = blockIdx.z / p.splits;
  const int split = blockIdx.z - tap * p.splits;
  const int pix0 = split * p.pix_per_split;
  const int pix1 = min(p.P, pix0 + p.pix_per_split);
  const int ksteps = (pix1 > pix0) ? (pix1 - pix0 + kWK - 1) / kWK : 0;
  const int dh = (p.mode == 0) ? (tap >> 2) - 1 : 0;
  const int dw = (p.mode == 0) ? (tap & 3) - 1 : 0;
  const int SI = (p.mode == 0) ? 2 : 1;

  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; ++s) { mbar_init(&b_full[s], kWProducerWarps / kHiGroups); mbar_init(&b_empty[s], 1); }  // one group's warps
    for (int s = 0; s < S::kRawStages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < S::kAStages; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&tmap_lo);
  if (warp == kMmaWarp) tmem_alloc<S::kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp < kProd0) {
    // ---------------------------------------------------------------- transposers: raw Lo tile -> TMEM lane (= channel)
    const int wq = warp & 3;
    const uint32_t grp = smem_u32(smem) + S::kRawOff + wq * (kWK * 128);  // this warp's 32-channel group
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + S::kACol0;
    const uint32_t cword = (lane & 3) * 4, cchunk = lane >> 2;
    for (int it = warp >> 2; it < ksteps; it += TG) {
      const int s = it % S::kRawStages, t = it % S::kAStages;
      const uint32_t pars = (it / S::kRawStages) & 1, part = (it / S::kAStages) & 1;
      mbar_wait(&raw_full[s], pars);
      float x[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const uint32_t addr = grp + s * S::kRawBytes + r * 128 + ((cchunk ^ (r & 7)) << 4) + cword;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[r]) : "r"(addr));
      }
      float hi[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) hi[j] = tf32_round_fast(x[j]);   // consumes every loaded word: the loads have landed
      fence_proxy_async_smem();                    // generic-proxy reads before the next TMA write of this stage
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[s]);
      mbar_wait(&a_empty[t], part ^ 1);
      tc_fence_after_sync();
      tmem_st_x32(t_lane + t * 64, hi);
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] -= hi[j];
      tmem_st_x32(t_lane + t * 64 + 32, x);
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[t]);
    }
  } else if (warp < kTmaWarp) {
    // ---------------------------------------------------------------- Hi producers (shared memory), then epilogue
    // kHiGroups groups of warps take K steps round-robin.  A group's step: wait for the stage -> split + store the tile
    // it holds in registers -> fence.proxy.async -> arrive -> ONLY THEN issue the global loads of its next step.  The
    // order matters: fence.proxy.async is a MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS, i.e. it waits for EVERY
    // outstanding memory operation of the thread; with loads prefetched several K steps ahead (round 1) each step
    // stalled for a full L2 round trip (~1.2 us per K step measured, 3x the MMA time).  Here no load is in flight at the
    // fence, and a group's load latency (~1.3 us under load) is covered by the other groups' steps.
    static_assert(kHiGroups <= S::kStages, "a producer group may be at most one barrier phase ahead on a stage");
    constexpr int kGroupWarps = kWProducerWarps / kHiGroups, kGroupThreads = 32 * kGroupWarps;
    const int grp = (warp - kProd0) / kGroupWarps;
    const int tid = threadIdx.x - (kProd0 + kGroupWarps * grp) * 32;  // inside the group
    const uint32_t smem0 = smem_u32(smem);
    constexpr int kBChunksPerRow = BN / 4;
    constexpr int kBRowsPerPass = kGroupThreads / kBChunksPerRow;
    constexpr int kBPasses = kWK / kBRowsPerPass;
    static_assert(kBRowsPerPass >= 1 && kBPasses * kBRowsPerPass == kWK, "Hi tile split over one producer group");
    const int b_cidx = tid % kBChunksPerRow;
    const int b_row0 = tid / kBChunksPerRow;
    const int hw_l = p.Hl * p.Wl;
    auto b_soff = [&](int i) {
      const int r = b_row0 + kBRowsPerPass * i;
      return static_cast<uint32_t>((b_cidx >> 3) * (kWK * 128) + r * 128) + swz32(b_cidx & 7, r);
    };
    float4 bbuf[kBPasses];
    auto issue_loads = [&](int it) {
      const int pbase = pix0 + it * kWK;
#pragma unroll
      for (int i = 0; i < kBPasses; ++i) {
        const int pix = pbase + b_row0 + kBRowsPerPass * i;
        bool ok = pix < pix1;
        size_t off = 0;
        if (ok) {
          const int img = static_cast<int>((static_cast<unsigned long long>(pix) * p.magic_hw) >> 40);
          const int rem = pix - img * hw_l;
          const int gi = static_cast<int>((static_cast<unsigned long long>(rem) * p.magic_w) >> 40);
          const int gj = rem - gi * p.Wl;
          const int sh = gi * SI + dh, sw = gj * SI + dw;
          ok = sh >= 0 && sh < p.Hh && sw >= 0 && sw < p.Wh;
          off = (static_cast<size_t>((img * p.Hh + sh) * p.Wh + sw)) * p.C2 + c2_0 + b_cidx * 4;
        }
        bbuf[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.hi + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (grp < ksteps) issue_loads(grp);
    for (int it = grp; it < ksteps; it += kHiGroups) {
      const int s = it % S::kStages;
      const uint32_t par = (it / S::kStages) & 1;
      mbar_wait(&b_empty[s], par ^ 1);
      const uint32_t stage = smem0 + s * S::kStageBytes;
#pragma unroll
      for (int i = 0; i < kBPasses; ++i) store_split<true>(stage + b_soff(i), S::kBBytes, bbuf[i]);
      fence_proxy_async_smem();   // no global load of this thread is in flight here (see above)
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[s]);
      if (it + kHiGroups < ksteps) issue_loads(it + kHiGroups);
    }
    // epilogue: TMEM -> partial[split][tap]
    if (ksteps > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
    }
    const int q = warp & 3, half = (warp - kProd0) >> 2;
    constexpr int kColsPerHalf = BN / 2;
    const int c1 = c1_0 + q * 32 + lane;
    const int taps = gridDim.z / p.splits;
    float* orow = p.partial + (static_cast<size_t>(split * taps + tap) * p.C1 + c1) * p.C2 + c2_0;
#pragma unroll 1
    for (int c16 = 0; c16 < kColsPerHalf; c16 += 16) {
      const int colx = half * kColsPerHalf + c16;
      float v[16];
      if (ksteps > 0) {
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + colx, v);
        float tt[16];
#pragma unroll
        for (int a = 1; a < S::kAccs; ++a) {
          tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + colx, tt);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += tt[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (c1 < p.C1) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(orow + colx + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before_sync();
  } else if (warp == kTmaWarp) {
    // ---------------------------------------------------------------- Lo TMA issuer: 4 boxes of 32 pixels x 32 channels
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      int s = 0;
      uint32_t par = 0;
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&raw_empty[s], par ^ 1);
        mbar_arrive_expect_tx(&raw_full[s], S::kRawBytes);
        const uint32_t dst = smem_u32(smem) + S::kRawOff + s * S::kRawBytes;
#pragma unroll
        for (int g = 0; g < kWM / 32; ++g)
          tma_load_2d(dst + g * (kWK * 128), &tmap_lo, &raw_full[s], c1_0 + 32 * g, pix0 + it * kWK);
        if (++s == S::kRawStages) { s = 0; par ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer: A from TMEM, B (MN-major) from smem
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      constexpr uint32_t idesc = make_idesc_tf32(kWM, BN, 0, 1);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem), p.lbo_b, p.sbo_b, 1);
      constexpr uint64_t kStageStep = S::kStageBytes >> 4, kLoStep = S::kBBytes >> 4;
      const uint32_t acc_corr = tmem_base + S::kMain * BN;
      int s = 0, t = 0;
      uint32_t par = 0, part = 0;
      uint64_t db0 = bdesc0;
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&a_full[t], part);
        mbar_wait(&b_full[s], par);
        tc_fence_after_sync();
        const uint32_t acc = it != 0 ? 1u : 0u;
        const uint32_t a_hi0 = tmem_base + S::kACol0 + t * 64;
#pragma unroll
        for (int k = 0; k < kWK / 8; ++k) {
          const uint32_t a_hi = a_hi0 + 8 * k, a_lo = a_hi + 32;
          const uint64_t db = db0 + 64 * k;  // 8 pixels = two 512-byte k-atoms
          umma_tf32_ts(acc_corr, a_lo, db, idesc, k == 0 ? acc : 1u);
          umma_tf32_ts(acc_corr, a_hi, db + kLoStep, idesc, 1u);
          umma_tf32_ts(tmem_base + (k % S::kMain) * BN, a_hi, db, idesc, k < S::kMain ? acc : 1u);
        }
        umma_commit(&a_empty[t]);
        umma_commit(&b_empty[s]);
        db0 += kStageStep;
        if (++s == S::kStages) { s = 0; par ^= 1; db0 = bdesc0; }
        if (++t == S::kAStages) { t = 0; part ^= 1; }
      }
      if (ksteps > 0) umma_commit(tmem_full_bar);
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}

template <int BN, int TG>
static int launch_wgrad_ta_tg(const CUtensorMap& tmap_lo, const WgradParams& p, dim3 grid, cudaStream_t st) {
  using S = WgradTaSmem<BN>;
  static_assert(S::kDynamic <= 227 * 1024, "shared memory per CTA");
  MDGAN_CUDA(configure_smem_once(wgrad_gemm_ta_kernel<BN, TG>, S::kDynamic));
  MDGAN_LAUNCH((wgrad_gemm_ta_kernel<BN, TG>), grid, dim3(wta_threads<TG>()), S::kDynamic, st, tmap_lo, p);
  return 0;
}

// MDGAN_WGRAD_TG = 1 (default) | 2: groups of transposer warps (see wgrad_gemm_ta_kernel).  Measured on B200 (round 2):
// no difference in kernel time, bit-identical results; the default stays at one group.
static int wgrad_tg() {
  static const int tg = [] {
    const char* e = getenv("MDGAN_WGRAD_TG");
    const int v = e ? atoi(e) : 1;
    return v < 1 ? 1 : (v > 2 ? 2 : v);
  }();
  return tg;
}

template <int BN>
static int launch_wgrad_ta(const CUtensorMap& tmap_lo, const WgradParams& p, dim3 grid, cudaStream_t st) {
  return wgrad_tg() == 1 ? launch_wgrad_ta_tg<BN, 1>(tmap_lo, p, grid, st) : launch_wgrad_ta_tg<BN, 2>(tmap_lo, p, grid, st);
}

// MDGAN_WGRAD_TA = 1 (default) | 0: Lo operand by TMA into tensor memory (tf32x3 only; same arithmetic, same bits).
static bool wgrad_ta_enabled() {
  static const bool on = [] {
    const char* e = getenv("MDGAN_WGRAD_TA");
    return e ? e[0] != '0' : true;
  }();
  return on;
}

template <int BN, int STAGES, bool X3>
static int launch_wgrad(const WgradParams& p, dim3 grid, cudaStream_t st) {
  using S = WgradSmem<BN, STAGES, X3>;
  static_assert(S::kDynamic <= 227 * 1024, "shared memory per CTA");
  MDGAN_CUDA(configure_smem_once(wgrad_gemm_kernel<BN, STAGES, X3>, S::kDynamic));
  MDGAN_LAUNCH((wgrad_gemm_kernel<BN, STAGES, X3>), grid, dim3(kWThreads), S::kDynamic, st, p);
  return 0;
}

// Tile width of mdgan_wgrad_gemm for a given C2 (both precisions): 128 when it divides C2, else 64.
static int wgrad_bn(int C2) { return C2 % 128 == 0 ? 128 : 64; }

}  // namespace mdgan

using namespace mdgan;

// Number of split-K partial slices mdgan_wgrad_gemm will write for this problem (the caller sizes `partial`
// as splits * taps * C1 * C2 floats).
extern "C" int mdgan_wgrad_splits(int n_img, int Hl, int Wl, int C1, int C2, int mode) {
  const int P = n_img * Hl * Wl;
  const int taps = mode == 0 ? 16 : 1;
  const int bn = wgrad_bn(C2);
  const int tiles = (C1 / kWM) * (C2 / bn) * taps;
  int splits = tiles >= 148 ? 1 : 148 / tiles;  // one wave of CTAs: never spill a few tiles into a second wave
  const int max_splits = (P + 4 * kWK - 1) / (4 * kWK);  // at least 128 pixels per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

extern "C" int mdgan_wgrad_gemm(const float* lo, const float* hi, float* partial, int n_img, int Hl, int Wl, int C1,
                                int C2, int mode, int splits, int precision, void* stream) {
  if (!lo || !hi || !partial) return MDGAN_ERR_BAD_ARG;
  if (C1 % kWM != 0 || C2 % 64 != 0 || (mode != 0 && mode != 2) || splits < 1) return MDGAN_ERR_UNSUPPORTED;
  WgradParams p{};
  p.lo = lo; p.hi = hi; p.partial = partial;
  p.n_img = n_img; p.Hl = Hl; p.Wl = Wl;
  p.Hh = mode == 0 ? 2 * Hl : Hl;
  p.Wh = mode == 0 ? 2 * Wl : Wl;
  p.C1 = C1; p.C2 = C2; p.mode = mode;
  p.P = n_img * Hl * Wl;
  p.splits = splits;
  p.pix_per_split = ceil_div(ceil_div(p.P, splits), kWK) * kWK;
  if (p.P >= (1 << 28) / (Hl * Wl > 0 ? 1 : 1) && static_cast<long long>(p.P) * Hl * Wl >= (1LL << 40)) return MDGAN_ERR_UNSUPPORTED;
  p.magic_hw = ((1ULL << 40) + static_cast<unsigned long long>(Hl * Wl) - 1) / static_cast<unsigned long long>(Hl * Wl);
  p.magic_w = ((1ULL << 40) + static_cast<unsigned long long>(Wl) - 1) / static_cast<unsigned long long>(Wl);
  p.lbo_a = p.lbo_b = kWK * 128;  // distance between 32-channel groups (encoding confirmed on hardware, round 1)
  p.sbo_a = p.sbo_b = 512;        // distance between 4-pixel swizzle atoms
  const int taps = mode == 0 ? 16 : 1;
  const int bn = wgrad_bn(C2);
  dim3 grid(C1 / kWM, C2 / bn, taps * splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision != 0 && precision != 1) return MDGAN_ERR_BAD_ARG;
  if (precision == 1 && wgrad_ta_enabled()) {
    CUtensorMap tmap_lo;
    const int rc = get_tmap_2d_f32(lo, static_cast<uint64_t>(p.P), static_cast<uint64_t>(C1), kWK, &tmap_lo);
    if (rc != 0) return rc;
    return bn == 128 ? launch_wgrad_ta<128>(tmap_lo, p, grid, st) : launch_wgrad_ta<64>(tmap_lo, p, grid, st);
  }
  if (precision == 1) return bn == 128 ? launch_wgrad<128, 3, true>(p, grid, st) : launch_wgrad<64, 4, true>(p, grid, st);
  return bn == 128 ? launch_wgrad<128, 6, false>(p, grid, st) : launch_wgrad<64, 6, false>(p, grid, st);
}
