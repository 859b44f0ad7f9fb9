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
// Two kernels:
//   conv_gemm_kernel    (single-pass tf32, and the tf32x3 predecessor MDGAN_CONV_TA=0): both operands in shared memory.
//     warps 0-7 : A producers (global -> registers kPrefetch K steps ahead -> TF32 hi/lo split -> one swizzled
//                 st.shared per operand), then the epilogue
//     warp  8   : B producer (TMA 2D tiled load, SWIZZLE_128B, mbarrier complete_tx)
//     warp  9   : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem stages
//   conv_gemm_ta_kernel (tf32x3 default): activation operand in tensor memory, see the comment above that kernel.
//
// Precision modes (template X3):
//   tf32   : one tcgen05.mma per K=8 slice; operands are fp32 words that their producers rounded to TF32.
//   tf32x3 : error-compensated split ("3xTF32"): every operand x is held as hi = tf32(x) and lo = x - hi; the
//            product is accumulated as lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator, which restores ~fp32
//            accuracy (the parity mode: plain TF32 through BatchNorm + (Leaky)ReLU gates is only good to a few
//            percent on gradients, exactly like cuDNN's TF32 path -- see DESIGN.md).  The A producers split their
//            own cp.async'd chunks in shared memory; the weight pack kernel pre-splits B.
#include <cstdlib>

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
  int accumulate;     // 1: dst += result (NCHW outputs only; sums feedbacks of workers sharing a batch)
  int lo_row_offset;  // tf32x3: row offset of the `lo` half of the packed weights
  int rows_per_tile;  // GEMM rows a CTA owns: 128, or fewer when a TMA box of whole images does not fill 128 rows (the
                      // remaining TMEM lanes carry garbage that is never stored)
  const float* gate;  // optional, NHWC like dst: dst = result * act'(gate) (backward of the activation that produced
  int gate_act;       //   `gate`, fused into the data-gradient GEMM that feeds it; 1 ReLU, 2 LeakyReLU(gate_slope))
  float gate_slope;
  float* bn_partial;  // optional [phases][row tiles][2][N_pad]: per-CTA column sums / sums of squares of the stored
                      //   values (the BatchNorm statistics of the layer, finalized by mdgan_bn_finalize)
  // BatchNorm-backward fusion (data-gradient GEMMs whose output is the gradient da of a BatchNorm+activation output):
  // with bnb_z (the BatchNorm INPUT, NHWC like dst) and bnb_stats ([groups][4][N]: mean, invstd, scale, shift) the
  // epilogue stores dy = da * act'(z*scale + shift) instead of da and bn_partial receives the column sums of dy and
  // of dy * xhat (xhat = (z - mean) * invstd): the reduction pass of the BatchNorm backward.
  const float* bnb_z;
  const float* bnb_stats;
  int bnb_act;
  float bnb_slope;
  int bnb_rows_per_group;  // GEMM rows (low-resolution positions) per BatchNorm pass; a CTA never straddles two
};

constexpr int kBM = 128;
constexpr int kBK = 32;  // fp32 elements per K step = 128 bytes
constexpr int kNumProducerWarps = 8;
constexpr int kThreads = (kNumProducerWarps + 2) * 32;
#ifndef MDGAN_CONV_PREFETCH
#define MDGAN_CONV_PREFETCH 4
#endif
constexpr int kPrefetch = MDGAN_CONV_PREFETCH;  // K steps of activation loads kept in flight in registers per producer thread

template <int BN, int STAGES, bool X3>
struct ConvGemmSmem {
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kHalfBytes = kABytes + kBBytes;             // [A_hi | B_hi] then (X3) [A_lo | B_lo]
  static constexpr int kStageBytes = (X3 ? 2 : 1) * kHalfBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024 B alignment
  // tf32x3 keeps several TMEM accumulators: the tensor core adds into its fp32 accumulator with truncation, so a
  // long accumulation chain drifts by ~3e-8 per add (measured: 1.5e-5 at K = 4096 with one accumulator).  The four
  // K=8 slices of a K step go to kMain accumulators (slice k -> accumulator k % kMain, a compile-time constant in
  // the unrolled issue loop), the two small correction products to their own accumulator, and the epilogue sums
  // them with ordinary (round-to-nearest) fp32 adds.
  static constexpr int kMain = X3 ? (BN > 64 ? 2 : 4) : 1;
  static constexpr int kAccs = X3 ? kMain + 1 : 1;
  static constexpr uint32_t kTmemCols = kAccs * BN <= 32 ? 32 : kAccs * BN <= 64 ? 64 : kAccs * BN <= 128 ? 128
                                        : kAccs * BN <= 256 ? 256 : 512;
  static_assert(kAccs * BN <= 512, "TMEM holds 512 columns");
};

__device__ __forceinline__ float round_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Epilogue of one 128 x BN tile: warp (q = TMEM lane quarter, half = column half) reads its 32 rows x BN/2 columns of
// the kAccs accumulators (summed with round-to-nearest fp32 adds), applies bias / tanh / TF32 rounding / the fused
// activation-backward gate and stores NHWC (float4) or scatters NCHW (optionally accumulating).
// Sum of v[0..16) over the 32 lanes of the warp, column by column, with 16 shuffles instead of 80: at every step a lane
// keeps half of its columns and hands the other half to its partner (xor 16, 8, 4, 2), then the pair is combined
// (xor 1).  On return lane l holds, in v[0], the 32-row total of column (l >> 1) & 15.  Fixed tree: deterministic.
__device__ __forceinline__ void warp_column_sums16(float (&v)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int mask = 16 >> step, cnt = 8 >> step;
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < cnt) {
        const float send = upper ? v[j] : v[j + cnt];
        const float keep = upper ? v[j + cnt] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// red: shared-memory scratch [4 row quarters][2 statistics][BN] for the fused BatchNorm statistics (STATS only).
template <int BN, int kAccs, bool STATS>
__device__ __forceinline__ void conv_epilogue(const ConvGemmParams& p, uint32_t tmem_base, int q, int half, int lane,
                                              int m0, int n0, int ph, int pw, float* red = nullptr) {
  constexpr int kColsPerHalf = (BN >= 32) ? BN / 2 : BN;
  if (BN >= 32 || half == 0) {
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool ok = m < p.M && row < p.rows_per_tile;
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
      if (kAccs > 1) {
        float t[16];
#pragma unroll
        for (int a = 1; a < kAccs; ++a) {
          tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + col, t);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += t[j];
        }
      }
        const int nbase = n0 + col;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = v[j];
          if (p.bias != nullptr && nbase + j < p.N) x += __ldg(p.bias + nbase + j);
          if (p.act == 1) x = tanhf(x);
          if (p.round_tf32) x = round_to_tf32(x);
          v[j] = x;
        }
        if (STATS && p.bn_partial != nullptr) {  // warp-uniform: every lane takes part in the shuffles
          float s1[16], s2[16];
          if (p.bnb_z != nullptr) {
            // dy = da * act'(y) and dy * xhat from the BatchNorm input z at the same NHWC position (N % 16 == 0)
            const float* st = p.bnb_stats + static_cast<size_t>(m0 / p.bnb_rows_per_group) * 4 * p.N + nbase;
            const float* zr = p.bnb_z + (static_cast<size_t>((img * Ho + oh) * Wo + ow)) * p.N + nbase;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 z4 = ok ? __ldg(reinterpret_cast<const float4*>(zr + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 mean = __ldg(reinterpret_cast<const float4*>(st + j));
              const float4 istd = __ldg(reinterpret_cast<const float4*>(st + p.N + j));
              const float4 sc = __ldg(reinterpret_cast<const float4*>(st + 2 * p.N + j));
              const float4 sh = __ldg(reinterpret_cast<const float4*>(st + 3 * p.N + j));
              const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, mm4[4] = {mean.x, mean.y, mean.z, mean.w};
              const float ii[4] = {istd.x, istd.y, istd.z, istd.w}, cc[4] = {sc.x, sc.y, sc.z, sc.w};
              const float hh[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float y = fmaf(zz[t], cc[t], hh[t]);
                const float g = y > 0.f ? 1.f : (p.bnb_act == 2 ? p.bnb_slope : (p.bnb_act == 1 ? 0.f : 1.f));
                const float dy = ok ? v[j + t] * g : 0.f;
                v[j + t] = dy;
                s1[j + t] = dy;
                s2[j + t] = dy * ((zz[t] - mm4[t]) * ii[t]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { s1[j] = ok ? v[j] : 0.f; s2[j] = s1[j] * s1[j]; }
          }
          warp_column_sums16(s1, lane);
          warp_column_sums16(s2, lane);
          if ((lane & 1) == 0) {
            red[(q * 2 + 0) * BN + col + (lane >> 1)] = s1[0];
            red[(q * 2 + 1) * BN + col + (lane >> 1)] = s2[0];
          }
        }
        if (!ok) continue;
        if (!p.out_nchw) {
          const size_t oidx = (static_cast<size_t>((img * Ho + oh) * Wo + ow)) * p.N + nbase;
          float* o = p.dst + oidx;
          if (nbase + 16 <= p.N && (p.N & 3) == 0) {
            if (p.gate != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p.gate + oidx + j));
                const float neg = p.gate_act == 2 ? p.gate_slope : 0.f;
                v[j] *= a.x > 0.f ? 1.f : neg;
                v[j + 1] *= a.y > 0.f ? 1.f : neg;
                v[j + 2] *= a.z > 0.f ? 1.f : neg;
                v[j + 3] *= a.w > 0.f ? 1.f : neg;
                if (p.round_tf32) {
                  v[j] = round_to_tf32(v[j]); v[j + 1] = round_to_tf32(v[j + 1]);
                  v[j + 2] = round_to_tf32(v[j + 2]); v[j + 3] = round_to_tf32(v[j + 3]);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nbase + j < p.N) {
                float x = v[j];
                if (p.gate != nullptr) {
                  x *= __ldg(p.gate + oidx + j) > 0.f ? 1.f : (p.gate_act == 2 ? p.gate_slope : 0.f);
                  if (p.round_tf32) x = round_to_tf32(x);
                }
                o[j] = x;
              }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nbase + j < p.N) {
              float* o = p.dst + (static_cast<size_t>(img * p.N + nbase + j) * Ho + oh) * Wo + ow;
              *o = p.accumulate ? (*o + v[j]) : v[j];
            }
        }
      }
    }
}

template <int BN, int STAGES, bool X3>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const ConvGemmParams p) {
  using S = ConvGemmSmem<BN, STAGES, X3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * p.rows_per_tile;
  const int n0 = blockIdx.y * BN;
  const int phase = blockIdx.z;  // UP: output parity phase (ph*2 + pw)
  const int ph = phase >> 1, pw = phase & 1;

  const int taps = (p.mode == 0) ? 16 : (p.mode == 1 ? 4 : 1);
  const int cchunks = p.C / kBK;
  const int ksteps = taps * cchunks;

  pdl_trigger();  // the next kernel's CTAs may take SMs as they drain; it waits for this grid before touching memory
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kNumProducerWarps + 1);  // one arrive per producer warp + the TMA thread
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 9) tmem_alloc<S::kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // barriers, TMEM and the descriptor prefetch overlapped the previous kernel's tail; its outputs are read below

  if (warp < kNumProducerWarps) {
    // ------------------------------------------------------------------ A producers
    // Each thread owns the 16-byte chunk `chunk` of rows row_in + 32 i.  Loads go global -> registers, kPrefetch K
    // steps ahead (their latency never sits on the critical path), are rounded / split into TF32 hi and lo parts in
    // registers and stored once to the swizzled stage -- one shared-memory pass per operand.
    const int chunk = threadIdx.x & 7;
    const int row_in = threadIdx.x >> 3;  // 0..31
    const int SI = (p.mode == 0) ? 2 : 1;
    int base_off[4], sh0[4], sw0[4];
    bool row_ok[4];
    uint32_t soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row_in + 32 * i;
      const int m = m0 + r;
      row_ok[i] = m < p.M;
      const int mm = row_ok[i] ? m : 0;
      const int img = mm / (p.Hg * p.Wg);
      const int rem = mm - img * (p.Hg * p.Wg);
      const int gi = rem / p.Wg, gj = rem - gi * p.Wg;
      sh0[i] = gi * SI;
      sw0[i] = gj * SI;
      base_off[i] = ((img * p.Hs + sh0[i]) * p.Ws + sw0[i]) * p.C + chunk * 4;
      soff[i] = r * 128 + ((chunk ^ (r & 7)) << 4);
    }
    const uint32_t a_smem0 = smem_u32(smem);
    int tap_l = 0, cc_l = 0;  // coordinates of the next K step to load
    auto issue_loads = [&](float4(&buf)[4]) {
      int dh, dw;
      if (p.mode == 0) { dh = (tap_l >> 2) - 1; dw = (tap_l & 3) - 1; }
      else if (p.mode == 1) { dh = ph - (tap_l >> 1); dw = pw - (tap_l & 1); }
      else { dh = 0; dw = 0; }
      const int tap_off = (dh * p.Ws + dw) * p.C + cc_l * kBK;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sh = sh0[i] + dh, sw = sw0[i] + dw;
        const bool ok = row_ok[i] && sh >= 0 && sh < p.Hs && sw >= 0 && sw < p.Ws;
        buf[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.src + base_off[i] + tap_off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++cc_l == cchunks) { cc_l = 0; ++tap_l; }
    };
    float4 buf[kPrefetch][4];
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u)
      if (u < ksteps) issue_loads(buf[u]);
    int s = 0;
    uint32_t par = 0;
    for (int it0 = 0; it0 < ksteps; it0 += kPrefetch) {
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) {
        const int it = it0 + u;
        if (it < ksteps) {
          mbar_wait(&empty_bar[s], par ^ 1);
          const uint32_t a_stage = a_smem0 + s * S::kStageBytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = buf[u][i];
            const float hx = tf32_round_fast(v.x), hy = tf32_round_fast(v.y), hz = tf32_round_fast(v.z),
                        hw = tf32_round_fast(v.w);
            sts128(a_stage + soff[i], hx, hy, hz, hw);
            if (X3)  // lo = x - hi (exact in fp32); the tensor core truncates it to TF32 (2^-21 relative to x) and
                     // the dropped lo*lo term is ~2^-22 relative
              sts128(a_stage + S::kHalfBytes + soff[i], v.x - hx, v.y - hy, v.z - hz, v.w - hw);
          }
          fence_proxy_async_smem();  // every thread publishes its own stores to the async proxy ...
          __syncwarp();              // ... the warp agrees they are all done, and one lane arrives for the 32
          if (lane == 0) mbar_arrive(&full_bar[s]);
          if (it + kPrefetch < ksteps) issue_loads(buf[u]);
          if (++s == STAGES) { s = 0; par ^= 1; }
        }
      }
    }

    // ------------------------------------------------------------------ epilogue
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    conv_epilogue<BN, X3 ? S::kAccs : 1, false>(p, tmem_base, warp & 3, warp >> 2, lane, m0, n0, ph, pw);
    tc_fence_before_sync();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ B producer (TMA)
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      const int row0 = (p.mode == 1 ? phase * p.N_pad : 0) + n0;
      int s = 0;
      uint32_t par = 0;
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&empty_bar[s], par ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], (X3 ? 2 : 1) * S::kBBytes);
        const uint32_t b_dst = smem_u32(smem + s * S::kStageBytes + S::kABytes);
        tma_load_2d(b_dst, &tmap_w, &full_bar[s], it * kBK, row0);
        if (X3) tma_load_2d(b_dst + S::kHalfBytes, &tmap_w, &full_bar[s], it * kBK, row0 + p.lo_row_offset);
        if (++s == STAGES) { s = 0; par ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    // One thread; the loop carries nothing but the stage descriptor and the barrier parity, every other quantity
    // (K-slice offset inside the swizzle atom, hi/lo offset, accumulator column) is an immediate of the unrolled
    // body: a tcgen05.mma 128 x BN x 8 costs max(44, BN/2) clk (tools/micro/umma_micro.cu), an issue loop that
    // rebuilds descriptors or takes a modulo per MMA costs more than that.
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 0, 0);
      const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      constexpr uint64_t kStageStep = S::kStageBytes >> 4, kBStep = S::kABytes >> 4, kLoStep = S::kHalfBytes >> 4;
      int s = 0;
      uint32_t par = 0;
      uint64_t da0 = desc0;
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&full_bar[s], par);
        tc_fence_after_sync();
        const uint32_t acc = it != 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint64_t da = da0 + 2 * k, db = da + kBStep;
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
      umma_commit(tmem_full_bar);
    }
  }
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}

// ================================================================================================================
// tf32x3 variant with the ACTIVATION operand in tensor memory.  The shared-memory port bounds the kernel above (per K
// step: 96 KB of MMA operand reads + 32 KB of producer stores + 32 KB of TMA writes against 768 clk of MMA time,
// profiles/r01_ncu_full_conv_wgrad.md).  Here the activations cross shared memory once, as raw fp32 (16 KB stored,
// 16 KB read back row-wise), are split into TF32 hi / lo in registers by four "transposer" warps that own one TMEM lane
// quarter each, and land in TMEM (tcgen05.st) where the twelve MMAs of the K step read them for free; only the weights
// (TMA) stay on the port: 16 + 16 + 32 + 48 = 112 KB per K step instead of 160 KB.
//   warps 0-3  : transposers: raw stage (row-wise, swizzle-aware LDS) -> hi/lo -> TMEM A stage
//   warps 4-11 : loaders (global -> registers kPrefetch K steps ahead -> raw fp32 stage), then the epilogue
//   warp  12   : B producer (TMA, hi + lo)        warp 13: TMEM allocator + single-thread MMA issuer
//   warp  14   : (TMA_A only) activation-tile TMA issuer
// With the TMA-fed activation tile the loader warps have nothing to load, so TG = 2 lets warps 4-7 act as a second
// transposer group (group g takes K steps g, g+2, ...).  Measured: no effect (see conv_tg()), the option is kept for the
// cross-variant bit-identity stress (tools/stress_conv.py).
constexpr int kTaThreads = (4 + kNumProducerWarps + 2) * 32;
constexpr int kTaThreadsTma = kTaThreads + 32;

template <int BN>
struct ConvTaSmem {
  static constexpr int kRawBytes = kBM * 128;
  static constexpr int kRawStages = 4;
  static constexpr int kBBytes = BN * 128;              // one of hi / lo
  static constexpr int kBStageBytes = 2 * kBBytes;      // [B_hi | B_lo]
  // MDGAN_CONV_DEEP (experiment build, `make VARIANT=deep`): 128-wide tiles with ONE main accumulator (+ the correction
  // accumulator) so that four TMEM stages of the activation operand and four weight stages fit -- tests whether the
  // depth of the operand pipeline, rather than bandwidth, paces the K step.
#ifdef MDGAN_CONV_DEEP
  static constexpr int kBStages = 4;
  static constexpr int kAStages = BN > 64 ? 4 : 3;
#else
  static constexpr int kBStages = 4;                    // 4 x 32 KB weight stages + 4 x 16 KB raw stages = 192 KB at BN = 128
  static constexpr int kAStages = BN > 64 ? 2 : 3;      // TMEM stages of [A_hi (32 columns) | A_lo (32 columns)]
#endif
  static constexpr int kBOff = kRawStages * kRawBytes;
  static constexpr int kBarOffset = kBOff + kBStages * kBStageBytes;
  static constexpr int kNumBars = 2 * kRawStages + 2 * kBStages + 2 * kAStages + 1;
  static constexpr int kRedOff = kBarOffset + kNumBars * 8 + 16;   // [4][2][BN] floats: fused BatchNorm statistics
  static constexpr int kTotal = kRedOff + 4 * 2 * BN * 4;
  static constexpr int kDynamic = kTotal + 1024;
#ifdef MDGAN_CONV_DEEP
  static constexpr int kMain = BN > 64 ? 1 : 4;
#else
  static constexpr int kMain = BN > 64 ? 2 : 4;
#endif
  static constexpr int kAccs = kMain + 1;
  static constexpr int kACol0 = kAccs * BN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kACol0 + kAStages * 64 <= 512, "TMEM holds 512 columns");
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// TMA_A: the raw activation tile is fetched by ONE 4-D tiled TMA load per K step (tmap_a: box = the 128 im2col rows of
// one tap x 32 channels, element strides = the conv stride, out-of-bounds = zero padding) instead of by the eight
// loader warps: the tile then crosses the L1/shared-memory arrays once (the TMA write) instead of three times
// (L1 fill, L1 read, shared store).  Possible whenever 128 consecutive rows of the (n, i, j) grid form a box.
// UP2 (BN = 128, UP mode with 64 output channels): one CTA computes BOTH column parities (pw = 0, 1) of output-row parity
// ph = blockIdx.z.  The two parities read the same source rows and overlapping source columns: the six shifted
// activation tiles (dh in {ph-1, ph}) x (dw in {0, -1, +1}) replace the 2 x 4 tiles of two single-phase CTAs, the dw = 0
// tile feeds both parities with ONE 128-wide MMA per K slice ([pw0 | pw1] weights side by side, accumulators side by
// side in TMEM), the dw = -1 / +1 tiles feed pw = 0 / 1 with 64-wide MMAs.  Per K slice: 64 + 46 + 46 clk of MMA instead
// of 4 x 46, 25 % less transposer work, half as many CTAs with 24 instead of 16 K steps each.
template <int BN, bool TMA_A, int TG, bool UP2 = false>
__global__ void __launch_bounds__(TMA_A ? kTaThreadsTma : kTaThreads, 1)
conv_gemm_ta_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                    const ConvGemmParams p) {
  using S = ConvTaSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* raw_empty = raw_full + S::kRawStages;
  uint64_t* b_full = raw_empty + S::kRawStages;
  uint64_t* b_empty = b_full + S::kBStages;
  uint64_t* a_full = b_empty + S::kBStages;
  uint64_t* a_empty = a_full + S::kAStages;
  uint64_t* tmem_full_bar = a_empty + S::kAStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* red = reinterpret_cast<float*>(smem + S::kRedOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * p.rows_per_tile;
  const int n0 = blockIdx.y * BN;
  static_assert(!UP2 || (BN == 128 && TMA_A), "UP2 pairs two 64-wide phases in the 128-wide TMA-fed kernel");
  const int phase = blockIdx.z;
  const int ph = UP2 ? phase : phase >> 1, pw = phase & 1;
  const int taps = (p.mode == 0) ? 16 : (p.mode == 1 ? 4 : 1);
  const int cchunks = p.C / kBK;
  const int ksteps = UP2 ? 6 * cchunks : taps * cchunks;

  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kRawStages; ++s) { mbar_init(&raw_full[s], TMA_A ? 1 : kNumProducerWarps); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < S::kBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < S::kAStages; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 12 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (TMA_A && warp == 14 && lane == 0) tma_prefetch_desc(&tmap_a);
  if (warp == 13) tmem_alloc<S::kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  static_assert(TG >= 1 && TG <= 2 && (TMA_A || TG == 1), "transposer groups: the loader warps double as transposers only when TMA feeds the tile");
  // a group's consecutive K steps are TG apart: it must not get more than one barrier phase ahead on any stage (an
  // mbarrier parity wait cannot tell two phases apart) -- three groups over two TMEM stages dead-locked in round 2
  static_assert(TG <= S::kAStages && TG <= S::kRawStages, "transposer groups must not outnumber the stages they cycle through");
  if (warp < 12) {
    if (warp < 4 * TG) {
      // ---------------------------------------------------------------- transposers: raw stage -> TMEM A stage
      const int r = (warp & 3) * 32 + lane;  // tile row == TMEM lane
      const uint32_t row_base = smem_u32(smem) + r * 128;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + S::kACol0;
      for (int it = warp >> 2; it < ksteps; it += TG) {
        const int s = it % S::kRawStages, t = it % S::kAStages;
        const uint32_t pars = (it / S::kRawStages) & 1, part = (it / S::kAStages) & 1;
        mbar_wait(&raw_full[s], pars);
        float x[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(row_base + s * S::kRawBytes + ((c ^ (r & 7)) << 4));
          x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
        float hi[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) hi[j] = tf32_round_fast(x[j]);   // consumes every loaded word: the loads have landed
        fence_proxy_async_smem();                    // order the generic-proxy reads before the next TMA write of the stage
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);   // the row is in registers: the stage may be refilled
        mbar_wait(&a_empty[t], part ^ 1);            // the MMAs that read this TMEM stage have completed
        tc_fence_after_sync();
        tmem_st_x32(t_lane + t * 64, hi);
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] -= hi[j];  // lo = x - hi (exact); the tensor core truncates it to TF32
        tmem_st_x32(t_lane + t * 64 + 32, x);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[t]);
      }
    }
    if (warp >= 4) {
      // ---------------------------------------------------------------- loaders (raw fp32 tile), then epilogue
      if (!TMA_A) {
        const int tid = threadIdx.x - 128;
        const int chunk = tid & 7;
        const int row_in = tid >> 3;  // 0..31
        const int SI = (p.mode == 0) ? 2 : 1;
        int base_off[4], sh0[4], sw0[4];
        bool row_ok[4];
        uint32_t soff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = row_in + 32 * i;
          const int m = m0 + r;
          row_ok[i] = m < p.M;
          const int mm = row_ok[i] ? m : 0;
          const int img = mm / (p.Hg * p.Wg);
          const int rem = mm - img * (p.Hg * p.Wg);
          const int gi = rem / p.Wg, gj = rem - gi * p.Wg;
          sh0[i] = gi * SI;
          sw0[i] = gj * SI;
          base_off[i] = ((img * p.Hs + sh0[i]) * p.Ws + sw0[i]) * p.C + chunk * 4;
          soff[i] = r * 128 + ((chunk ^ (r & 7)) << 4);
        }
        const uint32_t raw0 = smem_u32(smem);
        int tap_l = 0, cc_l = 0;
        auto issue_loads = [&](float4(&buf)[4]) {
          int dh, dw;
          if (p.mode == 0) { dh = (tap_l >> 2) - 1; dw = (tap_l & 3) - 1; }
          else if (p.mode == 1) { dh = ph - (tap_l >> 1); dw = pw - (tap_l & 1); }
          else { dh = 0; dw = 0; }
          const int tap_off = (dh * p.Ws + dw) * p.C + cc_l * kBK;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int sh = sh0[i] + dh, sw = sw0[i] + dw;
            const bool ok = row_ok[i] && sh >= 0 && sh < p.Hs && sw >= 0 && sw < p.Ws;
            buf[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.src + base_off[i] + tap_off)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (++cc_l == cchunks) { cc_l = 0; ++tap_l; }
        };
        float4 buf[kPrefetch][4];
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u)
          if (u < ksteps) issue_loads(buf[u]);
        int s = 0;
        uint32_t par = 0;
        for (int it0 = 0; it0 < ksteps; it0 += kPrefetch) {
#pragma unroll
          for (int u = 0; u < kPrefetch; ++u) {
            const int it = it0 + u;
            if (it < ksteps) {
              mbar_wait(&raw_empty[s], par ^ 1);
              const uint32_t stage = raw0 + s * S::kRawBytes;
#pragma unroll
              for (int i = 0; i < 4; ++i) sts128(stage + soff[i], buf[u][i].x, buf[u][i].y, buf[u][i].z, buf[u][i].w);
              __syncwarp();
              if (lane == 0) mbar_arrive(&raw_full[s]);
              if (it + kPrefetch < ksteps) issue_loads(buf[u]);
              if (++s == S::kRawStages) { s = 0; par ^= 1; }
            }
          }
        }
      }
      __syncwarp();
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
      if (UP2)  // columns [64 half, 64 half + 64) are the 64 channels of column parity pw = half
        conv_epilogue<BN, S::kAccs, true>(p, tmem_base, warp & 3, (warp - 4) >> 2, lane, m0, -64 * ((warp - 4) >> 2), ph,
                                          (warp - 4) >> 2, red);
      else
        conv_epilogue<BN, S::kAccs, true>(p, tmem_base, warp & 3, (warp - 4) >> 2, lane, m0, n0, ph, pw, red);
      tc_fence_before_sync();
    }
  } else if (TMA_A && warp == 14) {
    // ------------------------------------------------------------------ activation tile producer (TMA, 4-D box)
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      const int SIa = (p.mode == 0) ? 2 : 1;
      const int n_start = m0 / (p.Hg * p.Wg);
      const int h_start = (m0 - n_start * (p.Hg * p.Wg)) / p.Wg;
      int s = 0, tap_l = 0, cc_l = 0;
      uint32_t par = 0;
      for (int it = 0; it < ksteps; ++it) {
        int dh, dw;
        if (UP2) { dh = ph - tap_l / 3; const int j = tap_l % 3; dw = j == 0 ? 0 : (j == 1 ? -1 : 1); }  // tap_l = 3 a + j
        else if (p.mode == 0) { dh = (tap_l >> 2) - 1; dw = (tap_l & 3) - 1; }
        else if (p.mode == 1) { dh = ph - (tap_l >> 1); dw = pw - (tap_l & 1); }
        else { dh = 0; dw = 0; }
        mbar_wait(&raw_empty[s], par ^ 1);
        mbar_arrive_expect_tx(&raw_full[s], p.rows_per_tile * 128);
        tma_load_4d(smem_u32(smem) + s * S::kRawBytes, &tmap_a, &raw_full[s], cc_l * kBK, dw, h_start * SIa + dh, n_start);
        if (++cc_l == cchunks) { cc_l = 0; ++tap_l; }
        if (++s == S::kRawStages) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------ B producer (TMA, hi + lo)
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      const int row0 = (p.mode == 1 ? phase * p.N_pad : 0) + n0;
      int s = 0;
      uint32_t par = 0;
      int tap_l = 0, cc_l = 0;  // UP2 only
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&b_empty[s], par ^ 1);
        const uint32_t b_dst = smem_u32(smem + S::kBOff + s * S::kBStageBytes);
        if (UP2) {
          // packed weights: row (2 ph + pw) * 64 + n, column (2 a + b) * C + c; tile rows 0-63 = pw 0, 64-127 = pw 1
          const int a = tap_l / 3, j = tap_l % 3;
          const int r0 = (2 * ph) * 64, r1 = (2 * ph + 1) * 64;
          const int k0 = (2 * a) * p.C + cc_l * kBK, k1 = (2 * a + 1) * p.C + cc_l * kBK;
          mbar_arrive_expect_tx(&b_full[s], j == 0 ? 4 * 64 * 128 : 2 * 64 * 128);
          if (j == 0 || j == 1) {   // pw = 0: b = 0 (dw = 0) or b = 1 (dw = -1)
            tma_load_2d(b_dst, &tmap_w, &b_full[s], j == 0 ? k0 : k1, r0);
            tma_load_2d(b_dst + S::kBBytes, &tmap_w, &b_full[s], j == 0 ? k0 : k1, r0 + p.lo_row_offset);
          }
          if (j == 0 || j == 2) {   // pw = 1: b = 1 (dw = 0) or b = 0 (dw = +1)
            tma_load_2d(b_dst + 64 * 128, &tmap_w, &b_full[s], j == 0 ? k1 : k0, r1);
            tma_load_2d(b_dst + S::kBBytes + 64 * 128, &tmap_w, &b_full[s], j == 0 ? k1 : k0, r1 + p.lo_row_offset);
          }
          if (++cc_l == cchunks) { cc_l = 0; ++tap_l; }
        } else {
          mbar_arrive_expect_tx(&b_full[s], S::kBStageBytes);
          tma_load_2d(b_dst, &tmap_w, &b_full[s], it * kBK, row0);
          tma_load_2d(b_dst + S::kBBytes, &tmap_w, &b_full[s], it * kBK, row0 + p.lo_row_offset);
        }
        if (++s == S::kBStages) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer (A from TMEM, B from shared memory)
    if (elect_one()) {  // one elected lane; keeps the role's code warp-uniform for ptxas (see ptx.cuh)
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 0, 0);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem + S::kBOff), 16, 1024);
      constexpr uint64_t kBStageStep = S::kBStageBytes >> 4, kLoStep = S::kBBytes >> 4;
      const uint32_t acc_corr = tmem_base + S::kMain * BN;
      int sb = 0, t = 0;
      uint32_t parb = 0, part = 0;
      uint64_t db0 = bdesc0;
      int tap_l = 0, cc_l = 0;  // UP2 only
      for (int it = 0; it < ksteps; ++it) {
        mbar_wait(&a_full[t], part);
        mbar_wait(&b_full[sb], parb);
        tc_fence_after_sync();
        const uint32_t acc = it != 0 ? 1u : 0u;   // UP2: step 0 is a dw = 0 tile, which writes all 128 columns
        const uint32_t a_hi0 = tmem_base + S::kACol0 + t * 64;
        // UP2: dw = 0 -> one 128-wide MMA over [pw 0 | pw 1]; dw = -1 -> 64-wide on the left half; +1 -> right half
        uint32_t idesc_s = idesc, col_off = 0;
        uint64_t db_off = 0;
        if (UP2) {
          const int j = tap_l % 3;
          if (j != 0) idesc_s = make_idesc_tf32(kBM, 64, 0, 0);
          if (j == 2) { col_off = 64; db_off = (64 * 128) >> 4; }
          if (++cc_l == cchunks) { cc_l = 0; ++tap_l; }
        }
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint32_t a_hi = a_hi0 + 8 * k, a_lo = a_hi + 32;
          const uint64_t db = db0 + db_off + 2 * k;
          umma_tf32_ts(acc_corr + col_off, a_lo, db, idesc_s, k == 0 ? acc : 1u);
          umma_tf32_ts(acc_corr + col_off, a_hi, db + kLoStep, idesc_s, 1u);
          umma_tf32_ts(tmem_base + (k % S::kMain) * BN + col_off, a_hi, db, idesc_s, k < S::kMain ? acc : 1u);
        }
        umma_commit(&a_empty[t]);
        umma_commit(&b_empty[sb]);
        db0 += kBStageStep;
        if (++sb == S::kBStages) { sb = 0; parb ^= 1; db0 = bdesc0; }
        if (++t == S::kAStages) { t = 0; part ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
  }
  __syncthreads();
  if (warp == 13) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
  if (p.bn_partial != nullptr && threadIdx.x < 2 * BN) {
    // fused BatchNorm statistics: the CTA's column sums (quarters added in a fixed order) -> its partial slice
    const int stat = threadIdx.x / BN, col = threadIdx.x - stat * BN;
    float t = (red[(0 * 2 + stat) * BN + col] + red[(1 * 2 + stat) * BN + col]) +
              (red[(2 * 2 + stat) * BN + col] + red[(3 * 2 + stat) * BN + col]);
    if (UP2) {  // columns c and 64 + c are channel c of the two column parities: one slice entry per channel
      if (col < 64) {
        t += (red[(0 * 2 + stat) * BN + col + 64] + red[(1 * 2 + stat) * BN + col + 64]) +
             (red[(2 * 2 + stat) * BN + col + 64] + red[(3 * 2 + stat) * BN + col + 64]);
        p.bn_partial[((static_cast<size_t>(phase) * gridDim.x + blockIdx.x) * 2 + stat) * p.N_pad + col] = t;
      }
    } else {
      p.bn_partial[((static_cast<size_t>(phase) * gridDim.x + blockIdx.x) * 2 + stat) * p.N_pad + n0 + col] = t;
    }
  }
}

template <int BN, bool TMA_A, int TG>
static int launch_conv_gemm_ta(const CUtensorMap& tmap, const CUtensorMap& tmap_a, const ConvGemmParams& p, dim3 grid,
                               cudaStream_t st) {
  using S = ConvTaSmem<BN>;
  static_assert(S::kDynamic <= 227 * 1024, "shared memory per CTA");
  MDGAN_CUDA(configure_smem_once(conv_gemm_ta_kernel<BN, TMA_A, TG>, S::kDynamic));
  MDGAN_LAUNCH((conv_gemm_ta_kernel<BN, TMA_A, TG>), grid, dim3(TMA_A ? kTaThreadsTma : kTaThreads), S::kDynamic, st, tmap,
               tmap_a, p);
  return 0;
}

// MDGAN_CONV_TG = 1 (default) | 2: groups of transposer warps of the TMA-fed kernel (see conv_gemm_ta_kernel).  Measured
// on B200 (round 2, gpurun_out/r2c1_convbench_tg*.log): 1 and 2 groups give the same kernel times to within 1 % on every
// layer shape and bit-identical results -- the K step is paced by the shared-memory port, not by the transposer chain --
// so the default stays at the single group.
static int conv_tg() {
  static const int tg = [] {
    const char* e = getenv("MDGAN_CONV_TG");
    const int v = e ? atoi(e) : 1;
    return v < 1 ? 1 : (v > 2 ? 2 : v);
  }();
  return tg;
}

// MDGAN_CONV_UP2 = 1 (default) | 0: UP-mode layers with 64 output channels pair their column parities in one CTA.
static bool conv_up2_enabled() {
  static const bool on = [] {
    const char* e = getenv("MDGAN_CONV_UP2");
    return e ? e[0] != '0' : true;
  }();
  return on;
}

static int launch_conv_gemm_up2(const CUtensorMap& tmap, const CUtensorMap& tmap_a, const ConvGemmParams& p, dim3 grid,
                                cudaStream_t st) {
  using S = ConvTaSmem<128>;
  if (conv_tg() == 1) {
    MDGAN_CUDA(configure_smem_once(conv_gemm_ta_kernel<128, true, 1, true>, S::kDynamic));
    MDGAN_LAUNCH((conv_gemm_ta_kernel<128, true, 1, true>), grid, dim3(kTaThreadsTma), S::kDynamic, st, tmap, tmap_a, p);
  } else {
    MDGAN_CUDA(configure_smem_once(conv_gemm_ta_kernel<128, true, 2, true>, S::kDynamic));
    MDGAN_LAUNCH((conv_gemm_ta_kernel<128, true, 2, true>), grid, dim3(kTaThreadsTma), S::kDynamic, st, tmap, tmap_a, p);
  }
  return 0;
}

template <int BN>
static int launch_conv_gemm_ta_tma(const CUtensorMap& tmap, const CUtensorMap& tmap_a, const ConvGemmParams& p, dim3 grid,
                                   cudaStream_t st) {
  return conv_tg() == 1 ? launch_conv_gemm_ta<BN, true, 1>(tmap, tmap_a, p, grid, st)
                        : launch_conv_gemm_ta<BN, true, 2>(tmap, tmap_a, p, grid, st);
}

// MDGAN_CONV_TMA_PARTIAL = 1 (default) | 0: also use TMA boxes of whole images that do not fill 128 rows (7x7 grids).
static bool conv_tma_partial_enabled() {
  static const bool on = [] {
    const char* e = getenv("MDGAN_CONV_TMA_PARTIAL");
    return e ? e[0] != '0' : true;
  }();
  return on;
}

// MDGAN_CONV_TMA = 1 (default) | 0: fetch the activation tile with tiled TMA loads where the row tile is a box.
static bool conv_tma_enabled() {
  static const bool on = [] {
    const char* e = getenv("MDGAN_CONV_TMA");
    return e ? e[0] != '0' : true;
  }();
  return on;
}

// MDGAN_CONV_TA = 1 (default) | 0: activation operand through tensor memory (tf32x3 only; same arithmetic, same bits).
static bool conv_ta_enabled() {
  static const bool on = [] {
    const char* e = getenv("MDGAN_CONV_TA");
    return e ? e[0] != '0' : true;
  }();
  return on;
}

template <int BN, int STAGES, bool X3>
static int launch_conv_gemm(const CUtensorMap& tmap, const ConvGemmParams& p, dim3 grid, cudaStream_t st) {
  using S = ConvGemmSmem<BN, STAGES, X3>;
  static_assert(S::kDynamic <= 227 * 1024, "shared memory per CTA");
  MDGAN_CUDA(configure_smem_once(conv_gemm_kernel<BN, STAGES, X3>, S::kDynamic));
  MDGAN_LAUNCH((conv_gemm_kernel<BN, STAGES, X3>), grid, dim3(kThreads), S::kDynamic, st, tmap, p);
  return 0;
}

// Estimated cycles of one launch for tile width bn: waves of CTAs x K steps x MMA time per K step (four K=8 slices,
// three MMAs each in tf32x3), a 128 x bn x 8 MMA costing max(44, bn/2) clk, plus a fixed per-CTA prologue/epilogue.
static double conv_cost(int row_tiles, int phases, int n_pad, int bn, int ksteps, bool x3) {
  const long ctas = static_cast<long>(row_tiles) * phases * (n_pad / bn);
  const long waves = (ctas + 147) / 148;
  const double mma = bn / 2 > 44 ? bn / 2 : 44;
  return waves * (ksteps * 4.0 * (x3 ? 3 : 1) * mma + 3000.0 + 40.0 * bn);
}

}  // namespace mdgan

using namespace mdgan;

// See include/mdgan_b200.h for the contract.
// Row tiling of the GEMM view: rows a CTA owns and, for the TMA-fed kernel, the box (images x grid rows x grid columns).
static int conv_row_tiling(int Hg, int Wg, int precision, int* bn_img, int* bh, int* bw) {
  *bn_img = *bh = *bw = 0;
  if (precision == 1 && conv_ta_enabled() && conv_tma_enabled()) {
    const int rows_img = Hg * Wg;
    if (rows_img <= kBM && (kBM % rows_img == 0 || conv_tma_partial_enabled())) { *bn_img = kBM / rows_img; *bh = Hg; *bw = Wg; }
    else if (Wg <= kBM && kBM % Wg == 0 && Hg % (kBM / Wg) == 0) { *bn_img = 1; *bh = kBM / Wg; *bw = Wg; }
    if (*bn_img > 0) return *bn_img * *bh * *bw;
  }
  return kBM;
}

// true when mdgan_conv_gemm runs this problem with the paired-parity kernel (conv_gemm_ta_kernel<..., UP2>)
static bool conv_uses_up2(int mode, int N_pad, int Hg, int Wg, int precision) {
  if (mode != 1 || N_pad != 64 || precision != 1 || !conv_ta_enabled() || !conv_up2_enabled()) return false;
  int a, b, c;
  conv_row_tiling(Hg, Wg, precision, &a, &b, &c);
  return a > 0;  // TMA-fed activation tile
}

// Number of phase slices the fused statistics of this problem have (bn_partial is [phases][row tiles][2][N_pad]).
extern "C" int mdgan_conv_stat_phases(int mode, int N_pad, int Hg, int Wg, int precision) {
  if (mode != 1) return 1;
  return conv_uses_up2(mode, N_pad, Hg, Wg, precision) ? 2 : 4;
}

// GEMM rows one CTA of mdgan_conv_gemm owns for this row grid (callers size the fused-statistics buffer with it:
// row tiles = ceil(n_img*Hg*Wg / rows)), or 0 when the fused BatchNorm statistics are not available in this mode.
extern "C" int mdgan_conv_rows_per_tile(int Hg, int Wg, int precision) {
  if (precision != 1 || !conv_ta_enabled() || Hg <= 0 || Wg <= 0) return 0;
  int a, b, c;
  return conv_row_tiling(Hg, Wg, precision, &a, &b, &c);
}

extern "C" int mdgan_conv_gemm(const float* src, const float* wpacked, float* dst, const float* bias, int n_img,
                               int Hg, int Wg, int Hs, int Ws, int C, int mode, int N, int N_pad, int out_nchw,
                               int act, int round_tf32, int accumulate, int precision, int force_bn, const float* gate,
                               int gate_act, float gate_slope, float* bn_partial, const float* bnb_z,
                               const float* bnb_stats, int bnb_act, float bnb_slope, int bnb_groups, void* stream) {
  if (!src || !wpacked || !dst) return MDGAN_ERR_BAD_ARG;
  if (bn_partial && (precision != 1 || !conv_ta_enabled() || out_nchw || gate || act != 0 || accumulate))
    return MDGAN_ERR_UNSUPPORTED;
  if (bnb_z && (!bn_partial || !bnb_stats || bias || round_tf32 || N % 16 != 0 || bnb_groups < 1 || bnb_act < 0 || bnb_act > 2))
    return MDGAN_ERR_UNSUPPORTED;
  if (gate && (out_nchw || (gate_act != 1 && gate_act != 2))) return MDGAN_ERR_UNSUPPORTED;
  if (C <= 0 || C % kBK != 0 || mode < 0 || mode > 2) return MDGAN_ERR_UNSUPPORTED;
  if (N_pad % 16 != 0 || N > N_pad || N <= 0) return MDGAN_ERR_UNSUPPORTED;
  ConvGemmParams p{};
  p.src = src; p.dst = dst; p.bias = bias;
  p.n_img = n_img; p.Hg = Hg; p.Wg = Wg; p.Hs = Hs; p.Ws = Ws; p.C = C;
  p.mode = mode; p.N = N; p.N_pad = N_pad; p.out_nchw = out_nchw; p.act = act;
  p.M = n_img * Hg * Wg;
  p.round_tf32 = round_tf32;
  p.accumulate = accumulate;
  p.gate = gate; p.gate_act = gate_act; p.gate_slope = gate_slope;
  p.bn_partial = bn_partial;
  p.bnb_z = bnb_z; p.bnb_stats = bnb_stats; p.bnb_act = bnb_act; p.bnb_slope = bnb_slope;
  p.bnb_rows_per_group = bnb_z ? p.M / bnb_groups : p.M;
  if (bnb_z && (p.M % bnb_groups != 0)) return MDGAN_ERR_UNSUPPORTED;
  if (accumulate && !out_nchw) return MDGAN_ERR_UNSUPPORTED;
  if (p.M <= 0) return MDGAN_ERR_BAD_ARG;
  const int taps = mode == 0 ? 16 : (mode == 1 ? 4 : 1);
  const int phases = mode == 1 ? 4 : 1;
  if (precision != 0 && precision != 1) return MDGAN_ERR_BAD_ARG;
  // TMA-fed activation tile (tf32x3 TMEM-operand kernel): the rows of a CTA must be a box of the (n, i, j) grid --
  // whole images (as many as fit in 128 rows; 7x7 grids give 98-row tiles) or whole grid rows of one image.
  int bn_img = 0, bh = 0, bw = 0;
  p.rows_per_tile = conv_row_tiling(Hg, Wg, precision, &bn_img, &bh, &bw);
  const int row_tiles = ceil_div(p.M, p.rows_per_tile);
  if (bnb_z && p.bnb_rows_per_group % p.rows_per_tile != 0) return MDGAN_ERR_UNSUPPORTED;  // a CTA must not straddle passes
  const bool x3 = precision == 1;
  // Tile width: the candidate (dividing N_pad) with the lowest estimated time -- wide tiles use the tensor core
  // better, narrow ones fill the 148 SMs when the row grid is small.
  int bn = 16;
  double best = 1e300;
  for (int c : {128, 64, 32, 16}) {
    if (N_pad % c != 0) continue;
    const double t = conv_cost(row_tiles, phases, N_pad, c, taps * (C / kBK), x3);
    if (t < best) { best = t; bn = c; }
  }
  if (force_bn > 0) {
    if (N_pad % force_bn != 0) return MDGAN_ERR_BAD_ARG;
    bn = force_bn;
  }
  const uint64_t rows = static_cast<uint64_t>(N_pad) * phases;
  p.lo_row_offset = static_cast<int>(rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (force_bn == 0 && conv_uses_up2(mode, N_pad, Hg, Wg, precision)) {
    CUtensorMap tmap, tmap_a;
    int rc = get_tmap_2d_f32(wpacked, rows * 2, static_cast<uint64_t>(taps) * C, 64, &tmap);
    if (rc != 0) return rc;
    rc = get_tmap_im2col_f32(src, n_img, Hs, Ws, C, bn_img, bh, bw, 1, &tmap_a);
    if (rc != 0) return rc;
    return launch_conv_gemm_up2(tmap, tmap_a, p, dim3(row_tiles, 1, 2), st);
  }
  CUtensorMap tmap;
  int rc = get_tmap_2d_f32(wpacked, rows * (x3 ? 2 : 1), static_cast<uint64_t>(taps) * C, bn, &tmap);
  if (rc != 0) return rc;
  dim3 grid(row_tiles, N_pad / bn, phases);
  if (x3 && conv_ta_enabled()) {
    CUtensorMap tmap_a = tmap;
    bool tma_a = bn_img > 0;
    if (tma_a) {
      const int si = mode == 0 ? 2 : 1;
      rc = get_tmap_im2col_f32(src, n_img, Hs, Ws, C, bn_img, bh, bw, si, &tmap_a);
      if (rc != 0) return rc;  // the tiling above already committed to the box
    }
    if (tma_a) {
      switch (bn) {
        case 128: return launch_conv_gemm_ta_tma<128>(tmap, tmap_a, p, grid, st);
        case 64: return launch_conv_gemm_ta_tma<64>(tmap, tmap_a, p, grid, st);
        case 32: return launch_conv_gemm_ta_tma<32>(tmap, tmap_a, p, grid, st);
        case 16: return launch_conv_gemm_ta_tma<16>(tmap, tmap_a, p, grid, st);
        default: return MDGAN_ERR_UNSUPPORTED;
      }
    }
    switch (bn) {
      case 128: return launch_conv_gemm_ta<128, false, 1>(tmap, tmap_a, p, grid, st);
      case 64: return launch_conv_gemm_ta<64, false, 1>(tmap, tmap_a, p, grid, st);
      case 32: return launch_conv_gemm_ta<32, false, 1>(tmap, tmap_a, p, grid, st);
      case 16: return launch_conv_gemm_ta<16, false, 1>(tmap, tmap_a, p, grid, st);
      default: return MDGAN_ERR_UNSUPPORTED;
    }
  }
  switch (bn) {
    case 128: return x3 ? launch_conv_gemm<128, 3, true>(tmap, p, grid, st) : launch_conv_gemm<128, 6, false>(tmap, p, grid, st);
    case 64: return x3 ? launch_conv_gemm<64, 4, true>(tmap, p, grid, st) : launch_conv_gemm<64, 6, false>(tmap, p, grid, st);
    case 32: return x3 ? launch_conv_gemm<32, 4, true>(tmap, p, grid, st) : launch_conv_gemm<32, 6, false>(tmap, p, grid, st);
    case 16: return x3 ? launch_conv_gemm<16, 4, true>(tmap, p, grid, st) : launch_conv_gemm<16, 6, false>(tmap, p, grid, st);
    default: return MDGAN_ERR_UNSUPPORTED;
  }
}
