// Bandwidth kernels of the MD-GAN step: weight (un)packing, train-mode BatchNorm forward/backward fused with the
// activation, the discriminator head (4x4 valid conv -> sigmoid -> BCE, forward + backward), tanh backward,
// LeakyReLU backward, fused flat Adam.  All activations are NHWC fp32; parameters stay in PyTorch layout.
// Reference semantics: torch.nn.{BatchNorm2d,LeakyReLU,ReLU,Tanh,Sigmoid,BCELoss} and torch.optim.Adam as called
// from /root/reference/src/datasets/{CIFAR10,CelebA}.py and /root/reference/src/actors/{worker,server}.py
// (SURVEY.md Appendix B).
#include "common.cuh"

namespace mdgan {

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ----------------------------------------------------------------------------------------------- weight packing
// mode 0 (DOWN):  out[n][tap][c]      = W[n][c][kh][kw]              W: [N][C][4][4]
// mode 1 (UP):    out[ph][n][t][c]    = W[c][n][kh][kw]              W: [C][N][4][4], kh=(1-ph_h)+2a, kw=(1-ph_w)+2b
// mode 2 (DENSE): out[kk*N + n][c]    = W[c][n][kk]                  W: [C][N][KK]    (convT on a 1x1 input)
// Rows n >= N and columns c >= C are zero padding (N_pad rows per phase / C_pad columns per tap).
// split = 1 (tf32x3): out holds two matrices back to back, hi = tf32(w) and lo = tf32(w - hi).
__global__ void pack_weights_kernel(const float* __restrict__ W, float* __restrict__ out, int mode, int N, int C,
                                    int N_pad, int C_pad, int KK, long long total, int split) {
  pdl_enter();
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float v = 0.f;
  if (mode == 0) {
    const int c = idx % C_pad;
    const int tap = (idx / C_pad) % 16;
    const int n = idx / (16LL * C_pad);
    if (n < N && c < C) v = W[((long long)n * C + c) * 16 + tap];
  } else if (mode == 1) {
    const int c = idx % C_pad;
    const int t = (idx / C_pad) % 4;
    const int n = (idx / (4LL * C_pad)) % N_pad;
    const int ph = idx / (4LL * C_pad * N_pad);
    const int kh = (1 - (ph >> 1)) + 2 * (t >> 1);
    const int kw = (1 - (ph & 1)) + 2 * (t & 1);
    if (n < N && c < C) v = W[((long long)c * N + n) * 16 + kh * 4 + kw];
  } else {
    const int c = idx % C_pad;
    const long long row = idx / C_pad;
    const int n = row % N;
    const int kk = row / N;
    if (c < C) v = W[((long long)c * N + n) * KK + kk];
  }
  const float hi = to_tf32(v);
  out[idx] = hi;
  if (split) out[total + idx] = to_tf32(v - hi);
}

// All weight re-packs of one network in ONE launch (after its Adam step).  jobs: n_jobs records of kPackJobWords
// int64 words in device memory: {W ptr, out ptr, mode, N, C, N_pad, C_pad, KK, total elements, first block, split}.
// mode 0..2 as above; mode 3 is the discriminator head re-order wt[hw*C + c] = w[c*HW + hw] (N = HW, no TF32 rounding:
// the head runs in fp32 on CUDA cores).
// Modes 0 / 1 (the 4x4 stencils, > 95 % of the bytes) are tiled: a block owns one output row n and 32 channels, reads
// the 32 x 16 source taps as 64-byte runs, transposes them through shared memory and writes sixteen 128-byte runs
// (job blocks = N_pad * C_pad / 32); modes 2 / 3 go element-wise (job blocks = ceil(total / 256)).
constexpr int kPackJobWords = 11;
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const long long* __restrict__ jobs, int n_jobs) {
  __shared__ float tile[16][33];
  pdl_enter();
  int j = 0;
  while (j + 1 < n_jobs && (long long)blockIdx.x >= jobs[(j + 1) * kPackJobWords + 9]) ++j;
  const long long* jb = jobs + j * kPackJobWords;
  const float* __restrict__ W = reinterpret_cast<const float*>(jb[0]);
  float* __restrict__ out = reinterpret_cast<float*>(jb[1]);
  const int mode = (int)jb[2], N = (int)jb[3], C = (int)jb[4], N_pad = (int)jb[5], C_pad = (int)jb[6], KK = (int)jb[7];
  const long long total = jb[8];
  const int split = (int)jb[10];
  const long long blk = (long long)blockIdx.x - jb[9];
  if (mode <= 1) {
    const int chunks = C_pad >> 5;
    const int n = blk / chunks, c0 = (int)(blk - (long long)n * chunks) << 5;
    // source: mode 0 W[n][c][16], mode 1 W[c][n][16]; element e = cl * 16 + k of the 32 x 16 tile
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = threadIdx.x + 256 * r;
      const int cl = e >> 4, k = e & 15, c = c0 + cl;
      float v = 0.f;
      if (n < N && c < C) v = mode == 0 ? W[((long long)n * C + c) * 16 + k] : W[((long long)c * N + n) * 16 + k];
      tile[k][cl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = threadIdx.x + 256 * r;
      const int row = e >> 5, cl = e & 31;  // row: mode 0 the tap; mode 1 (ph, t) = (row >> 2, row & 3)
      float v;
      long long idx;
      if (mode == 0) {
        v = tile[row][cl];
        idx = ((long long)n * 16 + row) * C_pad + c0 + cl;
      } else {
        const int ph = row >> 2, t = row & 3;
        const int kh = (1 - (ph >> 1)) + 2 * (t >> 1), kw = (1 - (ph & 1)) + 2 * (t & 1);
        v = tile[kh * 4 + kw][cl];
        idx = (((long long)ph * N_pad + n) * 4 + t) * C_pad + c0 + cl;
      }
      const float hi = to_tf32(v);
      out[idx] = hi;
      if (split) out[total + idx] = to_tf32(v - hi);
    }
    return;
  }
  const long long idx = blk * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  if (mode == 2) {
    const int c = idx % C_pad;
    const long long row = idx / C_pad;
    const int n = row % N;
    const int kk = row / N;
    const float v = c < C ? W[((long long)c * N + n) * KK + kk] : 0.f;
    const float hi = to_tf32(v);
    out[idx] = hi;
    if (split) out[total + idx] = to_tf32(v - hi);
  } else {
    const int hw = idx / C, c = idx - (long long)hw * C;
    out[idx] = W[(long long)c * N + hw];
  }
}

// Reduce split-K partial slices and scatter to the PyTorch parameter layout.
// mode 0: grad[c1][c2][tap] = sum_s partial[s][tap][c1][c2]      (C1p x C2 slices, c1 < C1)
// mode 2: grad[c1][n][kk]   = sum_s partial[s][0][c1][kk*N + n]  (dense first generator layer; C2 = KK*N)
__global__ void wgrad_unpack_kernel(const float* __restrict__ partial, float* __restrict__ grad, int mode, int splits,
                                    int taps, int C1, int C1p, int C2, int N, int KK, long long total) {
  pdl_enter();
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long src;
  int tap;
  if (mode == 0) {
    tap = idx % 16;
    const int c2 = (idx / 16) % C2;
    const int c1 = idx / (16LL * C2);
    src = ((long long)tap * C1p + c1) * C2 + c2;
  } else {
    const int kk = idx % KK;
    const int n = (idx / KK) % N;
    const int c1 = idx / ((long long)KK * N);
    src = (long long)c1 * C2 + (long long)kk * N + n;
  }
  const long long slice = (long long)taps * C1p * C2;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[s * slice + src];
  grad[idx] = acc;
}

// Tiled form of mode 0 (C2 % 32 == 0): a block owns one c1 and 32 consecutive c2; it reads the sixteen tap rows of every
// split as 128-byte runs (summing the splits in order), transposes through shared memory and writes the 512 gradient
// words grad[c1][c2_0 .. c2_0+32)[16 taps], which are contiguous.  The element-wise kernel above reads with a stride of
// one tap slice between neighbouring threads (eight-fold sector amplification).
__global__ void __launch_bounds__(256) wgrad_unpack_tiled_kernel(const float* __restrict__ partial, float* __restrict__ grad,
                                                                 int splits, int C1, int C1p, int C2) {
  __shared__ float tile[16][33];
  pdl_enter();
  const int chunks = C2 >> 5;
  const int c1 = blockIdx.x / chunks, c2_0 = (blockIdx.x - c1 * chunks) << 5;
  const long long slice = 16LL * C1p * C2;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int tap = (threadIdx.x >> 5) + 8 * r, cl = threadIdx.x & 31;
    const float* src = partial + ((long long)tap * C1p + c1) * C2 + c2_0 + cl;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += src[s * slice];
    tile[tap][cl] = acc;
  }
  __syncthreads();
  float* dst = grad + ((long long)c1 * C2 + c2_0) * 16;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int e = threadIdx.x + 256 * r;
    dst[e] = tile[e & 15][e >> 4];
  }
}

// out[i] = sum_s partial[s][i].  Block = 32 outputs x 8 slice lanes (fixed summation order: lane-strided partial sums,
// then lanes 0..7), so a few hundred slices of a small tensor are read with 8-way parallelism per output.
__global__ void __launch_bounds__(256) reduce_slices_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int slices, long long n) {
  __shared__ float red[8][32];
  pdl_enter();
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const long long idx = blockIdx.x * 32LL + o;
  float acc = 0.f;
  if (idx < n) {
#pragma unroll 4
    for (int s = sl; s < slices; s += 8) acc += partial[s * n + idx];
  }
  red[sl][o] = acc;
  __syncthreads();
  if (sl == 0 && idx < n) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += red[l][o];
    out[idx] = t;
  }
}

// ----------------------------------------------------------------------------------------------- BatchNorm (train)
// x is [G*Pg, C] (G independent passes of Pg rows each, e.g. real || X_d).  Statistics kernel: grid = (G * chunks,
// C / 32); a block reduces rows_per_chunk rows of one 32-channel slab (256 threads = 32 row lanes x 8 float4 quads)
// into partial[chunk][2][C]; the LAST block of a slab to finish (device counter per slab) finalizes its 32 channels:
// per group (in order) mean / biased var -> scale, shift; running stats with momentum and unbiased variance;
// num_batches_tracked += G.  stats layout: [G][4][C] = mean, invstd, scale, shift.
constexpr int kBnSlab = 32;

__device__ __forceinline__ void bn_block_reduce(float (&s)[4], float (&ss)[4], float* sm, float* __restrict__ partial_row,
                                                int slab, int C) {
  // sm: [32 row lanes][2][32]
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sm[(rl * 2 + 0) * kBnSlab + q * 4 + j] = s[j];
    sm[(rl * 2 + 1) * kBnSlab + q * 4 + j] = ss[j];
  }
  __syncthreads();
  if (threadIdx.x < 2 * kBnSlab) {
    const int stat = threadIdx.x >> 5, ch = threadIdx.x & 31;
    float acc = 0.f;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) acc += sm[(l * 2 + stat) * kBnSlab + ch];
    partial_row[stat * C + slab * kBnSlab + ch] = acc;
  }
}

// true in every thread of the last block of this slab to arrive (all partials of the slab are then visible)
__device__ __forceinline__ bool bn_last_block(unsigned int* counters, int slab) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(counters + slab, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x == 0) counters[slab] = 0;
  }
  return last;
}

// sum over the chunks of group g of partial[.][stat 0/1][c]: 8 lanes (lane & 7) per channel, fp64, butterfly
__device__ __forceinline__ void bn_chunk_sums(const float* __restrict__ partial, int g, int chunks_per_group, int C,
                                              int c, double& s, double& ss) {
  const int cl = threadIdx.x & 7;
  s = 0.0;
  ss = 0.0;
  const float* pg = partial + (long long)g * chunks_per_group * 2 * C + c;
#pragma unroll 4
  for (int ch = cl; ch < chunks_per_group; ch += 8) {
    s += __ldcg(pg + (long long)ch * 2 * C);
    ss += __ldcg(pg + (long long)ch * 2 * C + C);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
}

__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, float* __restrict__ partial, unsigned int* __restrict__ counters,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                float* __restrict__ running_var, long long* __restrict__ nbt, float* __restrict__ stats, int G, int Pg,
                int C, int chunks_per_group, int rows_per_chunk, float eps, float momentum) {
  __shared__ float sm[32 * 2 * kBnSlab];
  pdl_enter();
  const int slab = blockIdx.y;
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int g = blockIdx.x / chunks_per_group, ch = blockIdx.x % chunks_per_group;
  const int r0 = ch * rows_per_chunk;
  const int r1 = min(Pg, r0 + rows_per_chunk);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  const float* base = x + ((long long)g * Pg) * C + slab * kBnSlab + q * 4;
#pragma unroll 4
  for (int r = r0 + rl; r < r1; r += 32) {
    const float4 v = *reinterpret_cast<const float4*>(base + (long long)r * C);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
  }
  bn_block_reduce(s, ss, sm, partial + (long long)blockIdx.x * 2 * C, slab, C);
  if (!bn_last_block(counters, slab)) return;
  if (slab == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += G;
  const int c = slab * kBnSlab + (threadIdx.x >> 3);  // 8 lanes per channel
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  const float gm = gamma[c], bt = beta[c];
  for (int gi = 0; gi < G; ++gi) {
    double sum, sq;
    bn_chunk_sums(partial, gi, chunks_per_group, C, c, sum, sq);
    const double mean = sum / Pg;
    double var = sq / Pg - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gm * invstd;
    const float unbiased = (float)(var * ((double)Pg / (double)(Pg > 1 ? Pg - 1 : 1)));
    rm = (1.f - momentum) * rm + momentum * (float)mean;
    rv = (1.f - momentum) * rv + momentum * unbiased;
    if ((threadIdx.x & 7) == 0) {
      float* st = stats + (long long)gi * 4 * C;
      st[c] = (float)mean;
      st[C + c] = invstd;
      st[2 * C + c] = sc;
      st[3 * C + c] = bt - (float)mean * sc;
    }
  }
  if ((threadIdx.x & 7) == 0) {
    if (running_mean) running_mean[c] = rm;
    if (running_var) running_var[c] = rv;
  }
}

// Finalize kernel for statistics reduced by the producing GEMM (mdgan_conv_gemm, bn_partial): block = 8 channels x 32
// lanes; a channel's lanes walk its partial slices (phase, tile, fold) in a fixed strided order, accumulate in fp64 and
// are combined by a butterfly; then exactly the arithmetic of bn_stats_kernel's last block.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ partial, int phases, int row_tiles, int tiles_per_group, int col_stride,
                   int fold, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt,
                   float* __restrict__ stats, int G, int Pg, int C, float eps, float momentum) {
  pdl_enter();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;  // whole warps: C % 8 == 0 is required by the launcher
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += G;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  const float gm = gamma[c], bt = beta[c];
  const int per_group = phases * tiles_per_group * fold;
  for (int g = 0; g < G; ++g) {
    double sum = 0.0, sq = 0.0;
    auto slice = [&](int i) {
      const int f = i % fold, r = i / fold;
      const int t = r % tiles_per_group, ph = r / tiles_per_group;
      return partial + (static_cast<long long>(ph) * row_tiles + g * tiles_per_group + t) * 2 * col_stride + f * C + c;
    };
    int i = lane;
    for (; i + 96 < per_group; i += 128) {   // four slices per lane in flight (the loads are independent, the adds ordered)
      const float* s0 = slice(i); const float* s1 = slice(i + 32); const float* s2 = slice(i + 64); const float* s3 = slice(i + 96);
      const float a0 = __ldcg(s0), b0 = __ldcg(s0 + col_stride), a1 = __ldcg(s1), b1 = __ldcg(s1 + col_stride);
      const float a2 = __ldcg(s2), b2 = __ldcg(s2 + col_stride), a3 = __ldcg(s3), b3 = __ldcg(s3 + col_stride);
      sum += a0; sum += a1; sum += a2; sum += a3;
      sq += b0; sq += b1; sq += b2; sq += b3;
    }
    for (; i < per_group; i += 32) {
      const float* sl = slice(i);
      sum += __ldcg(sl);
      sq += __ldcg(sl + col_stride);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const double mean = sum / Pg;
    double var = sq / Pg - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gm * invstd;
    const float unbiased = (float)(var * ((double)Pg / (double)(Pg > 1 ? Pg - 1 : 1)));
    rm = (1.f - momentum) * rm + momentum * (float)mean;
    rv = (1.f - momentum) * rv + momentum * unbiased;
    if (lane == 0) {
      float* st = stats + (long long)g * 4 * C;
      st[c] = (float)mean;
      st[C + c] = invstd;
      st[2 * C + c] = sc;
      st[3 * C + c] = bt - (float)mean * sc;
    }
  }
  if (lane == 0) {
    if (running_mean) running_mean[c] = rm;
    if (running_var) running_var[c] = rv;
  }
}

// act: 0 none, 1 ReLU, 2 LeakyReLU(slope)
__device__ __forceinline__ float act_fwd(float y, int act, float slope) {
  if (act == 1) return y > 0.f ? y : 0.f;
  if (act == 2) return y > 0.f ? y : y * slope;
  return y;
}
__device__ __forceinline__ float act_grad(float y, int act, float slope) {
  if (act == 1) return y > 0.f ? 1.f : 0.f;
  if (act == 2) return y > 0.f ? 1.f : slope;
  return 1.f;
}

// Stage 3: out = act(x*scale + shift).  One float4 per thread and step, grid-stride; the (row, channel, pass) of an
// element comes from 32-bit divisions (the tensors of this path hold < 2^31 elements: checked by the launcher) -- the
// 64-bit div / mod of the first version cost more instructions than the memory traffic they indexed.
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, float* __restrict__ out, int Pg, int C,
                long long total4, int act, float slope, int round_tf32) {
  pdl_enter();
  const unsigned C4 = (unsigned)C >> 2, n4 = (unsigned)total4, stride = gridDim.x * blockDim.x;
  for (unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const unsigned row = i4 / C4, c = (i4 - row * C4) << 2, g = row / (unsigned)Pg;
    const float* st = stats + (size_t)g * 4 * C;
    const float4 v = reinterpret_cast<const float4*>(x)[i4];
    const float4 sc = __ldg(reinterpret_cast<const float4*>(st + 2 * C + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(st + 3 * C + c));
    float4 o;
    o.x = act_fwd(fmaf(v.x, sc.x, sh.x), act, slope);
    o.y = act_fwd(fmaf(v.y, sc.y, sh.y), act, slope);
    o.z = act_fwd(fmaf(v.z, sc.z, sh.z), act, slope);
    o.w = act_fwd(fmaf(v.w, sc.w, sh.w), act, slope);
    if (round_tf32) { o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w); }
    reinterpret_cast<float4*>(out)[i4] = o;
  }
}

// Backward statistics (same grid / last-block scheme as bn_stats_kernel): dy = da * act'(y); per channel
// sums[g][2][C] = (sum dy, sum dy*xhat); dgamma / dbeta = totals over all groups (optional).
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ da, const float* __restrict__ x, const float* __restrict__ stats,
                    float* __restrict__ partial, unsigned int* __restrict__ counters, float* __restrict__ sums,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int Pg, int C, int chunks_per_group,
                    int rows_per_chunk, int act, float slope) {
  __shared__ float sm[32 * 2 * kBnSlab];
  pdl_enter();
  const int slab = blockIdx.y;
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int g = blockIdx.x / chunks_per_group, ch = blockIdx.x % chunks_per_group;
  const int r0 = ch * rows_per_chunk;
  const int r1 = min(Pg, r0 + rows_per_chunk);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  {
    const float* st = stats + (long long)g * 4 * C + slab * kBnSlab + q * 4;
    float mean[4], invstd[4], sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mean[j] = st[j]; invstd[j] = st[C + j]; sc[j] = st[2 * C + j]; sh[j] = st[3 * C + j]; }
    const long long base = ((long long)g * Pg) * C + slab * kBnSlab + q * 4;
#pragma unroll 2
    for (int r = r0 + rl; r < r1; r += 32) {
      const float4 xv = *reinterpret_cast<const float4*>(x + base + (long long)r * C);
      const float4 dv = *reinterpret_cast<const float4*>(da + base + (long long)r * C);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = fmaf(xs[j], sc[j], sh[j]);
        const float dy = ds[j] * act_grad(y, act, slope);
        s[j] += dy;
        ss[j] += dy * ((xs[j] - mean[j]) * invstd[j]);
      }
    }
  }
  bn_block_reduce(s, ss, sm, partial + (long long)blockIdx.x * 2 * C, slab, C);
  if (!bn_last_block(counters, slab)) return;
  const int c = slab * kBnSlab + (threadIdx.x >> 3);
  double tg = 0.0, tb = 0.0;
  for (int gi = 0; gi < G; ++gi) {
    double sum, sq;
    bn_chunk_sums(partial, gi, chunks_per_group, C, c, sum, sq);
    if ((threadIdx.x & 7) == 0) {
      sums[(long long)gi * 2 * C + c] = (float)sum;
      sums[(long long)gi * 2 * C + C + c] = (float)sq;
    }
    tb += sum;
    tg += sq;
  }
  if ((threadIdx.x & 7) == 0) {
    if (dgamma) dgamma[c] = (float)tg;
    if (dbeta) dbeta[c] = (float)tb;
  }
}

// Backward stage 3: dx = scale * (dy - mean(dy) - xhat * mean(dy*xhat))   (32-bit index math, grid-stride: see bn_apply)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ x, const float* __restrict__ stats,
                    const float* __restrict__ sums, float* __restrict__ dx, int Pg, int C, long long total4, int act,
                    float slope, int round_tf32) {
  pdl_enter();
  const unsigned C4 = (unsigned)C >> 2, n4 = (unsigned)total4, stride = gridDim.x * blockDim.x;
  const float invP = 1.f / (float)Pg;
  for (unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const unsigned row = i4 / C4, c = (i4 - row * C4) << 2, g = row / (unsigned)Pg;
    const float* st = stats + (size_t)g * 4 * C + c;
    const float* sm = sums + (size_t)g * 2 * C + c;
    const float4 xv = reinterpret_cast<const float4*>(x)[i4];
    const float4 dv = reinterpret_cast<const float4*>(da)[i4];
    const float4 mean4 = __ldg(reinterpret_cast<const float4*>(st)), istd4 = __ldg(reinterpret_cast<const float4*>(st + C));
    const float4 sc4 = __ldg(reinterpret_cast<const float4*>(st + 2 * C)), sh4 = __ldg(reinterpret_cast<const float4*>(st + 3 * C));
    const float4 s14 = __ldg(reinterpret_cast<const float4*>(sm)), s24 = __ldg(reinterpret_cast<const float4*>(sm + C));
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    const float mean[4] = {mean4.x, mean4.y, mean4.z, mean4.w}, istd[4] = {istd4.x, istd4.y, istd4.z, istd4.w};
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
    const float s1[4] = {s14.x, s14.y, s14.z, s14.w}, s2[4] = {s24.x, s24.y, s24.z, s24.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y = fmaf(xs[j], sc[j], sh[j]);
      const float dy = ds[j] * act_grad(y, act, slope);
      const float xhat = (xs[j] - mean[j]) * istd[j];
      o[j] = sc[j] * (dy - s1[j] * invP - xhat * (s2[j] * invP));
      if (round_tf32) o[j] = to_tf32(o[j]);
    }
    reinterpret_cast<float4*>(dx)[i4] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Backward twin of bn_finalize_kernel: partial slices hold the column sums of dy and of dy * xhat reduced by the
// data-gradient GEMM that produced dy (mdgan_conv_gemm, bnb_*).  sums[g][2][C] per pass, dgamma / dbeta over all passes.
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int phases, int row_tiles, int tiles_per_group, int col_stride,
                       float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int C) {
  pdl_enter();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int per_group = phases * tiles_per_group;
  double tg = 0.0, tb = 0.0;
  for (int g = 0; g < G; ++g) {
    double sum = 0.0, sq = 0.0;
    auto slice = [&](int i) {
      const int t = i % tiles_per_group, ph = i / tiles_per_group;
      return partial + (static_cast<long long>(ph) * row_tiles + g * tiles_per_group + t) * 2 * col_stride + c;
    };
    int i = lane;
    for (; i + 96 < per_group; i += 128) {   // four slices per lane in flight
      const float* s0 = slice(i); const float* s1 = slice(i + 32); const float* s2 = slice(i + 64); const float* s3 = slice(i + 96);
      const float a0 = __ldcg(s0), b0 = __ldcg(s0 + col_stride), a1 = __ldcg(s1), b1 = __ldcg(s1 + col_stride);
      const float a2 = __ldcg(s2), b2 = __ldcg(s2 + col_stride), a3 = __ldcg(s3), b3 = __ldcg(s3 + col_stride);
      sum += a0; sum += a1; sum += a2; sum += a3;
      sq += b0; sq += b1; sq += b2; sq += b3;
    }
    for (; i < per_group; i += 32) {
      const float* sl = slice(i);
      sum += __ldcg(sl);
      sq += __ldcg(sl + col_stride);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
      sums[(long long)g * 2 * C + c] = (float)sum;
      sums[(long long)g * 2 * C + C + c] = (float)sq;
    }
    tb += sum;
    tg += sq;
  }
  if (lane == 0) {
    if (dgamma) dgamma[c] = (float)tg;
    if (dbeta) dbeta[c] = (float)tb;
  }
}

// dx = scale * (dy - mean(dy) - xhat * mean(dy * xhat)) with dy already gated by the activation (see above)
__global__ void __launch_bounds__(256)
bn_bwd_apply_dy_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                       const float* __restrict__ sums, float* __restrict__ dx, int Pg, int C, long long total4,
                       int round_tf32) {
  pdl_enter();
  const unsigned C4 = (unsigned)C >> 2, n4 = (unsigned)total4, stride = gridDim.x * blockDim.x;
  const float invP = 1.f / (float)Pg;
  for (unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const unsigned row = i4 / C4, c = (i4 - row * C4) << 2, g = row / (unsigned)Pg;
    const float* st = stats + (size_t)g * 4 * C + c;
    const float* sm = sums + (size_t)g * 2 * C + c;
    const float4 xv = reinterpret_cast<const float4*>(x)[i4];
    const float4 dv = reinterpret_cast<const float4*>(dy)[i4];
    const float4 mean4 = __ldg(reinterpret_cast<const float4*>(st)), istd4 = __ldg(reinterpret_cast<const float4*>(st + C));
    const float4 sc4 = __ldg(reinterpret_cast<const float4*>(st + 2 * C));
    const float4 s14 = __ldg(reinterpret_cast<const float4*>(sm)), s24 = __ldg(reinterpret_cast<const float4*>(sm + C));
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    const float mean[4] = {mean4.x, mean4.y, mean4.z, mean4.w}, istd[4] = {istd4.x, istd4.y, istd4.z, istd4.w};
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
    const float s1[4] = {s14.x, s14.y, s14.z, s14.w}, s2[4] = {s24.x, s24.y, s24.z, s24.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xhat = (xs[j] - mean[j]) * istd[j];
      o[j] = sc[j] * (ds[j] - s1[j] * invP - xhat * (s2[j] * invP));
      if (round_tf32) o[j] = to_tf32(o[j]);
    }
    reinterpret_cast<float4*>(dx)[i4] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// dz = da * act'(a) for an activation with no BatchNorm in front (a = act(z); sign(a) == sign(z) for slope > 0).
__global__ void act_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a, float* __restrict__ dz,
                               long long total4, int act, float slope, int round_tf32) {
  pdl_enter();
  long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const float4 av = reinterpret_cast<const float4*>(a)[i4];
  const float4 dv = reinterpret_cast<const float4*>(da)[i4];
  float4 o;
  o.x = dv.x * act_grad(av.x, act, slope);
  o.y = dv.y * act_grad(av.y, act, slope);
  o.z = dv.z * act_grad(av.z, act, slope);
  o.w = dv.w * act_grad(av.w, act, slope);
  if (round_tf32) { o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w); }
  reinterpret_cast<float4*>(dz)[i4] = o;
}

// d(pre-tanh) = s * (1 - x^2) * scale   (generator output layer; s = group-summed feedback, x = tanh output)
__global__ void tanh_bwd_kernel(const float* __restrict__ s, const float* __restrict__ x, float* __restrict__ out,
                                long long n, float scale) {
  pdl_enter();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i];
  out[i] = s[i] * (1.f - xv * xv) * scale;
}

// ----------------------------------------------------------------------------------------------- discriminator head
// wt[hw*C + c] = w[c*HW + hw]: the head weight re-ordered once per optimiser step to the NHWC order of the activations
// so that the per-sample dot products below read both operands with coalesced float4 loads.
__global__ void head_pack_kernel(const float* __restrict__ w, float* __restrict__ wt, int HW, int C) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW * C) return;
  const int hw = i / C, c = i - hw * C;
  wt[i] = w[c * HW + hw];
}

// logits[n] = <a[n, :], wt>,  a is NHWC [n, HW, C], wt from head_pack_kernel.
// p = sigmoid(logit); per-sample BCE term with the log clamp at -100; dlogit = dBCE/dlogit * (1/b).
// Labels: samples of group g (n / b) use label[g].  One 256-thread block per sample.
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ a, const float* __restrict__ wt, const float* __restrict__ label,
                float* __restrict__ prob, float* __restrict__ loss_terms, float* __restrict__ dlogit,
                float* __restrict__ loss, unsigned int* __restrict__ counter, int n_total, int b, int G, int L) {
  __shared__ float red[8];
  __shared__ bool last;
  pdl_enter();
  const int n = blockIdx.x;
  const float4* row = reinterpret_cast<const float4*>(a + (long long)n * L);
  const float4* w4 = reinterpret_cast<const float4*>(wt);
  float acc = 0.f;
  for (int i = threadIdx.x; i < (L >> 2); i += 256) {
    const float4 v = row[i];
    const float4 u = __ldg(w4 + i);
    acc = fmaf(v.x, u.x, acc);
    acc = fmaf(v.y, u.y, acc);
    acc = fmaf(v.z, u.z, acc);
    acc = fmaf(v.w, u.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float logit = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) logit += red[i];
    const float y = label[n / b];
    const float p = 1.f / (1.f + expf(-logit));
    const float lp = fmaxf(logf(p), -100.f);
    const float l1p = fmaxf(log1pf(-p), -100.f);
    prob[n] = p;
    loss_terms[n] = (y - 1.f) * l1p - y * lp;
    const float pq = (1.f - p) * p;
    // BCELoss backward: (p - y) / max(p(1-p), 1e-12) / b; sigmoid backward: * p(1-p)
    dlogit[n] = ((p - y) / fmaxf(pq, 1e-12f)) * (1.f / (float)b) * pq;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // The last block to finish reduces the per-sample terms in a fixed order: loss[g] = mean over the b samples of
  // group g; loss[G] = sum over groups (the reference's d_loss).
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

// da[n, l] = dlogit[n] * wt[l];  dw[c*HW + hw] = sum_n dlogit[n] * a[n, l]  (dw optional, PyTorch layout)
__global__ void head_bwd_kernel(const float* __restrict__ a, const float* __restrict__ wt,
                                const float* __restrict__ dlogit, float* __restrict__ da, float* __restrict__ dw,
                                int n_total, int HW, int C) {
  pdl_enter();
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = HW * C;
  if (l >= L) return;
  const int hw = l / C, c = l - hw * C;
  const float wv = wt[l];
  float acc = 0.f;
  for (int n = 0; n < n_total; ++n) {
    const float d = dlogit[n];
    da[(long long)n * L + l] = d * wv;
    if (dw) acc = fmaf(d, a[(long long)n * L + l], acc);
  }
  if (dw) dw[c * HW + hw] = acc;
}

// ----------------------------------------------------------------------------------------------- Adam (torch.optim.Adam)
// step_count[0] lives on the device so the launch is CUDA-graph friendly; step_count[1] is a block counter: the last
// block to finish (every block has read the step by then) writes the incremented step and clears the counter.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, int* __restrict__ step_count, float lr, float beta1,
                            float beta2, float eps, int vec4) {
  __shared__ float s_step_size, s_bc2_sqrt;
  __shared__ int s_t;
  pdl_enter();
  if (threadIdx.x == 0) {
    const int t = *reinterpret_cast<volatile int*>(step_count) + 1;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
    s_t = t;
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  auto update = [&](float gi, float& pi, float& mi, float& vi) {
    mi = mi + (gi - mi) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * beta2 + (1.f - beta2) * gi * gi;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  const long long n4 = vec4 ? (n >> 2) : 0;  // float4 body (all four buffers 16-byte aligned), scalar tail
  for (long long i = tid; i < n4; i += nthr) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    update(g4.x, p4.x, m4.x, v4.x);
    update(g4.y, p4.y, m4.y, v4.y);
    update(g4.z, p4.z, m4.z, v4.z);
    update(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (long long i = n4 * 4 + tid; i < n; i += nthr) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(g[i], pi, mi, vi);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* cnt = reinterpret_cast<unsigned int*>(step_count + 1);
    if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
      *cnt = 0;
      step_count[0] = s_t;
    }
  }
}

// ----------------------------------------------------------------------------------------------- misc
// out[r][c] (cols_out >= cols_in, zero padded), optionally rounded to TF32: pads z [kb, 100] to [kb, 128].
__global__ void pad_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols_in,
                                int cols_out, int round_tf32) {
  pdl_enter();
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols_out) return;
  const int c = idx % cols_out;
  const long long r = idx / cols_out;
  float v = c < cols_in ? in[r * cols_in + c] : 0.f;
  out[idx] = round_tf32 ? to_tf32(v) : v;
}

// out[i] = sum_k in_k[i] over `count` equally sized slices spaced `stride` floats apart (feedback group sum).
__global__ void sum_slices_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int count,
                                  long long stride) {
  pdl_enter();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int k = 0; k < count; ++k) acc += in[k * stride + i];
  out[i] = acc;
}

}  // namespace mdgan

using namespace mdgan;

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }
// BatchNorm apply kernels: grid-stride, at most 8 blocks of 256 threads per SM (two to four float4 per thread at b = 64)
static inline unsigned bn_apply_blocks(long long total4) {
  const unsigned b = blocks_for(total4, 256);
  return b < 148u * 8u ? b : 148u * 8u;
}

extern "C" int mdgan_pack_weights(const float* W, float* out, int mode, int N, int C, int N_pad, int C_pad, int KK,
                                  int split, void* stream) {
  if (!W || !out || mode < 0 || mode > 2) return MDGAN_ERR_BAD_ARG;
  long long total;
  if (mode == 0) total = (long long)N_pad * 16 * C_pad;
  else if (mode == 1) total = 4LL * N_pad * 4 * C_pad;
  else total = (long long)KK * N * C_pad;
  MDGAN_LAUNCH(pack_weights_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, W, out, mode, N, C,
               N_pad, C_pad, KK, total, split);
  return 0;
}

extern "C" int mdgan_pack_job_words(void) { return kPackJobWords; }

extern "C" int mdgan_pack_weights_multi(const long long* jobs_dev, int n_jobs, int total_blocks, void* stream) {
  if (!jobs_dev || n_jobs <= 0 || total_blocks <= 0) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(pack_weights_multi_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream, jobs_dev, n_jobs);
  return 0;
}

extern "C" int mdgan_wgrad_unpack(const float* partial, float* grad, int mode, int splits, int C1, int C1p, int C2,
                                  int N, int KK, void* stream) {
  if (!partial || !grad || (mode != 0 && mode != 2)) return MDGAN_ERR_BAD_ARG;
  const int taps = mode == 0 ? 16 : 1;
  const long long total = mode == 0 ? (long long)C1 * C2 * 16 : (long long)C1 * N * KK;
  if (mode == 0 && C2 % 32 == 0) {
    MDGAN_LAUNCH(wgrad_unpack_tiled_kernel, dim3((unsigned)(C1 * (C2 / 32))), dim3(256), 0, (cudaStream_t)stream, partial,
                 grad, splits, C1, C1p, C2);
    return 0;
  }
  MDGAN_LAUNCH(wgrad_unpack_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, partial, grad, mode,
               splits, taps, C1, C1p, C2, N, KK, total);
  return 0;
}

extern "C" int mdgan_reduce_slices(const float* partial, float* out, int slices, long long n, void* stream) {
  if (!partial || !out || slices <= 0) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(reduce_slices_kernel, dim3(blocks_for(n, 32)), dim3(256), 0, (cudaStream_t)stream, partial, out, slices, n);
  return 0;
}

// Chunking shared by the BN forward and backward reductions: about four blocks per SM over (chunks x 32-channel
// slabs), at least 32 rows per chunk.
static void bn_chunks(int Pg, int G, int C, int* chunks_per_group, int* rows_per_chunk) {
  const int slabs = C / 32 > 0 ? C / 32 : 1;
  int cpg = 592 / (G * slabs);
  if (cpg > (Pg + 31) / 32) cpg = (Pg + 31) / 32;
  if (cpg < 1) cpg = 1;
  *rows_per_chunk = (Pg + cpg - 1) / cpg;
  *chunks_per_group = (Pg + *rows_per_chunk - 1) / *rows_per_chunk;
}

extern "C" long long mdgan_bn_workspace_floats(int G, int Pg, int C) {
  int cpg, rpc;
  bn_chunks(Pg, G, C, &cpg, &rpc);
  return (long long)G * cpg * 2 * C;
}

extern "C" int mdgan_bn_forward(const float* x, float* out, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, long long* num_batches_tracked, float* stats, float* workspace,
                                unsigned int* counters, int G, int Pg, int C, float eps, float momentum, int act,
                                float slope, int round_tf32, void* stream) {
  if (!x || !out || !gamma || !beta || !stats || !workspace || !counters) return MDGAN_ERR_BAD_ARG;
  if (C % 32 != 0 || C > 1024 || G < 1 || Pg < 1) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int cpg, rpc;
  bn_chunks(Pg, G, C, &cpg, &rpc);
  MDGAN_LAUNCH(bn_stats_kernel, dim3(G * cpg, C / 32), dim3(256), 0, st, x, workspace, counters, gamma, beta, running_mean,
               running_var, num_batches_tracked, stats, G, Pg, C, cpg, rpc, eps, momentum);
  const long long total4 = (long long)G * Pg * C / 4;
  if (total4 >= (1LL << 29)) return MDGAN_ERR_UNSUPPORTED;   // 32-bit element indices in the apply kernels
  MDGAN_LAUNCH(bn_apply_kernel, dim3(bn_apply_blocks(total4)), dim3(256), 0, st, x, stats, out, Pg, C, total4, act, slope,
               round_tf32);
  return 0;
}

extern "C" int mdgan_bn_finalize(const float* partial, int phases, int row_tiles, int tiles_per_group, int col_stride,
                                 int fold, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, long long* num_batches_tracked, float* stats, int G, int Pg, int C,
                                 float eps, float momentum, void* stream) {
  if (!partial || !gamma || !beta || !stats) return MDGAN_ERR_BAD_ARG;
  if (C % 8 != 0 || G < 1 || Pg < 1 || phases < 1 || fold < 1 || tiles_per_group < 1 || G * tiles_per_group > row_tiles ||
      fold * C > col_stride)
    return MDGAN_ERR_UNSUPPORTED;
  MDGAN_LAUNCH(bn_finalize_kernel, dim3(C / 8), dim3(256), 0, (cudaStream_t)stream, partial, phases, row_tiles,
               tiles_per_group, col_stride, fold, gamma, beta, running_mean, running_var, num_batches_tracked, stats, G, Pg,
               C, eps, momentum);
  return 0;
}

extern "C" int mdgan_bn_apply(const float* x, const float* stats, float* out, int G, int Pg, int C, int act, float slope,
                              int round_tf32, void* stream) {
  if (!x || !stats || !out) return MDGAN_ERR_BAD_ARG;
  if (C % 4 != 0 || G < 1 || Pg < 1) return MDGAN_ERR_UNSUPPORTED;
  const long long total4 = (long long)G * Pg * C / 4;
  if (total4 >= (1LL << 29)) return MDGAN_ERR_UNSUPPORTED;
  MDGAN_LAUNCH(bn_apply_kernel, dim3(bn_apply_blocks(total4)), dim3(256), 0, (cudaStream_t)stream, x, stats, out, Pg, C,
               total4, act, slope, round_tf32);
  return 0;
}

extern "C" int mdgan_bn_backward(const float* da, const float* x, const float* stats, float* dx, float* dgamma,
                                 float* dbeta, float* sums, float* workspace, unsigned int* counters, int G, int Pg, int C,
                                 int act, float slope, int round_tf32, void* stream) {
  if (!da || !x || !stats || !dx || !sums || !workspace || !counters) return MDGAN_ERR_BAD_ARG;
  if (C % 32 != 0 || C > 1024 || G < 1 || Pg < 1) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int cpg, rpc;
  bn_chunks(Pg, G, C, &cpg, &rpc);
  MDGAN_LAUNCH(bn_bwd_stats_kernel, dim3(G * cpg, C / 32), dim3(256), 0, st, da, x, stats, workspace, counters, sums, dgamma,
               dbeta, G, Pg, C, cpg, rpc, act, slope);
  const long long total4 = (long long)G * Pg * C / 4;
  if (total4 >= (1LL << 29)) return MDGAN_ERR_UNSUPPORTED;
  MDGAN_LAUNCH(bn_bwd_apply_kernel, dim3(bn_apply_blocks(total4)), dim3(256), 0, st, da, x, stats, sums, dx, Pg, C, total4,
               act, slope, round_tf32);
  return 0;
}

extern "C" int mdgan_bn_bwd_finalize(const float* partial, int phases, int row_tiles, int tiles_per_group, int col_stride,
                                     float* sums, float* dgamma, float* dbeta, int G, int C, void* stream) {
  if (!partial || !sums) return MDGAN_ERR_BAD_ARG;
  if (C % 8 != 0 || G < 1 || phases < 1 || tiles_per_group < 1 || G * tiles_per_group > row_tiles || C > col_stride)
    return MDGAN_ERR_UNSUPPORTED;
  MDGAN_LAUNCH(bn_bwd_finalize_kernel, dim3(C / 8), dim3(256), 0, (cudaStream_t)stream, partial, phases, row_tiles,
               tiles_per_group, col_stride, sums, dgamma, dbeta, G, C);
  return 0;
}

extern "C" int mdgan_bn_bwd_apply_dy(const float* dy, const float* x, const float* stats, const float* sums, float* dx,
                                     int G, int Pg, int C, int round_tf32, void* stream) {
  if (!dy || !x || !stats || !sums || !dx) return MDGAN_ERR_BAD_ARG;
  if (C % 4 != 0 || G < 1 || Pg < 1) return MDGAN_ERR_UNSUPPORTED;
  const long long total4 = (long long)G * Pg * C / 4;
  if (total4 >= (1LL << 29)) return MDGAN_ERR_UNSUPPORTED;
  MDGAN_LAUNCH(bn_bwd_apply_dy_kernel, dim3(bn_apply_blocks(total4)), dim3(256), 0, (cudaStream_t)stream, dy, x, stats, sums,
               dx, Pg, C, total4, round_tf32);
  return 0;
}

extern "C" int mdgan_act_backward(const float* da, const float* a, float* dz, long long n, int act, float slope,
                                  int round_tf32, void* stream) {
  if (!da || !a || !dz || n % 4 != 0) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(act_bwd_kernel, dim3(blocks_for(n / 4, 256)), dim3(256), 0, (cudaStream_t)stream, da, a, dz, n / 4, act,
               slope, round_tf32);
  return 0;
}

extern "C" int mdgan_tanh_backward(const float* s, const float* x, float* out, long long n, float scale,
                                   void* stream) {
  if (!s || !x || !out) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(tanh_bwd_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, s, x, out, n, scale);
  return 0;
}

extern "C" int mdgan_head_pack(const float* w, float* wt, int HW, int C, void* stream) {
  if (!w || !wt || HW <= 0 || C <= 0) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(head_pack_kernel, dim3(blocks_for((long long)HW * C, 256)), dim3(256), 0, (cudaStream_t)stream, w, wt, HW, C);
  return 0;
}

extern "C" int mdgan_head_forward(const float* a, const float* wt, const float* label, float* prob, float* loss_terms,
                                  float* dlogit, float* loss, unsigned int* counter, int G, int b, int HW, int C,
                                  void* stream) {
  if (!a || !wt || !label || !prob || !loss_terms || !dlogit || !loss || !counter) return MDGAN_ERR_BAD_ARG;
  if (C % 4 != 0) return MDGAN_ERR_UNSUPPORTED;
  const int n_total = G * b;
  MDGAN_LAUNCH(head_fwd_kernel, dim3(n_total), dim3(256), 0, (cudaStream_t)stream, a, wt, label, prob, loss_terms, dlogit,
               loss, counter, n_total, b, G, HW * C);
  return 0;
}

extern "C" int mdgan_head_backward(const float* a, const float* wt, const float* dlogit, float* da, float* dw,
                                   int n_total, int HW, int C, void* stream) {
  if (!a || !wt || !dlogit || !da) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(head_bwd_kernel, dim3(blocks_for((long long)HW * C, 128)), dim3(128), 0, (cudaStream_t)stream, a, wt, dlogit,
               da, dw, n_total, HW, C);
  return 0;
}

extern "C" int mdgan_adam_step(float* p, const float* g, float* m, float* v, long long n, int* step_count, float lr,
                               float beta1, float beta2, float eps, void* stream) {
  if (!p || !g || !m || !v || !step_count) return MDGAN_ERR_BAD_ARG;
  const int vec4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  unsigned blocks = blocks_for(vec4 ? (n + 3) / 4 : n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  MDGAN_LAUNCH(adam_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, step_count, lr, beta1, beta2,
               eps, vec4);
  return 0;
}

extern "C" int mdgan_pad_rows(const float* in, float* out, int rows, int cols_in, int cols_out, int round_tf32,
                              void* stream) {
  if (!in || !out || cols_out < cols_in) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(pad_rows_kernel, dim3(blocks_for((long long)rows * cols_out, 256)), dim3(256), 0, (cudaStream_t)stream, in,
               out, rows, cols_in, cols_out, round_tf32);
  return 0;
}

extern "C" int mdgan_sum_slices(const float* in, float* out, long long n, int count, long long stride, void* stream) {
  if (!in || !out) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(sum_slices_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, in, out, n, count, stride);
  return 0;
}
