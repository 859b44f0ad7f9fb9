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

// Reduce split-K partial slices and scatter to the PyTorch parameter layout.
// mode 0: grad[c1][c2][tap] = sum_s partial[s][tap][c1][c2]      (C1p x C2 slices, c1 < C1)
// mode 2: grad[c1][n][kk]   = sum_s partial[s][0][c1][kk*N + n]  (dense first generator layer; C2 = KK*N)
__global__ void wgrad_unpack_kernel(const float* __restrict__ partial, float* __restrict__ grad, int mode, int splits,
                                    int taps, int C1, int C1p, int C2, int N, int KK, long long total) {
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

__global__ void reduce_slices_kernel(const float* __restrict__ partial, float* __restrict__ out, int slices,
                                     long long n) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float acc = 0.f;
  for (int s = 0; s < slices; ++s) acc += partial[s * n + idx];
  out[idx] = acc;
}

// ----------------------------------------------------------------------------------------------- BatchNorm (train)
// x is [G*Pg, C] (G independent passes of Pg rows each, e.g. real || X_d).  Stage 1: per-chunk partial sums.
__global__ void bn_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int Pg, int C,
                                  int chunks_per_group, int rows_per_chunk) {
  extern __shared__ float sm[];  // [row_lanes][2][C]
  const int quads = C >> 2;
  const int row_lanes = blockDim.x / quads;
  const int q = threadIdx.x % quads, rl = threadIdx.x / quads;
  const int g = blockIdx.x / chunks_per_group, ch = blockIdx.x % chunks_per_group;
  const int r0 = ch * rows_per_chunk;
  const int r1 = min(Pg, r0 + rows_per_chunk);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (rl < row_lanes) {
    const float* base = x + ((long long)g * Pg) * C + q * 4;
    for (int r = r0 + rl; r < r1; r += row_lanes) {
      const float4 v = *reinterpret_cast<const float4*>(base + (long long)r * C);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sm[(rl * 2 + 0) * C + q * 4 + j] = s[j];
      sm[(rl * 2 + 1) * C + q * 4 + j] = ss[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < row_lanes; ++l) acc += sm[l * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = acc;
  }
}

// Stage 2: per group (in order) mean / biased var -> scale, shift; running stats with momentum 0.1 and unbiased
// variance, num_batches_tracked += G.  stats layout: [G][4][C] = mean, invstd, scale, shift.
// Block = 32 channels x 8 chunk-lanes: the chunk partials of a channel are summed by 8 threads in fp64 and combined
// through shared memory (a single thread walking all ~300 chunks cost 40-60 us of pure latency per launch).
constexpr int kFinLanes = 8;
__global__ void __launch_bounds__(32 * kFinLanes)
bn_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt,
                   float* __restrict__ stats, int G, int Pg, int C, int chunks_per_group, float eps, float momentum) {
  __shared__ double red[2][kFinLanes][32];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += G;
  const bool ok = c < C;
  float rm = 0.f, rv = 0.f;
  if (ok && lane == 0) { rm = running_mean ? running_mean[c] : 0.f; rv = running_var ? running_var[c] : 0.f; }
  for (int g = 0; g < G; ++g) {
    double s = 0.0, ss = 0.0;
    if (ok) {
      for (int ch = lane; ch < chunks_per_group; ch += kFinLanes) {
        const float* pp = partial + ((long long)(g * chunks_per_group + ch)) * 2 * C;
        s += pp[c];
        ss += pp[C + c];
      }
    }
    red[0][lane][cl] = s;
    red[1][lane][cl] = ss;
    __syncthreads();
    if (ok && lane == 0) {
      s = 0.0; ss = 0.0;
#pragma unroll
      for (int l = 0; l < kFinLanes; ++l) { s += red[0][l][cl]; ss += red[1][l][cl]; }
      const double mean = s / Pg;
      double var = ss / Pg - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      const float sc = gamma[c] * invstd;
      float* st = stats + (long long)g * 4 * C;
      st[c] = (float)mean;
      st[C + c] = invstd;
      st[2 * C + c] = sc;
      st[3 * C + c] = beta[c] - (float)mean * sc;
      const float unbiased = (float)(var * ((double)Pg / (double)(Pg > 1 ? Pg - 1 : 1)));
      rm = (1.f - momentum) * rm + momentum * (float)mean;
      rv = (1.f - momentum) * rv + momentum * unbiased;
    }
    __syncthreads();
  }
  if (ok && lane == 0) {
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

// Stage 3: out = act(x*scale + shift)
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, float* __restrict__ out,
                                int Pg, int C, long long total4, int act, float slope, int round_tf32) {
  long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const long long e = i4 * 4;
  const int c = e % C;
  const int g = (e / C) / Pg;
  const float* st = stats + (long long)g * 4 * C;
  const float4 v = *reinterpret_cast<const float4*>(x + e);
  const float4 sc = *reinterpret_cast<const float4*>(st + 2 * C + c);
  const float4 sh = *reinterpret_cast<const float4*>(st + 3 * C + c);
  float4 o;
  o.x = act_fwd(fmaf(v.x, sc.x, sh.x), act, slope);
  o.y = act_fwd(fmaf(v.y, sc.y, sh.y), act, slope);
  o.z = act_fwd(fmaf(v.z, sc.z, sh.z), act, slope);
  o.w = act_fwd(fmaf(v.w, sc.w, sh.w), act, slope);
  if (round_tf32) { o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w); }
  *reinterpret_cast<float4*>(out + e) = o;
}

// Backward stage 1: dy = da * act'(y); partial sums of dy and dy*xhat per channel.
__global__ void bn_bwd_partial_kernel(const float* __restrict__ da, const float* __restrict__ x,
                                      const float* __restrict__ stats, float* __restrict__ partial, int Pg, int C,
                                      int chunks_per_group, int rows_per_chunk, int act, float slope) {
  extern __shared__ float sm[];
  const int quads = C >> 2;
  const int row_lanes = blockDim.x / quads;
  const int q = threadIdx.x % quads, rl = threadIdx.x / quads;
  const int g = blockIdx.x / chunks_per_group, ch = blockIdx.x % chunks_per_group;
  const int r0 = ch * rows_per_chunk;
  const int r1 = min(Pg, r0 + rows_per_chunk);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (rl < row_lanes) {
    const float* st = stats + (long long)g * 4 * C + q * 4;
    float mean[4], invstd[4], sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mean[j] = st[j]; invstd[j] = st[C + j]; sc[j] = st[2 * C + j]; sh[j] = st[3 * C + j]; }
    const long long base = ((long long)g * Pg) * C + q * 4;
    for (int r = r0 + rl; r < r1; r += row_lanes) {
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sm[(rl * 2 + 0) * C + q * 4 + j] = s[j];
      sm[(rl * 2 + 1) * C + q * 4 + j] = ss[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < row_lanes; ++l) acc += sm[l * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = acc;
  }
}

// Backward stage 2: sums[g][2][C] (sum dy, sum dy*xhat); dgamma/dbeta = totals over all groups (optional).
// Same 32-channel x 8-lane layout as bn_finalize_kernel.
__global__ void __launch_bounds__(32 * kFinLanes)
bn_bwd_finalize_kernel(const float* __restrict__ partial, float* __restrict__ sums, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, int G, int C, int chunks_per_group) {
  __shared__ double red[2][kFinLanes][32];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool ok = c < C;
  double tg = 0.0, tb = 0.0;
  for (int g = 0; g < G; ++g) {
    double s = 0.0, ss = 0.0;
    if (ok) {
      for (int ch = lane; ch < chunks_per_group; ch += kFinLanes) {
        const float* pp = partial + ((long long)(g * chunks_per_group + ch)) * 2 * C;
        s += pp[c];
        ss += pp[C + c];
      }
    }
    red[0][lane][cl] = s;
    red[1][lane][cl] = ss;
    __syncthreads();
    if (ok && lane == 0) {
      s = 0.0; ss = 0.0;
#pragma unroll
      for (int l = 0; l < kFinLanes; ++l) { s += red[0][l][cl]; ss += red[1][l][cl]; }
      sums[(long long)g * 2 * C + c] = (float)s;
      sums[(long long)g * 2 * C + C + c] = (float)ss;
      tb += s;
      tg += ss;
    }
    __syncthreads();
  }
  if (ok && lane == 0) {
    if (dgamma) dgamma[c] = (float)tg;
    if (dbeta) dbeta[c] = (float)tb;
  }
}

// Backward stage 3: dx = scale * (dy - mean(dy) - xhat * mean(dy*xhat))
__global__ void bn_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ x,
                                    const float* __restrict__ stats, const float* __restrict__ sums,
                                    float* __restrict__ dx, int Pg, int C, long long total4, int act, float slope,
                                    int round_tf32) {
  long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const long long e = i4 * 4;
  const int c = e % C;
  const int g = (e / C) / Pg;
  const float* st = stats + (long long)g * 4 * C + c;
  const float* sm = sums + (long long)g * 2 * C + c;
  const float4 xv = *reinterpret_cast<const float4*>(x + e);
  const float4 dv = *reinterpret_cast<const float4*>(da + e);
  const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
  const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
  float o[4];
  const float invP = 1.f / (float)Pg;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float mean = st[j], invstd = st[C + j], sc = st[2 * C + j], sh = st[3 * C + j];
    const float y = fmaf(xs[j], sc, sh);
    const float dy = ds[j] * act_grad(y, act, slope);
    const float xhat = (xs[j] - mean) * invstd;
    o[j] = sc * (dy - sm[j] * invP - xhat * (sm[C + j] * invP));
    if (round_tf32) o[j] = to_tf32(o[j]);
  }
  *reinterpret_cast<float4*>(dx + e) = make_float4(o[0], o[1], o[2], o[3]);
}

// dz = da * act'(a) for an activation with no BatchNorm in front (a = act(z); sign(a) == sign(z) for slope > 0).
__global__ void act_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a, float* __restrict__ dz,
                               long long total4, int act, float slope, int round_tf32) {
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
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i];
  out[i] = s[i] * (1.f - xv * xv) * scale;
}

// ----------------------------------------------------------------------------------------------- discriminator head
// wt[hw*C + c] = w[c*HW + hw]: the head weight re-ordered once per optimiser step to the NHWC order of the activations
// so that the per-sample dot products below read both operands with coalesced float4 loads.
__global__ void head_pack_kernel(const float* __restrict__ w, float* __restrict__ wt, int HW, int C) {
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
                float* __restrict__ prob, float* __restrict__ loss_terms, float* __restrict__ dlogit, int n_total, int b,
                int L) {
  __shared__ float red[8];
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
  }
}

// loss[g] = mean over the b samples of group g; loss[G] = sum over groups (the reference's d_loss).
__global__ void head_loss_kernel(const float* __restrict__ loss_terms, float* __restrict__ loss, int G, int b) {
  __shared__ float red[32];
  float total = 0.f;
  for (int g = 0; g < G; ++g) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < b; i += blockDim.x) acc += loss_terms[g * b + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int wv = 0; wv < (blockDim.x >> 5); ++wv) s += red[wv];
      s /= (float)b;
      loss[g] = s;
      total += s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[G] = total;
}

// da[n, l] = dlogit[n] * wt[l];  dw[c*HW + hw] = sum_n dlogit[n] * a[n, l]  (dw optional, PyTorch layout)
__global__ void head_bwd_kernel(const float* __restrict__ a, const float* __restrict__ wt,
                                const float* __restrict__ dlogit, float* __restrict__ da, float* __restrict__ dw,
                                int n_total, int HW, int C) {
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
// step_count lives on the device so the launch is CUDA-graph friendly; adam_tick_kernel increments it afterwards.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, int* __restrict__ step_count, float lr, float beta1,
                            float beta2, float eps) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const int t = *step_count + 1;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * beta2 + (1.f - beta2) * gi * gi;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}
__global__ void adam_tick_kernel(int* step_count) { *step_count += 1; }

// ----------------------------------------------------------------------------------------------- misc
// out[r][c] (cols_out >= cols_in, zero padded), optionally rounded to TF32: pads z [kb, 100] to [kb, 128].
__global__ void pad_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols_in,
                                int cols_out, int round_tf32) {
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
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int k = 0; k < count; ++k) acc += in[k * stride + i];
  out[i] = acc;
}

}  // namespace mdgan

using namespace mdgan;

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

extern "C" int mdgan_pack_weights(const float* W, float* out, int mode, int N, int C, int N_pad, int C_pad, int KK,
                                  int split, void* stream) {
  if (!W || !out || mode < 0 || mode > 2) return MDGAN_ERR_BAD_ARG;
  long long total;
  if (mode == 0) total = (long long)N_pad * 16 * C_pad;
  else if (mode == 1) total = 4LL * N_pad * 4 * C_pad;
  else total = (long long)KK * N * C_pad;
  pack_weights_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(W, out, mode, N, C, N_pad, C_pad, KK,
                                                                                total, split);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_wgrad_unpack(const float* partial, float* grad, int mode, int splits, int C1, int C1p, int C2,
                                  int N, int KK, void* stream) {
  if (!partial || !grad || (mode != 0 && mode != 2)) return MDGAN_ERR_BAD_ARG;
  const int taps = mode == 0 ? 16 : 1;
  const long long total = mode == 0 ? (long long)C1 * C2 * 16 : (long long)C1 * N * KK;
  wgrad_unpack_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(partial, grad, mode, splits, taps, C1,
                                                                                C1p, C2, N, KK, total);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_reduce_slices(const float* partial, float* out, int slices, long long n, void* stream) {
  reduce_slices_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(partial, out, slices, n);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

// Chunking shared by the BN forward and backward reductions: enough blocks to cover the machine, whole rows.
static void bn_chunks(int Pg, int G, int* chunks_per_group, int* rows_per_chunk) {
  int cpg = (296 + G - 1) / G;
  if (cpg > (Pg + 31) / 32) cpg = (Pg + 31) / 32;
  if (cpg < 1) cpg = 1;
  *rows_per_chunk = (Pg + cpg - 1) / cpg;
  *chunks_per_group = (Pg + *rows_per_chunk - 1) / *rows_per_chunk;
}

extern "C" long long mdgan_bn_workspace_floats(int G, int Pg, int C) {
  int cpg, rpc;
  bn_chunks(Pg, G, &cpg, &rpc);
  return (long long)G * cpg * 2 * C;
}

extern "C" int mdgan_bn_forward(const float* x, float* out, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, long long* num_batches_tracked, float* stats, float* workspace,
                                int G, int Pg, int C, float eps, float momentum, int act, float slope, int round_tf32,
                                void* stream) {
  if (!x || !out || !gamma || !beta || !stats || !workspace) return MDGAN_ERR_BAD_ARG;
  if (C % 4 != 0 || C > 1024 || 256 % (C / 4) != 0) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int cpg, rpc;
  bn_chunks(Pg, G, &cpg, &rpc);
  const int row_lanes = 256 / (C / 4);
  bn_partial_kernel<<<G * cpg, 256, row_lanes * 2 * C * sizeof(float), st>>>(x, workspace, Pg, C, cpg, rpc);
  MDGAN_CHECK_LAUNCH();
  bn_finalize_kernel<<<blocks_for(C, 32), 32 * kFinLanes, 0, st>>>(workspace, gamma, beta, running_mean, running_var,
                                                         num_batches_tracked, stats, G, Pg, C, cpg, eps, momentum);
  MDGAN_CHECK_LAUNCH();
  const long long total4 = (long long)G * Pg * C / 4;
  bn_apply_kernel<<<blocks_for(total4, 256), 256, 0, st>>>(x, stats, out, Pg, C, total4, act, slope, round_tf32);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_bn_backward(const float* da, const float* x, const float* stats, float* dx, float* dgamma,
                                 float* dbeta, float* sums, float* workspace, int G, int Pg, int C, int act,
                                 float slope, int round_tf32, void* stream) {
  if (!da || !x || !stats || !dx || !sums || !workspace) return MDGAN_ERR_BAD_ARG;
  if (C % 4 != 0 || C > 1024 || 256 % (C / 4) != 0) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int cpg, rpc;
  bn_chunks(Pg, G, &cpg, &rpc);
  const int row_lanes = 256 / (C / 4);
  bn_bwd_partial_kernel<<<G * cpg, 256, row_lanes * 2 * C * sizeof(float), st>>>(da, x, stats, workspace, Pg, C, cpg,
                                                                               rpc, act, slope);
  MDGAN_CHECK_LAUNCH();
  bn_bwd_finalize_kernel<<<blocks_for(C, 32), 32 * kFinLanes, 0, st>>>(workspace, sums, dgamma, dbeta, G, C, cpg);
  MDGAN_CHECK_LAUNCH();
  const long long total4 = (long long)G * Pg * C / 4;
  bn_bwd_apply_kernel<<<blocks_for(total4, 256), 256, 0, st>>>(da, x, stats, sums, dx, Pg, C, total4, act, slope,
                                                               round_tf32);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_act_backward(const float* da, const float* a, float* dz, long long n, int act, float slope,
                                  int round_tf32, void* stream) {
  if (!da || !a || !dz || n % 4 != 0) return MDGAN_ERR_BAD_ARG;
  act_bwd_kernel<<<blocks_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(da, a, dz, n / 4, act, slope, round_tf32);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_tanh_backward(const float* s, const float* x, float* out, long long n, float scale,
                                   void* stream) {
  if (!s || !x || !out) return MDGAN_ERR_BAD_ARG;
  tanh_bwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s, x, out, n, scale);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_head_pack(const float* w, float* wt, int HW, int C, void* stream) {
  if (!w || !wt || HW <= 0 || C <= 0) return MDGAN_ERR_BAD_ARG;
  head_pack_kernel<<<blocks_for((long long)HW * C, 256), 256, 0, (cudaStream_t)stream>>>(w, wt, HW, C);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_head_forward(const float* a, const float* wt, const float* label, float* prob, float* loss_terms,
                                  float* dlogit, float* loss, int G, int b, int HW, int C, void* stream) {
  if (!a || !wt || !label || !prob || !loss_terms || !dlogit || !loss) return MDGAN_ERR_BAD_ARG;
  if (C % 4 != 0) return MDGAN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_total = G * b;
  head_fwd_kernel<<<n_total, 256, 0, st>>>(a, wt, label, prob, loss_terms, dlogit, n_total, b, HW * C);
  MDGAN_CHECK_LAUNCH();
  head_loss_kernel<<<1, 256, 0, st>>>(loss_terms, loss, G, b);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_head_backward(const float* a, const float* wt, const float* dlogit, float* da, float* dw,
                                   int n_total, int HW, int C, void* stream) {
  if (!a || !wt || !dlogit || !da) return MDGAN_ERR_BAD_ARG;
  head_bwd_kernel<<<blocks_for((long long)HW * C, 128), 128, 0, (cudaStream_t)stream>>>(a, wt, dlogit, da, dw, n_total,
                                                                                       HW, C);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_adam_step(float* p, const float* g, float* m, float* v, long long n, int* step_count, float lr,
                               float beta1, float beta2, float eps, void* stream) {
  if (!p || !g || !m || !v || !step_count) return MDGAN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned blocks = blocks_for(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, step_count, lr, beta1, beta2, eps);
  MDGAN_CHECK_LAUNCH();
  adam_tick_kernel<<<1, 1, 0, st>>>(step_count);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_pad_rows(const float* in, float* out, int rows, int cols_in, int cols_out, int round_tf32,
                              void* stream) {
  if (!in || !out || cols_out < cols_in) return MDGAN_ERR_BAD_ARG;
  pad_rows_kernel<<<blocks_for((long long)rows * cols_out, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols_in,
                                                                                                cols_out, round_tf32);
  MDGAN_CHECK_LAUNCH();
  return 0;
}

extern "C" int mdgan_sum_slices(const float* in, float* out, long long n, int count, long long stride, void* stream) {
  if (!in || !out) return MDGAN_ERR_BAD_ARG;
  sum_slices_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, count, stride);
  MDGAN_CHECK_LAUNCH();
  return 0;
}
