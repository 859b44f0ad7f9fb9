// Peer-memory exchange of the MD-GAN iteration over NVLink / NVSwitch: the generated batch is pushed from process 0
// into every process' copy of X with plain stores to peer-mapped memory, the workers' feedback kernels store straight
// into process 0's feedback slices, and cross-GPU ordering uses epoch flags in peer-mapped memory (release / acquire
// at system scope) instead of collectives.  Everything here is an ordinary stream-ordered kernel, so the whole
// iteration -- exchange included -- is captured in one CUDA graph.
//   replaces: the isend/recv of generated batches   /root/reference/src/actors/server.py:238-246, worker.py:181-182
//             the send/irecv of the error feedback  /root/reference/src/actors/worker.py:232-233, server.py:234
//             the N retain_graph VJPs + grads_sum   /root/reference/src/actors/server.py:266-302 (sum over workers)
#include "common.cuh"

namespace mdgan {

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// flags[i] (addresses in peer-mapped memory) <- *epoch + 1, after everything this stream did before is visible
// system-wide; advance = 1 also increments *epoch (the last exchange kernel of the iteration on this process).
__global__ void peer_signal_kernel(const unsigned long long* __restrict__ flag_addrs, int n, int* epoch, int advance) {
  pdl_enter();
  const int target = *reinterpret_cast<volatile int*>(epoch) + 1;
  __threadfence_system();
  if ((int)threadIdx.x < n) st_release_sys(reinterpret_cast<int*>(flag_addrs[threadIdx.x]), target);
  __syncthreads();
  if (advance && threadIdx.x == 0) *epoch = target;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until every flags[i] (local memory, written by peers) >= *epoch + 1.  A flag that does not arrive within
// timeout_ns of WALL time (globaltimer) means a peer process died: err[0] is set and the kernel TRAPS -- the context is
// lost and every later call of this process fails loudly, instead of the rest of the captured iteration consuming a
// stale generated batch / feedback and corrupting the discriminator or generator state.
__global__ void peer_wait_kernel(const int* __restrict__ flags, int n, int* epoch, int advance, int* err,
                                 unsigned long long timeout_ns) {
  pdl_enter();
  const int target = *reinterpret_cast<volatile int*>(epoch) + 1;
  if ((int)threadIdx.x < n) {
    const int* f = flags + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < target) {
      __nanosleep(spins < 64 ? 32 : 256);   // a healthy exchange answers within microseconds: poll fast first
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
        atomicExch(err, 1);
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (advance && threadIdx.x == 0) *epoch = target;
}

// dst_j[i] = src[i] for every destination j (peer-mapped addresses; one of them may be local).  float4 body.
__global__ void __launch_bounds__(256) peer_push_kernel(const float* __restrict__ src,
                                                        const unsigned long long* __restrict__ dst_addrs, int n_dst,
                                                        long long n4) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int j = 0; j < n_dst; ++j) reinterpret_cast<float4*>(dst_addrs[j])[i] = v;
  }
}

// mc[i] = src[i] through the NVSwitch multicast mapping of the symmetric buffer: ONE store per element leaves this GPU
// and the switch replicates it into every process' copy (the local one included), instead of one store per peer --
// the generated batch crosses GPU 0's NVLink ports once, not (N - 1) times.
__global__ void __launch_bounds__(256) peer_push_mc_kernel(const float* __restrict__ src, float* mc, long long n4) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float4*>(mc) + i),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
}

// out[s*n_per + j] = scale * (1 - x^2) * sum_{n = s, s+k, ... < N} F[n*n_per + j]: the group sum of the workers'
// feedbacks that share generated batch s (ascending worker order, fixed), fused with the generator's tanh backward.
__global__ void tanh_bwd_slices_kernel(const float* __restrict__ F, const float* __restrict__ x, float* __restrict__ out,
                                       long long n_per4, int k, int N, float scale) {
  pdl_enter();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_per4 * k) return;
  const int s = i / n_per4;
  const long long j = i - (long long)s * n_per4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  bool first = true;
  for (int n = s; n < N; n += k) {
    const float4 f = __ldcg(reinterpret_cast<const float4*>(F) + (long long)n * n_per4 + j);
    if (first) { acc = f; first = false; }
    else { acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w; }
  }
  const float4 xv = reinterpret_cast<const float4*>(x)[i];
  float4 o;
  o.x = acc.x * (1.f - xv.x * xv.x) * scale;
  o.y = acc.y * (1.f - xv.y * xv.y) * scale;
  o.z = acc.z * (1.f - xv.z * xv.z) * scale;
  o.w = acc.w * (1.f - xv.w * xv.w) * scale;
  reinterpret_cast<float4*>(out)[i] = o;
}

}  // namespace mdgan

using namespace mdgan;

extern "C" int mdgan_peer_signal(const unsigned long long* flag_addrs_dev, int n, int* epoch, int advance, void* stream) {
  if (!flag_addrs_dev || !epoch || n < 0 || n > 64) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(peer_signal_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, flag_addrs_dev, n, epoch, advance);
  return 0;
}

extern "C" int mdgan_peer_wait(const int* flags, int n, int* epoch, int advance, int* err, long long timeout_ms,
                               void* stream) {
  if (!flags || !epoch || !err || n < 0 || n > 64 || timeout_ms <= 0) return MDGAN_ERR_BAD_ARG;
  MDGAN_LAUNCH(peer_wait_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, flags, n, epoch, advance, err,
               static_cast<unsigned long long>(timeout_ms) * 1000000ULL);
  return 0;
}

extern "C" int mdgan_peer_push(const float* src, const unsigned long long* dst_addrs_dev, int n_dst, long long n,
                               void* stream) {
  if (!src || !dst_addrs_dev || n_dst < 1 || n_dst > 64 || n % 4 != 0) return MDGAN_ERR_BAD_ARG;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  MDGAN_LAUNCH(peer_push_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, dst_addrs_dev, n_dst,
               n / 4);
  return 0;
}

extern "C" int mdgan_peer_push_multicast(const float* src, float* mc_dst, long long n, void* stream) {
  if (!src || !mc_dst || n % 4 != 0) return MDGAN_ERR_BAD_ARG;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  MDGAN_LAUNCH(peer_push_mc_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, mc_dst, n / 4);
  return 0;
}

extern "C" int mdgan_tanh_backward_slices(const float* F, const float* x, float* out, long long n_per_slot, int k, int N,
                                          float scale, void* stream) {
  if (!F || !x || !out || k < 1 || N < 1 || n_per_slot % 4 != 0) return MDGAN_ERR_BAD_ARG;
  const long long total4 = n_per_slot / 4 * k;
  MDGAN_LAUNCH(tanh_bwd_slices_kernel, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, F, x,
               out, n_per_slot / 4, k, N, scale);
  return 0;
}
