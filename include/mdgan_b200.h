/* mdgan_b200.h -- C ABI of libmdgan_b200.so, the sm_100a kernel library behind the MD-GAN training step.
 *
 * The reference (owengombas/distributed-gan) has no native code and no FFI: every number on its hot path is
 * produced by a torch library call made from src/actors/{server,worker}.py and the model files
 * src/datasets/{CIFAR10,CelebA,MNIST}.py.  Each entry point below therefore names the torch call site(s) it
 * replaces (reference file:line, relative to /root/reference).  The binding a maintainer would add on the
 * reference side is a ctypes stub (Python is the reference's own language); see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers to fp32 unless stated otherwise;
 *   - activations are NHWC contiguous, image-side tensors (real / generated batches, feedback) NCHW contiguous,
 *     parameters keep the PyTorch layout of the reference's state_dict;
 *   - `stream` is a cudaStream_t (0 = default stream); launches are asynchronous and stream-ordered, the
 *     library keeps no global state besides a tensor-map cache and is safe to capture into a CUDA graph;
 *   - every launcher returns int: 0 = ok, > 0 = cudaError_t, < 0 = MDGAN_ERR_* below.  Nothing throws, nothing
 *     falls back to another implementation: unsupported shapes are errors.
 *   - "round_tf32": store the result rounded (RNA) to TF32 because its consumer is a tensor-core operand
 *     (tcgen05 kind::tf32 would otherwise truncate).
 */
#ifndef MDGAN_B200_H
#define MDGAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MDGAN_ERR_BAD_ARG (-1)
#define MDGAN_ERR_UNSUPPORTED (-2)
#define MDGAN_ERR_DRIVER (-3)

#define MDGAN_MODE_DOWN 0  /* k4 s2 p1 gather: Conv2d forward, ConvTranspose2d data-grad   */
#define MDGAN_MODE_UP 1    /* 4 output phases: ConvTranspose2d forward, Conv2d data-grad   */
#define MDGAN_MODE_DENSE 2 /* plain GEMM: ConvTranspose2d(k,s1,p0) on a 1x1 input          */

/* precision of the tensor-core kernels.  TF32: one tcgen05.mma (kind::tf32) per K-slice on operands their
 * producers rounded to TF32 -- as accurate as cuDNN's TF32 path.  TF32X3: error-compensated 3xTF32 (hi/lo operand
 * split, three MMAs into the fp32 TMEM accumulator) -- ~fp32 accuracy; the parity mode. */
#define MDGAN_PRECISION_TF32 0
#define MDGAN_PRECISION_TF32X3 1

#define MDGAN_ACT_NONE 0
#define MDGAN_ACT_RELU 1
#define MDGAN_ACT_LRELU 2

int mdgan_abi_version(void);
/* 0 iff the current CUDA device is compute capability 10.x (the library only contains sm_100a code). */
int mdgan_check_device(void);

/* ---- weight packing (once per optimiser step) --------------------------------------------------------------
 * Repack a PyTorch-layout conv weight into the K-major operand of mdgan_conv_gemm, rounded to TF32.
 *   mode DOWN : W [N][C][4][4]  -> out [N_pad][16*C_pad]           (Conv2d.weight, or ConvTranspose2d.weight for dgrad)
 *   mode UP   : W [C][N][4][4]  -> out [4*N_pad][4*C_pad]          (ConvTranspose2d.weight, or Conv2d.weight for dgrad)
 *   mode DENSE: W [C][N][k][k]  -> out [KK*N][C_pad]               (first generator layer)
 * Replaces: the implicit weight access of nn.Conv2d / nn.ConvTranspose2d in CIFAR10.py:83-98,116-133,
 * CelebA.py:78-93,113-131.  split = 1 (precision tf32x3): `out` receives two matrices back to back,
 * hi = tf32(w) and lo = tf32(w - hi). */
int mdgan_pack_weights(const float* W, float* out, int mode, int N, int C, int N_pad, int C_pad, int KK, int split,
                       void* stream);
/* All re-packs of one network in one launch.  jobs_dev: n_jobs records of mdgan_pack_job_words() (= 11) int64 words in
 * DEVICE memory: {W ptr, out ptr, mode, N, C, N_pad, C_pad, KK, total elements of the hi matrix, first block, split};
 * job j owns blocks [first_block_j, first_block_{j+1}) of 256 threads; mode 3 = mdgan_head_pack (N = HW, no rounding). */
int mdgan_pack_job_words(void);
int mdgan_pack_weights_multi(const long long* jobs_dev, int n_jobs, int total_blocks, void* stream);

/* ---- implicit-GEMM convolution, tcgen05 (kind::tf32) + TMEM + TMA ---------------------------------------------
 * src  NHWC [n_img][Hs][Ws][C] (C % 32 == 0), wpacked from mdgan_pack_weights, rows = the low-resolution grid
 * (n_img, Hg, Wg).  Output: DOWN/DENSE -> [n_img][Hg][Wg][N]; UP -> [n_img][2Hg][2Wg][N]; NHWC, or NCHW when
 * out_nchw = 1 (image-side outputs).  bias (optional, [N]) is added, act = 1 applies tanh.  accumulate = 1 (NCHW only)
 * adds into dst: the feedbacks of all workers that share one generated batch are summed in place
 * (actors/server.py:271-297).  force_bn = 0 lets the launcher pick the tile width.  gate (optional, NHWC outputs only,
 * same shape as dst): dst = result * act'(gate), gate_act = MDGAN_ACT_RELU / MDGAN_ACT_LRELU(gate_slope) -- the backward
 * of the activation whose OUTPUT is `gate`, fused into the data-gradient GEMM that feeds it.
 * bn_partial (optional; tf32x3 precision, plain NHWC outputs only): [phases][row tiles][2][N_pad] floats -- every CTA
 * also writes the per-column sum and sum of squares of the values it stores (fixed reduction tree), i.e. the train-mode
 * BatchNorm statistics of the layer without another pass over its output; mdgan_bn_finalize turns them into
 * mean / invstd / scale / shift + running statistics.  mdgan_conv_rows_per_tile gives the GEMM rows one CTA owns
 * (row tiles = ceil(n_img*Hg*Wg / rows); 0 = fused statistics unavailable in this mode).
 * bnb_z / bnb_stats (optional, with bn_partial; data-gradient GEMMs): dst is the gradient of a BatchNorm+activation
 * OUTPUT; with the BatchNorm input bnb_z (NHWC like dst), its statistics bnb_stats ([bnb_groups][4][N] from
 * mdgan_bn_forward / mdgan_bn_finalize) and the activation (bnb_act, bnb_slope) the epilogue stores
 * dy = da * act'(z*scale + shift) and bn_partial receives the column sums of dy and dy * xhat -- the reduction pass of
 * the BatchNorm backward; finish with mdgan_bn_bwd_finalize + mdgan_bn_bwd_apply_dy instead of mdgan_bn_backward.
 * Replaces: nn.Conv2d(k4,s2,p1) forward CIFAR10.py:88,92 / CelebA.py:81,85,88; nn.ConvTranspose2d forward
 * CIFAR10.py:118-130 / CelebA.py:113-131 (+ torch.tanh CIFAR10.py:131 / CelebA.py:140); and their data
 * gradients computed by loss.backward() (actors/worker.py:204,227) and torch.autograd.grad (actors/server.py:286). */
int mdgan_conv_gemm(const float* src, const float* wpacked, float* dst, const float* bias, int n_img, int Hg, int Wg,
                    int Hs, int Ws, int C, int mode, int N, int N_pad, int out_nchw, int act, int round_tf32,
                    int accumulate, int precision, int force_bn, const float* gate, int gate_act, float gate_slope,
                    float* bn_partial, const float* bnb_z, const float* bnb_stats, int bnb_act, float bnb_slope,
                    int bnb_groups, void* stream);
int mdgan_conv_rows_per_tile(int Hg, int Wg, int precision);
/* phase slices of bn_partial for this problem: 1 (DOWN / DENSE), 4 (UP), 2 (UP with 64 output channels: one CTA
 * computes both column parities of an output-row parity). */
int mdgan_conv_stat_phases(int mode, int N_pad, int Hg, int Wg, int precision);

/* ---- weight gradient, tcgen05 (kind::tf32, MN-major operands), split-K over pixels ----------------------------
 * partial[split][tap][C1][C2] = sum_p lo[p][c1] * hi[gather(p, tap)][c2]; mode DOWN = 16 taps of the k4 s2 p1
 * stencil (lo [n][Hl][Wl][C1], hi [n][2Hl][2Wl][C2]); mode DENSE = 1 tap, identity gather.  C1 % 128 == 0,
 * C2 % 64 == 0.  mdgan_wgrad_splits gives the slice count to allocate; mdgan_wgrad_unpack reduces the slices in
 * a fixed order and writes the PyTorch-layout gradient ([C1][C2][4][4], or [C1][N][KK] for DENSE, c1 < C1 real).
 * Replaces: the weight gradients of loss.backward() (actors/worker.py:204) and torch.autograd.grad
 * (actors/server.py:286-292). */
int mdgan_wgrad_splits(int n_img, int Hl, int Wl, int C1, int C2, int mode);
int mdgan_wgrad_gemm(const float* lo, const float* hi, float* partial, int n_img, int Hl, int Wl, int C1, int C2,
                     int mode, int splits, int precision, void* stream);
int mdgan_wgrad_unpack(const float* partial, float* grad, int mode, int splits, int C1, int C1p, int C2, int N, int KK,
                       void* stream);
int mdgan_reduce_slices(const float* partial, float* out, int slices, long long n, void* stream);

/* ---- image-side ("thin", 1 or 3 channel) layers on CUDA cores -------------------------------------------------
 * mdgan_thin_down : img NCHW [n][CI][Hi][Wi], W [N][CI][4][4] -> out NHWC [n][Hi/2][Wi/2][N] (+ LeakyReLU).
 *   Replaces nn.Conv2d(3,64,4,2,1)+LeakyReLU CIFAR10.py:85-86 / CelebA.py:78,96 and the data gradient of the
 *   last ConvTranspose2d (CIFAR10.py:130, CelebA.py:131).
 * mdgan_thin_wgrad: feat NHWC [n][Hl][Wl][C1], img NCHW [n][CI][2Hl][2Wl] -> partial [slices][C1][CI*16];
 *   reduce with mdgan_reduce_slices.  Weight gradient of those two layers.
 * mdgan_thin_up   : src NHWC [n][H][W][C], W [C][N][4][4], N in {1,3} -> out NCHW [n][N][2H][2W] (+ tanh, or += ):
 *   the last ConvTranspose2d of the generator with its tanh (CIFAR10.py:130-131, CelebA.py:131,140) and the data
 *   gradient of the first discriminator conv, i.e. the error feedback written (accumulated) into its slot
 *   (actors/worker.py:227-233).  fp32 FMAs on CUDA cores in both precision modes. */
int mdgan_thin_down(const float* img, const float* W, float* out, int n_img, int CI, int Hi, int Wi, int N, int act,
                    float slope, int round_tf32, void* stream);
int mdgan_thin_up(const float* src, const float* W, float* out, int n_img, int H, int Wd, int C, int N, int act_tanh,
                  int accumulate, void* stream);
int mdgan_thin_wgrad_slices(int n_img, int Hl, int Wl);
int mdgan_thin_wgrad(const float* feat, const float* img, float* partial, int n_img, int CI, int Hl, int Wl, int C1,
                     void* stream);

/* ---- train-mode BatchNorm2d fused with the following activation -----------------------------------------------
 * x [G*Pg][C]: G independent passes (e.g. real || X_d) of Pg = b*H*W rows; statistics per pass, running stats
 * updated once per pass in order (momentum, unbiased variance), *num_batches_tracked (int64) += G.
 * stats [G][4][C] = mean, invstd, scale, shift (kept for backward).  workspace: mdgan_bn_workspace_floats floats.
 * counters: C/32 (<= 32) device uint32, zero before the first call and left zero by every call (the last block of a
 * 32-channel slab to finish its partial sums finalizes that slab: 2 launches per pass instead of 3).  C % 32 == 0.
 * Replaces nn.BatchNorm2d + ReLU/LeakyReLU CIFAR10.py:89-94,119-128 / CelebA.py:97-99,134-137 and their backward. */
long long mdgan_bn_workspace_floats(int G, int Pg, int C);
int mdgan_bn_forward(const float* x, float* out, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float* stats, float* workspace,
                     unsigned int* counters, int G, int Pg, int C, float eps, float momentum, int act, float slope,
                     int round_tf32, void* stream);
/* The two halves of mdgan_bn_forward when the producing GEMM already reduced the statistics (mdgan_conv_gemm's
 * bn_partial): mdgan_bn_finalize sums, in fp64 and in a fixed order, the partial slices of every pass g -- slices
 * (phase, tile) with tile in [g*tiles_per_group, (g+1)*tiles_per_group), `fold` column groups of stride C each
 * (fold = k*k for the generator's first layer, whose GEMM columns are (position, channel); 1 otherwise), slice layout
 * [2][col_stride] -- and writes stats / running statistics / num_batches_tracked exactly like mdgan_bn_forward;
 * mdgan_bn_apply is the normalise + activation pass. */
int mdgan_bn_finalize(const float* partial, int phases, int row_tiles, int tiles_per_group, int col_stride, int fold,
                      const float* gamma, const float* beta, float* running_mean, float* running_var,
                      long long* num_batches_tracked, float* stats, int G, int Pg, int C, float eps, float momentum,
                      void* stream);
int mdgan_bn_apply(const float* x, const float* stats, float* out, int G, int Pg, int C, int act, float slope,
                   int round_tf32, void* stream);
int mdgan_bn_backward(const float* da, const float* x, const float* stats, float* dx, float* dgamma, float* dbeta,
                      float* sums, float* workspace, unsigned int* counters, int G, int Pg, int C, int act, float slope,
                      int round_tf32, void* stream);
int mdgan_bn_bwd_finalize(const float* partial, int phases, int row_tiles, int tiles_per_group, int col_stride,
                          float* sums, float* dgamma, float* dbeta, int G, int C, void* stream);
int mdgan_bn_bwd_apply_dy(const float* dy, const float* x, const float* stats, const float* sums, float* dx, int G, int Pg,
                          int C, int round_tf32, void* stream);
int mdgan_act_backward(const float* da, const float* a, float* dz, long long n, int act, float slope, int round_tf32,
                       void* stream);
/* out = s * (1 - x^2) * scale: backward of the generator's tanh on the group-summed feedback, with the
 * 1/(b*N) of actors/server.py:268 folded in. */
int mdgan_tanh_backward(const float* s, const float* x, float* out, long long n, float scale, void* stream);

/* ---- discriminator head: Conv2d(C,1,k,1,0) on a kxk map + Sigmoid + BCELoss -----------------------------------
 * a NHWC [G*b][HW][C]; wt [HW][C] = the PyTorch weight [1][C][k][k] re-ordered by mdgan_head_pack (once per
 * optimiser step); label[g] in {0,1}; dw is written in the PyTorch layout.  prob/loss_terms/dlogit are [G*b];
 * loss[g] = mean BCE of pass g (log clamped at -100), loss[G] = sum over passes.  dlogit already contains 1/b and
 * BCELoss' max(p(1-p), 1e-12) guard.  counter: one device uint32, zero before the first call (the last block to
 * finish reduces the per-sample terms into loss[] in a fixed order and clears it again).
 * Replaces CIFAR10.py:96-97,106 / CelebA.py:91-93,100-101 and nn.BCELoss actors/worker.py:96,199-204,222-227. */
int mdgan_head_pack(const float* w, float* wt, int HW, int C, void* stream);
int mdgan_head_forward(const float* a, const float* wt, const float* label, float* prob, float* loss_terms,
                       float* dlogit, float* loss, unsigned int* counter, int G, int b, int HW, int C, void* stream);
int mdgan_head_backward(const float* a, const float* wt, const float* dlogit, float* da, float* dw, int n_total, int HW,
                        int C, void* stream);

/* ---- the reference's MLP plugin (datasets/MNIST.py:74-120) on CUDA cores --------------------------------------
 * mdgan_sgemm: C [M][N] row-major = epilogue(sum_k A(m,k) B(k,n)) in fp32 FMA, k ascending (bitwise repeatable).
 *   A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs] (element strides), which covers the three products of an
 *   nn.Linear without a transposed copy:  forward y = x W^T (A = x, B(k,n) = W[n][k]);  data gradient dx = dy W;
 *   weight gradient dW = dy^T x, written in the PyTorch [out][in] layout.
 *   Epilogue, in this order: + bias[n] (optional) -> act (MDGAN_ACT_NONE / MDGAN_ACT_LRELU with `slope` /
 *   MDGAN_ACT_TANH) -> dropout keep mask (optional uint8 [M][N]: kept elements times mask_scale = 1/(1-p), dropped
 *   ones zero; the mask is drawn on the host from the worker's torch RNG in the reference's call order, so it equals
 *   F.dropout's noise, MNIST.py:89-94) -> LeakyReLU-backward gate (optional fp32 [M][N]: the gated layer's output; v
 *   where it is > 0, v * gate_slope elsewhere) -> += C (accumulate: the error feedback summed into its slot,
 *   actors/worker.py:227-233).  Replaces nn.Linear + F.leaky_relu + F.dropout + torch.tanh MNIST.py:80-96,106-120 and
 *   their autograd backward (actors/worker.py:204,227; actors/server.py:286-292).
 * mdgan_col_sum: out[n] = sum_m x[m][n]: the bias gradients.
 * mdgan_linear_head_forward / _backward: Linear(L,1) + sigmoid + BCELoss(mean) (MNIST.py:96, actors/worker.py:96,
 *   199-204,222-227) with the conventions of mdgan_head_forward (label per pass, loss[G] = sum, dlogit includes 1/b,
 *   counter zero before the first call); w [L] = Linear.weight [1][L], bias [1].  Backward: da = dlogit * w through the
 *   dropout mask (optional) and the LeakyReLU gate of the layer that produced a; dw [L], dbias [1] optional. */
#define MDGAN_ACT_TANH 3
int mdgan_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int a_rs, int a_cs, int b_rs, int b_cs,
                const float* bias, int act, float slope, const unsigned char* mask, float mask_scale, const float* gate,
                float gate_slope, int accumulate, void* stream);
int mdgan_col_sum(const float* x, float* out, int M, int N, void* stream);
int mdgan_linear_head_forward(const float* a, const float* w, const float* bias, const float* label, float* prob,
                              float* loss_terms, float* dlogit, float* loss, void* counter, int G, int b, int L,
                              void* stream);
int mdgan_linear_head_backward(const float* a, const float* w, const float* dlogit, const unsigned char* mask,
                               float mask_scale, float gate_slope, float* da, float* dw, float* dbias, int n_total, int L,
                               void* stream);

/* ---- torch.optim.Adam on a flat parameter buffer (actors/server.py:111-113,308-312; actors/worker.py:97-99,206).
 * step_count points to two device int32: [0] the number of steps already taken (incremented on the device by the
 * last block of the launch), [1] a block counter that must be zero before the first call. */
int mdgan_adam_step(float* p, const float* g, float* m, float* v, long long n, int* step_count, float lr, float beta1,
                    float beta2, float eps, void* stream);

/* ---- small helpers ----------------------------------------------------------------------------------------- */
int mdgan_pad_rows(const float* in, float* out, int rows, int cols_in, int cols_out, int round_tf32, void* stream);
/* out[i] = sum_k in[k*stride + i]: sums the feedbacks of the workers that share one generated batch
 * (actors/server.py:271-297, collapsed by linearity of the VJP). */
int mdgan_sum_slices(const float* in, float* out, long long n, int count, long long stride, void* stream);

/* ---- peer-memory exchange over NVLink / NVSwitch (one process per GPU; buffers allocated in peer-mapped
 * "symmetric" memory, e.g. torch.distributed._symmetric_memory) ---------------------------------------------------
 * Replaces the per-iteration messages of the reference: generated batches server.py:238-246 / worker.py:181-182,
 * error feedback worker.py:232-233 / server.py:234, and the sum over workers server.py:266-302.  All four are ordinary
 * stream-ordered kernels (capturable in a CUDA graph); no collective library call is involved.
 *   mdgan_peer_signal: *(int*)flag_addrs_dev[i] = *epoch + 1 for i < n (peer-mapped addresses, release at system
 *     scope after a system fence); advance = 1 also stores *epoch + 1 to *epoch.
 *   mdgan_peer_wait  : spins until flags[i] >= *epoch + 1 for i < n (local memory written by peers, acquire at
 *     system scope); advance as above.  A flag that does not arrive within timeout_ms of wall time (a peer process
 *     died) sets *err = 1 and TRAPS the kernel: the context is lost and later calls fail, rather than the rest of
 *     the iteration running on stale data.
 *   mdgan_peer_push  : dst_j[0..n) = src[0..n) for the n_dst peer-mapped destinations dst_addrs_dev[j] (n % 4 == 0).
 *   mdgan_peer_push_multicast: the same through the NVSwitch multicast address mc_dst of the symmetric buffer
 *     (multimem.st: one store leaves the GPU, the switch writes every copy, the local one included).
 *   mdgan_tanh_backward_slices: out[s][j] = scale * (1 - x[s][j]^2) * sum_{w = s, s+k, .. < N} F[w][j], j < n_per_slot:
 *     the feedbacks of the workers sharing generated batch s, summed in ascending worker order, fused with the
 *     generator's tanh backward (F: [N][n_per_slot] slices written by the workers' feedback kernels). */
int mdgan_peer_signal(const unsigned long long* flag_addrs_dev, int n, int* epoch, int advance, void* stream);
int mdgan_peer_wait(const int* flags, int n, int* epoch, int advance, int* err, long long timeout_ms, void* stream);
int mdgan_peer_push(const float* src, const unsigned long long* dst_addrs_dev, int n_dst, long long n, void* stream);
int mdgan_peer_push_multicast(const float* src, float* mc_dst, long long n, void* stream);
int mdgan_tanh_backward_slices(const float* F, const float* x, float* out, long long n_per_slot, int k, int N,
                               float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDGAN_B200_H */
