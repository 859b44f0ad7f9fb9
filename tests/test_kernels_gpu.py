"""Per-kernel parity: every sm_100a kernel family against the torch op it replaces (fp64 CPU truth).

Tolerances (scale-normalised max error |got - ref|max / |ref|max against an fp64 reference):
  precision tf32   : single-pass kind::tf32 on operands rounded to nearest -> ~5e-4 measured, bar 3e-3;
  precision tf32x3 : error-compensated 3xTF32 (the default / parity mode) -> ~fp32, bar 2e-5;
  CUDA-core kernels: 2e-5 (1e-4 for quantities with cancellation such as BatchNorm backward).
"""
import pytest
import torch
import torch.nn.functional as F

from util import nchw, nhwc, relerr

pytestmark = pytest.mark.gpu
TF32_TOL = 3e-3
FP32_TOL = 2e-5
PRECISIONS = [pytest.param(0, TF32_TOL, id="tf32"), pytest.param(1, FP32_TOL, id="tf32x3")]


@pytest.fixture(scope="module")
def ops():
    from mdgan_b200 import _lib, ops as _ops

    assert _lib.load().mdgan_check_device() == 0, "libmdgan_b200.so only contains sm_100a code"
    return _ops


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("n,C,N,H,bn", [(8, 64, 128, 16, 0), (8, 64, 128, 16, 128), (8, 64, 128, 16, 32),
                                        (4, 128, 256, 8, 0), (3, 256, 512, 8, 0), (64, 64, 128, 32, 0)])
@pytest.mark.parametrize("prec,tol", PRECISIONS)
def test_conv_down(ops, dev, n, C, N, H, bn, prec, tol):
    torch.manual_seed(1)
    x, W = torch.randn(n, C, H, H), torch.randn(N, C, 4, 4) * 0.05
    bias = torch.randn(N) * 0.1
    ref = F.conv2d(x.double(), W.double(), bias.double(), stride=2, padding=1)
    out = torch.empty(n, H // 2, H // 2, N, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), ops.pack_down(W.to(dev), precision=prec), ops.MODE_DOWN, N, out,
                  (n, H // 2, H // 2), (H, H), bias=bias.to(dev), precision=prec, force_bn=bn)
    assert relerr(nchw(out), ref) < tol


@pytest.mark.parametrize("n,C,N,H,nchw_out", [(8, 512, 256, 4, False), (8, 256, 128, 8, False), (4, 128, 64, 16, False),
                                              (4, 64, 3, 32, True), (4, 128, 3, 16, True), (5, 128, 1, 14, True),
                                              (3, 256, 128, 7, False)])
@pytest.mark.parametrize("prec,tol", PRECISIONS)
def test_conv_up(ops, dev, n, C, N, H, nchw_out, prec, tol):
    torch.manual_seed(2)
    x, W = torch.randn(n, C, H, H), torch.randn(C, N, 4, 4) * 0.05
    ref = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1)
    shape = (n, N, 2 * H, 2 * H) if nchw_out else (n, 2 * H, 2 * H, N)
    out = torch.empty(shape, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), ops.pack_up(W.to(dev), precision=prec), ops.MODE_UP, N, out, (n, H, H), (H, H),
                  out_nchw=nchw_out, precision=prec)
    assert relerr(out if nchw_out else nchw(out), ref) < tol


@pytest.mark.parametrize("N", [64, 128])
def test_conv_up_gated(ops, dev, N):
    """gate: the LeakyReLU backward of the layer below, fused into the data-gradient GEMM's epilogue."""
    torch.manual_seed(14)
    n, C, H = 5, 128, 7
    x, W = torch.randn(n, C, H, H), torch.randn(C, N, 4, 4) * 0.05
    a = torch.randn(n, N, 2 * H, 2 * H)  # the activation OUTPUT of the layer below (sign == sign of its input)
    ref = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1) * torch.where(a > 0, 1.0, 0.2).double()
    out = torch.empty(n, 2 * H, 2 * H, N, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), ops.pack_up(W.to(dev), precision=1), ops.MODE_UP, N, out, (n, H, H), (H, H), precision=1,
                  gate=nhwc(a).to(dev), gate_act=ops.ACT_LRELU, gate_slope=0.2)
    assert relerr(nchw(out), ref) < 2e-5


def test_conv_up_accumulate(ops, dev):
    """accumulate=1: dst += result (how feedbacks of workers sharing a generated batch are summed)."""
    torch.manual_seed(12)
    n, C, N, H = 4, 64, 3, 16
    x, W = torch.randn(n, C, H, H), torch.randn(C, N, 4, 4) * 0.05
    ref = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1)
    base = torch.randn(n, N, 2 * H, 2 * H)
    out = base.clone().to(dev)
    wq = ops.pack_up(W.to(dev), precision=1)
    for _ in range(2):
        ops.conv_gemm(nhwc(x).to(dev), wq, ops.MODE_UP, N, out, (n, H, H), (H, H), out_nchw=True, accumulate=True,
                      precision=1)
    assert relerr(out, base.double() + 2 * ref) < FP32_TOL


def test_conv_up_tanh(ops, dev):
    torch.manual_seed(3)
    n, C, N, H = 4, 128, 3, 16
    x, W = torch.randn(n, C, H, H), torch.randn(C, N, 4, 4) * 0.05
    ref = torch.tanh(F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1))
    out = torch.empty(n, N, 2 * H, 2 * H, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), ops.pack_up(W.to(dev)), ops.MODE_UP, N, out, (n, H, H), (H, H), out_nchw=True,
                  act_tanh=True)
    assert relerr(out, ref) < TF32_TOL


@pytest.mark.parametrize("n,C,N,k", [(16, 100, 512, 4), (130, 100, 256, 7), (128, 100, 512, 4)])
@pytest.mark.parametrize("prec,tol", PRECISIONS)
def test_conv_dense(ops, dev, n, C, N, k, prec, tol):
    torch.manual_seed(4)
    z, W = torch.randn(n, C, 1, 1), torch.randn(C, N, k, k) * 0.05
    ref = F.conv_transpose2d(z.double(), W.double())
    wp = ops.pack_dense(W.to(dev), precision=prec)
    zp = torch.zeros(n, wp.shape[1], device=dev)
    ops.pad_rows(z.view(n, C).to(dev), zp, round_tf32=(prec == 0))
    out = torch.empty(n, k, k, N, device=dev)
    ops.conv_gemm(zp, wp, ops.MODE_DENSE, k * k * N, out, (n, 1, 1), (1, 1), precision=prec)
    assert relerr(nchw(out), ref) < tol


@pytest.mark.parametrize("n,C1,C2,Hl", [(8, 128, 64, 8), (8, 256, 128, 4), (16, 128, 64, 16), (5, 512, 256, 4),
                                        (3, 128, 64, 7)])
@pytest.mark.parametrize("prec,tol", PRECISIONS)
def test_wgrad(ops, dev, n, C1, C2, Hl, prec, tol):
    torch.manual_seed(5)
    x, dout = torch.randn(n, C2, 2 * Hl, 2 * Hl), torch.randn(n, C1, Hl, Hl)
    W = torch.zeros(C1, C2, 4, 4, dtype=torch.double, requires_grad=True)
    F.conv2d(x.double(), W, stride=2, padding=1).backward(dout.double())
    splits = ops.wgrad_splits(n, Hl, Hl, C1, C2, ops.MODE_DOWN)
    partial = torch.empty(splits * 16 * C1 * C2, device=dev)
    grad = torch.empty(C1, C2, 4, 4, device=dev)
    ops.wgrad_gemm(nhwc(dout).to(dev), nhwc(x).to(dev), partial, (n, Hl, Hl), ops.MODE_DOWN, splits, precision=prec)
    ops.wgrad_unpack(partial, grad, ops.MODE_DOWN, splits, C1, C1, C2)
    assert relerr(grad, W.grad) < tol


@pytest.mark.parametrize("prec,tol", PRECISIONS)
def test_wgrad_dense(ops, dev, prec, tol):
    torch.manual_seed(6)
    n, C, N, k = 32, 100, 512, 4
    z, dout = torch.randn(n, C, 1, 1), torch.randn(n, N, k, k)
    W = torch.zeros(C, N, k, k, dtype=torch.double, requires_grad=True)
    F.conv_transpose2d(z.double(), W).backward(dout.double())
    zp = torch.zeros(n, 128, device=dev)
    ops.pad_rows(z.view(n, C).to(dev), zp, round_tf32=(prec == 0))
    hi = nhwc(dout).to(dev).view(n, 1, 1, k * k * N)
    splits = ops.wgrad_splits(n, 1, 1, 128, k * k * N, ops.MODE_DENSE)
    partial = torch.empty(splits * 128 * k * k * N, device=dev)
    grad = torch.empty(C, N, k, k, device=dev)
    ops.wgrad_gemm(zp.view(n, 1, 1, 128), hi, partial, (n, 1, 1), ops.MODE_DENSE, splits, precision=prec)
    ops.wgrad_unpack(partial, grad, ops.MODE_DENSE, splits, C, 128, k * k * N, N=N, KK=k * k)
    assert relerr(grad, W.grad) < tol


@pytest.mark.parametrize("CI,N,H", [(3, 64, 32), (3, 128, 16), (1, 64, 28), (1, 128, 28)])
def test_thin_down_and_wgrad(ops, dev, CI, N, H):
    torch.manual_seed(7)
    n = 6
    img, W = torch.randn(n, CI, H, H), torch.randn(N, CI, 4, 4) * 0.1
    ref = F.leaky_relu(F.conv2d(img.double(), W.double(), stride=2, padding=1), 0.2)
    out = torch.empty(n, H // 2, H // 2, N, device=dev)
    ops.thin_down(img.to(dev), W.to(dev), out, act=ops.ACT_LRELU, slope=0.2)
    assert relerr(nchw(out), ref) < FP32_TOL
    dout = torch.randn(n, N, H // 2, H // 2)
    Wd = torch.zeros(N, CI, 4, 4, dtype=torch.double, requires_grad=True)
    F.conv2d(img.double(), Wd, stride=2, padding=1).backward(dout.double())
    partial = torch.empty(ops.thin_wgrad_slices(n, H // 2, H // 2) * N * CI * 16, device=dev)
    grad = torch.empty(N, CI, 4, 4, device=dev)
    ops.thin_wgrad(nhwc(dout).to(dev), img.to(dev), partial, grad)
    assert relerr(grad, Wd.grad) < FP32_TOL


@pytest.mark.parametrize("n,C,N,H", [(4, 64, 3, 32), (4, 128, 3, 16), (5, 128, 1, 14), (3, 64, 1, 14), (64, 64, 3, 16)])
def test_thin_up(ops, dev, n, C, N, H):
    """ConvTranspose2d(C -> N, 4, 2, 1) (+ tanh) / the data gradient of Conv2d(N -> C, 4, 2, 1), NCHW output, += mode."""
    torch.manual_seed(13)
    x, W = torch.randn(n, C, H, H), torch.randn(C, N, 4, 4) * 0.05
    ref = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1)
    out = torch.empty(n, N, 2 * H, 2 * H, device=dev)
    ops.thin_up(nhwc(x).to(dev), W.to(dev), out)
    assert relerr(out, ref) < FP32_TOL
    ops.thin_up(nhwc(x).to(dev), W.to(dev), out, act_tanh=True)
    assert relerr(out, torch.tanh(ref)) < FP32_TOL
    base = torch.randn(n, N, 2 * H, 2 * H)
    out = base.clone().to(dev)
    for _ in range(2):
        ops.thin_up(nhwc(x).to(dev), W.to(dev), out, accumulate=True)
    assert relerr(out, base.double() + 2 * ref) < FP32_TOL
    # the same call as the data gradient of the first discriminator conv
    xin = torch.randn(n, N, 2 * H, 2 * H, dtype=torch.double, requires_grad=True)
    Wc = torch.randn(C, N, 4, 4) * 0.05
    F.conv2d(xin, Wc.double(), stride=2, padding=1).backward(x.double())
    ops.thin_up(nhwc(x).to(dev), Wc.to(dev), out)
    assert relerr(out, xin.grad) < FP32_TOL


@pytest.mark.parametrize("G,b,H,C,act,slope", [(2, 8, 8, 128, 2, 0.2), (1, 16, 4, 512, 1, 0.0), (1, 4, 16, 64, 1, 0.0)])
def test_batchnorm_fwd_bwd(ops, dev, G, b, H, C, act, slope):
    torch.manual_seed(8)
    x = torch.randn(G * b, C, H, H) * 1.7 + 0.3
    gamma, beta = torch.randn(C) * 0.1 + 1, torch.randn(C) * 0.1
    dout = torch.randn(G * b, C, H, H)
    bn = torch.nn.BatchNorm2d(C).double()
    bn.weight.data.copy_(gamma)
    bn.bias.data.copy_(beta)
    actf = (lambda t: F.leaky_relu(t, slope)) if act == 2 else F.relu
    outs, dxs = [], []
    for g in range(G):  # G independent train-mode passes, running stats updated in order
        xg = x[g * b:(g + 1) * b].double().requires_grad_(True)
        y = actf(bn(xg))
        y.backward(dout[g * b:(g + 1) * b].double())
        outs.append(y.detach())
        dxs.append(xg.grad)
    Pg = b * H * H
    xd = nhwc(x).to(dev)
    out, dx = torch.empty_like(xd), torch.empty_like(xd)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    stats, sums = torch.zeros(G * 4 * C, device=dev), torch.zeros(G * 2 * C, device=dev)
    ws = torch.empty(ops.bn_workspace_floats(G, Pg, C), device=dev)
    dgamma, dbeta = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    cnt = ops.bn_counters(dev)
    ops.bn_forward(xd, out, gamma.to(dev), beta.to(dev), rm, rv, nbt, stats, ws, cnt, G, Pg, C, act, slope)
    ops.bn_backward(nhwc(dout).to(dev), xd, stats, dx, dgamma, dbeta, sums, ws, cnt, G, Pg, C, act, slope)
    assert int(cnt.abs().sum().item()) == 0
    assert relerr(nchw(out), torch.cat(outs)) < FP32_TOL
    assert relerr(nchw(dx), torch.cat(dxs)) < 1e-4
    assert relerr(rm, bn.running_mean) < FP32_TOL and relerr(rv, bn.running_var) < FP32_TOL
    assert int(nbt.item()) == G == int(bn.num_batches_tracked.item())
    assert relerr(dgamma, bn.weight.grad) < 1e-4 and relerr(dbeta, bn.bias.grad) < 1e-4


@pytest.mark.parametrize("C,k", [(256, 4), (512, 4), (128, 7)])
def test_head_bce(ops, dev, C, k):
    torch.manual_seed(9)
    G, b = 2, 8
    a = torch.randn(G * b, C, k, k) * 0.3
    w = torch.randn(1, C, k, k) * 0.05
    a[0] *= 40  # saturate one sample: exercises the log clamp and the 1e-12 guard
    labels = torch.tensor([1.0, 0.0])
    ad = a.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    p = torch.sigmoid(F.conv2d(ad, wd)).view(-1)
    losses = [F.binary_cross_entropy(p[g * b:(g + 1) * b], torch.full((b,), labels[g].item(), dtype=torch.double))
              for g in range(G)]
    sum(losses).backward()
    prob, terms, dlogit = (torch.zeros(G * b, device=dev) for _ in range(3))
    loss = torch.zeros(G + 1, device=dev)
    a_d = nhwc(a).to(dev)
    wt = ops.head_pack(w.to(dev), torch.empty(k * k * C, device=dev))
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    for _ in range(2):  # twice: the fused loss reduction must leave its block counter cleared
        loss.zero_()
        ops.head_forward(a_d, wt, labels.to(dev), prob, terms, dlogit, loss, counter, G, b, k * k, C)
    assert int(counter.item()) == 0
    da, dw = torch.empty_like(a_d), torch.empty_like(w, device=dev)
    ops.head_backward(a_d, wt, dlogit, da, dw, G * b, k * k, C)
    assert relerr(prob, p) < 1e-5
    assert relerr(loss[:G], torch.stack(losses)) < 1e-5 and relerr(loss[G], sum(losses)) < 1e-5
    assert relerr(nchw(da), ad.grad) < 1e-4 and relerr(dw, wd.grad) < 1e-4


@pytest.mark.parametrize("prec", [0, 1])
def test_pack_plan_matches_single_packs(ops, dev, prec):
    """One mdgan_pack_weights_multi launch == the individual pack launches, bit for bit (incl. the head re-order)."""
    torch.manual_seed(21)
    Wd, Wu = torch.randn(128, 64, 4, 4, device=dev), torch.randn(256, 128, 4, 4, device=dev)
    Wz, Wh = torch.randn(100, 96, 7, 7, device=dev), torch.randn(1, 128, 7, 7, device=dev)
    ref = [ops.pack_down(Wd, precision=prec), ops.pack_up(Wu, precision=prec), ops.pack_dense(Wz, precision=prec),
           ops.head_pack(Wh, torch.empty(49 * 128, device=dev))]
    outs = [torch.full_like(r, float("nan")) for r in ref]
    plan = ops.PackPlan(dev)
    plan.add_down(Wd, outs[0])
    plan.add_up(Wu, outs[1])
    plan.add_dense(Wz, outs[2])
    plan.add_head(Wh, outs[3])
    plan.finalize().run()
    for o, r in zip(outs, ref):
        assert torch.equal(o, r)


def test_adam_matches_torch(ops, dev):
    torch.manual_seed(10)
    n = 10007   # not a multiple of 4: float4 body + scalar tail
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-4, betas=(0.5, 0.999))
    p, m, v = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(2, dtype=torch.int32, device=dev)   # (steps taken, block counter)
    for it in range(5):
        g = torch.randn(n) * (10.0 ** (-it))
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(dev), m, v, step, 2e-4, 0.5, 0.999)
    assert step.tolist() == [5, 0]
    assert (p.cpu() - ref.data).abs().max().item() < 2e-7


def test_tanh_backward_and_sum_slices(ops, dev):
    torch.manual_seed(11)
    s, x = torch.randn(4, 3, 8, 8), torch.tanh(torch.randn(4, 3, 8, 8))
    out = torch.empty_like(s, device=dev)
    ops.tanh_backward(s.to(dev), x.to(dev), out, 0.125)
    assert relerr(out, s * (1 - x * x) * 0.125) < 1e-6
    f = torch.randn(6, 5, 7)
    acc = torch.empty(2, 5, 7, device=dev)  # slots n % 2 of 6 workers, k = 2
    fd = f.to(dev)
    for slot in range(2):
        ops.sum_slices(fd[slot:], acc[slot], 3, 2 * 35)
    assert relerr(acc[0], f[0] + f[2] + f[4]) < 1e-6 and relerr(acc[1], f[1] + f[3] + f[5]) < 1e-6


# ----------------------------------------------------------------------------------------------- fused BatchNorm
def _bn_ref(z64, gamma, beta, G, slope):
    """train-mode BatchNorm2d per pass + LeakyReLU(slope) (slope 0 = ReLU) in fp64; z64 NCHW with G passes along dim 0."""
    outs = []
    for zc in z64.chunk(G):
        outs.append(F.leaky_relu(F.batch_norm(zc, None, None, gamma.double(), beta.double(), True, 0.1, 1e-5), slope))
    return torch.cat(outs)


@pytest.mark.parametrize("mode,n,C,N,H,G", [("down", 16, 64, 128, 16, 2), ("down", 8, 128, 256, 8, 1), ("up", 16, 128, 64, 8, 1),
                                            ("up", 16, 256, 128, 4, 1), ("up", 6, 256, 128, 7, 1), ("up", 32, 128, 64, 14, 1)])
def test_conv_fused_bn_statistics(ops, dev, mode, n, C, N, H, G):
    """conv_gemm(bn_partial) + bn_finalize + bn_apply == conv -> train-mode BatchNorm -> activation, including the
    running statistics, for DOWN (two passes), UP and the paired-parity UP kernel (64 channels)."""
    torch.manual_seed(21)
    x = torch.randn(n, C, H, H)
    gamma, beta = torch.rand(N) + 0.5, torch.randn(N) * 0.1
    if mode == "down":
        W = torch.randn(N, C, 4, 4) * 0.05
        z64 = F.conv2d(x.double(), W.double(), stride=2, padding=1)
        wp, m, grid, Ho = ops.pack_down(W.to(dev), precision=1), ops.MODE_DOWN, (n, H // 2, H // 2), H // 2
    else:
        W = torch.randn(C, N, 4, 4) * 0.05
        z64 = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1)
        wp, m, grid, Ho = ops.pack_up(W.to(dev), precision=1), ops.MODE_UP, (n, H, H), 2 * H
    ref = _bn_ref(z64, gamma, beta, G, 0.2)
    plan = ops.conv_stats_plan(grid, m, G, 1, N)
    assert plan is not None
    row_tiles, tpg, phases = plan
    Np = ops.n_pad_for(N)
    part = torch.full((phases * row_tiles * 2 * Np,), float("nan"), device=dev)
    z = torch.empty(n, Ho, Ho, N, device=dev)
    a = torch.empty_like(z)
    ops.conv_gemm(nhwc(x).to(dev), wp, m, N, z, grid, (H, H), precision=1, bn_partial=part)
    stats = torch.zeros(G * 4 * N, device=dev)
    rm, rv = torch.zeros(N, device=dev), torch.ones(N, device=dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    Pg = (n // G) * Ho * Ho
    ops.bn_finalize(part, plan, Np, 1, gamma.to(dev), beta.to(dev), rm, rv, nbt, stats, G, Pg, N)
    ops.bn_apply(z, stats, a, G, Pg, N, ops.ACT_LRELU, 0.2)
    assert relerr(nchw(z), z64) < FP32_TOL
    assert relerr(nchw(a), ref) < 1e-4
    assert int(nbt) == G
    bn = torch.nn.BatchNorm2d(N).double()
    bn.train()
    for zc in z64.chunk(G):
        bn(zc)
    assert relerr(rm, bn.running_mean) < 1e-4 and relerr(rv, bn.running_var) < 1e-4


def test_conv_dense_fused_bn_statistics(ops, dev):
    """first generator layer: GEMM columns are (position, channel); the finalize folds the k*k positions."""
    torch.manual_seed(22)
    n, C, N, k = 128, 100, 256, 4
    zin, W = torch.randn(n, C, 1, 1), torch.randn(C, N, k, k) * 0.05
    gamma, beta = torch.rand(N) + 0.5, torch.randn(N) * 0.1
    z64 = F.conv_transpose2d(zin.double(), W.double())
    ref = _bn_ref(z64, gamma, beta, 1, 0.0)
    wp = ops.pack_dense(W.to(dev), precision=1)
    zp = torch.zeros(n, wp.shape[1], device=dev)
    ops.pad_rows(zin.view(n, C).to(dev), zp)
    plan = ops.conv_stats_plan((n, 1, 1), ops.MODE_DENSE, 1, 1, k * k * N)
    assert plan is not None
    part = torch.full((plan[2] * plan[0] * 2 * k * k * N,), float("nan"), device=dev)
    z = torch.empty(n, k, k, N, device=dev)
    a = torch.empty_like(z)
    ops.conv_gemm(zp, wp, ops.MODE_DENSE, k * k * N, z, (n, 1, 1), (1, 1), precision=1, bn_partial=part)
    stats = torch.zeros(4 * N, device=dev)
    ops.bn_finalize(part, plan, k * k * N, k * k, gamma.to(dev), beta.to(dev), None, None, None, stats, 1, n * k * k, N)
    ops.bn_apply(z, stats, a, 1, n * k * k, N, ops.ACT_RELU, 0.0)
    assert relerr(nchw(a), ref) < 1e-4


@pytest.mark.parametrize("mode,n,C,N,H,G", [("up", 16, 128, 64, 8, 2), ("up", 8, 256, 128, 4, 1), ("down", 16, 128, 256, 16, 1)])
def test_conv_fused_bn_backward(ops, dev, mode, n, C, N, H, G):
    """Data-gradient GEMM with the BatchNorm-backward reduction in its epilogue (bnb) + bn_bwd_finalize +
    bn_bwd_apply_dy == autograd through  z -> BatchNorm -> LeakyReLU -> [conv whose data gradient the GEMM is]."""
    torch.manual_seed(23)
    gamma, beta = torch.rand(N) + 0.5, torch.randn(N) * 0.1
    g = torch.randn(n, C, H, H)          # gradient arriving at the GEMM's source
    if mode == "up":    # data gradient of a Conv2d(N -> C): da = conv_transpose(g, W), W [C, N, 4, 4]
        W = torch.randn(C, N, 4, 4) * 0.05
        Ho = 2 * H
        wp, m, grid = ops.pack_up(W.to(dev), precision=1), ops.MODE_UP, (n, H, H)
        fwd = lambda a_: F.conv2d(a_, W.double(), stride=2, padding=1)
    else:               # data gradient of a ConvTranspose2d(N -> C): da = conv(g, W), W [N, C, 4, 4]
        W = torch.randn(N, C, 4, 4) * 0.05
        Ho = H // 2
        wp, m, grid = ops.pack_down(W.to(dev), precision=1), ops.MODE_DOWN, (n, Ho, Ho)
        fwd = lambda a_: F.conv_transpose2d(a_, W.double(), stride=2, padding=1)
    z = torch.randn(n, N, Ho, Ho)
    z64 = z.double().requires_grad_(True)
    gm, bt = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    outs = []
    for zc in z64.chunk(G):
        outs.append(F.leaky_relu(F.batch_norm(zc, None, None, gm, bt, True, 0.1, 1e-5), 0.2))
    fwd(torch.cat(outs)).backward(g.double())
    # forward statistics on the device (unfused kernels), then the fused backward
    Pg = (n // G) * Ho * Ho
    zd = nhwc(z).to(dev)
    stats = torch.zeros(G * 4 * N, device=dev)
    ws = torch.empty(ops.bn_workspace_floats(G, Pg, N), device=dev)
    cnt = ops.bn_counters(dev)
    ops.bn_forward(zd, torch.empty_like(zd), gamma.to(dev), beta.to(dev), None, None, None, stats, ws, cnt, G, Pg, N,
                   ops.ACT_LRELU, 0.2)
    plan = ops.conv_stats_plan(grid, m, G, 1, N)
    assert plan is not None
    Np = ops.n_pad_for(N)
    part = torch.full((plan[2] * plan[0] * 2 * Np,), float("nan"), device=dev)
    dy = torch.empty(n, Ho, Ho, N, device=dev)
    ops.conv_gemm(nhwc(g).to(dev), wp, m, N, dy, grid, (H, H), precision=1, bn_partial=part,
                  bnb=(zd, stats, ops.ACT_LRELU, 0.2, G))
    sums = torch.zeros(G * 2 * N, device=dev)
    dgamma, dbeta = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
    ops.bn_bwd_finalize(part, plan, Np, sums, dgamma, dbeta, G, N)
    dz = torch.empty_like(dy)
    ops.bn_bwd_apply_dy(dy, zd, stats, sums, dz, G, Pg, N)
    assert relerr(nchw(dz), z64.grad) < 1e-4
    assert relerr(dgamma, gm.grad) < 1e-4 and relerr(dbeta, bt.grad) < 1e-4
