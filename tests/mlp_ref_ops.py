"""Plain-torch statement of what the MLP kernels of csrc/mlp.cu (and the shared Adam / tanh-backward kernels) compute,
argument for argument with the wrappers in mdgan_b200/ops.py.  Used two ways: the GPU kernel tests compare the CUDA
kernels with these functions evaluated in fp64, and the CPU host-logic test patches them in place of the kernels to
run mlp_nets.MlpGenNet / MlpDiscNet + engine.MDGANEngine against the oracle without a GPU (mask order, RNG plumbing,
layer wiring).  Test infrastructure only."""
import torch

ACT_NONE, ACT_LRELU, ACT_TANH = 0, 2, 3


def _tail(v, bias, act, slope, mask, mask_scale, gate, gate_slope):
    if bias is not None:
        v = v + bias.to(v.dtype)
    if act == ACT_LRELU:
        v = torch.where(v > 0, v, v * slope)
    elif act == ACT_TANH:
        v = torch.tanh(v)
    if mask is not None:
        v = torch.where(mask.bool(), v * torch.tensor(mask_scale, dtype=v.dtype, device=v.device), torch.zeros_like(v))
    if gate is not None:
        v = torch.where(gate > 0, v, v * gate_slope)
    return v


def linear_forward(x, W, bias, out, act=ACT_NONE, slope=0.0, mask=None, mask_scale=1.0):
    out.copy_(_tail(x.to(out.dtype) @ W.to(out.dtype).t(), bias, act, slope, mask, mask_scale, None, 1.0))
    return out


def linear_dgrad(dy, W, out, gate=None, gate_slope=1.0, mask=None, mask_scale=1.0, accumulate=False):
    v = _tail(dy.to(out.dtype) @ W.to(out.dtype), None, ACT_NONE, 0.0, mask, mask_scale,
              gate.to(out.dtype) if gate is not None else None, gate_slope)
    out.copy_(out + v if accumulate else v)
    return out


def linear_wgrad(dy, x, dW):
    dW.copy_(dy.to(dW.dtype).t() @ x.to(dW.dtype))
    return dW


def col_sum(x, out):
    out.copy_(x.to(out.dtype).sum(0))
    return out


def linear_head_forward(a, w, bias, label, prob, loss_terms, dlogit, loss, counter, G, b):
    n = G * b
    logit = a[:n].to(prob.dtype) @ w.to(prob.dtype)
    if bias is not None:
        logit = logit + bias.to(prob.dtype)[0]
    y = label[:G].to(prob.dtype).repeat_interleave(b)
    p = torch.sigmoid(logit)
    lp, l1p = torch.log(p).clamp_min(-100.0), torch.log1p(-p).clamp_min(-100.0)
    terms = (y - 1) * l1p - y * lp
    pq = (1 - p) * p
    prob[:n].copy_(p)
    loss_terms[:n].copy_(terms)
    dlogit[:n].copy_(((p - y) / pq.clamp_min(1e-12)) * (1.0 / b) * pq)
    per = terms.view(G, b).mean(1)
    loss[:G].copy_(per)
    loss[G] = per.sum()


def linear_head_backward(a, w, dlogit, da, dw, dbias, mask=None, mask_scale=1.0, gate_slope=1.0):
    n = a.shape[0]
    d = dlogit[:n].to(da.dtype)
    v = d[:, None] * w.to(da.dtype)[None, :]
    da.copy_(_tail(v, None, ACT_NONE, 0.0, mask, mask_scale, a.to(da.dtype), gate_slope))
    if dw is not None:
        dw.copy_(d @ a.to(da.dtype))
    if dbias is not None:
        dbias.copy_(d.sum().reshape(1))


def tanh_backward(s, x, out, scale):
    out.copy_(s * (1 - x * x) * scale)
    return out


def tanh_backward_slices(F, x, out, k, N, scale):
    n_per = x.numel() // k
    Fv = F.reshape(N, n_per)
    acc = torch.stack([sum(Fv[n] for n in range(s, N, k)) for s in range(k)]).reshape(x.shape)
    out.copy_(acc * (1 - x * x) * scale)
    return out


def adam_step(p, g, m, v, step_count, lr, beta1, beta2, eps=1e-8):
    """elementwise.cu adam_kernel = torch.optim.Adam's single-tensor update."""
    t = int(step_count[0]) + 1
    bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
    m.add_((g - m) * (1.0 - beta1))
    v.mul_(beta2).add_((1.0 - beta2) * g * g)
    p.sub_((lr / bc1) * (m / (v.sqrt() / (bc2 ** 0.5) + eps)))
    step_count[0] = t


ALL = ("linear_forward", "linear_dgrad", "linear_wgrad", "col_sum", "linear_head_forward", "linear_head_backward",
       "tanh_backward", "tanh_backward_slices", "adam_step")


def patch(monkeypatch, ops_module) -> None:
    for name in ALL:
        monkeypatch.setattr(ops_module, name, globals()[name])
