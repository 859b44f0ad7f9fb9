"""Layer-plan extraction for the datasets.<NAME> plugin models.

The actor API hands the engine ready-made `nn.Module`s (bootstrap.py builds `Generator()` / `Discriminator()` from
the plugin, /root/reference/src/bootstrap.py:75-76,102-103).  Plugins are written both in module style
(CIFAR10.py: nn.Sequential) and functional style (CelebA.py:95-101,134-142: F.leaky_relu / torch.tanh in
forward), so instead of walking `children()` the plan is recorded from one dry-run forward on a tiny CPU batch
under a TorchFunctionMode: every torch-level op (conv2d, conv_transpose2d, batch_norm, relu, leaky_relu, tanh,
sigmoid, view/squeeze/flatten) is captured with the *parameter tensors it received*, which are mapped back to
`state_dict` keys by identity.  Anything else (unknown strides, unknown ops) raises: there is no fallback path,
unsupported models are refused loudly.  `extract_mlp_plan` does the same for the reference's second model family, the
Linear / LeakyReLU / dropout MLP of datasets/MNIST.py:74-120.

The dry run restores the module's buffers (BatchNorm running stats / num_batches_tracked) and the global RNG
state afterwards, so it is invisible to the training run.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.overrides import TorchFunctionMode

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = "none", "relu", "lrelu", "tanh", "sigmoid"


class UnsupportedModelError(NotImplementedError):
    pass


@dataclass
class BNSpec:
    weight: str
    bias: str
    running_mean: str
    running_var: str
    num_batches_tracked: Optional[str]
    eps: float
    momentum: float


@dataclass
class ConvLayer:
    """One conv / conv-transpose with its (optional) BatchNorm and activation."""
    kind: str                  # "down" (Conv2d k4 s2 p1), "up" (ConvT k4 s2 p1), "dense_up" (ConvT k,s1,p0 on 1x1),
                               # "head" (Conv2d k,s1,p0 on a kxk map -> 1 channel)
    weight: str                # state_dict key
    bias: Optional[str]
    c_in: int
    c_out: int
    k: int
    h_in: int
    h_out: int
    bn: Optional[BNSpec] = None
    act: str = ACT_NONE
    slope: float = 0.0


@dataclass
class NetPlan:
    role: str                  # "generator" | "discriminator"
    layers: List[ConvLayer] = field(default_factory=list)
    in_shape: Tuple[int, ...] = ()
    out_shape: Tuple[int, ...] = ()


class _Recorder(TorchFunctionMode):
    def __init__(self):
        super().__init__()
        self.depth = 0
        self.ops: List[Tuple[str, tuple, dict, object]] = []

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", str(func))
        if self.depth > 0:
            return func(*args, **kwargs)
        self.depth += 1
        try:
            out = func(*args, **kwargs)
        finally:
            self.depth -= 1
        self.ops.append((name, args, kwargs, out))
        return out


_IGNORED = {"view", "squeeze", "flatten", "reshape", "contiguous", "size", "dim", "__get__", "to", "detach",
            "is_cuda", "shape", "__getitem__", "unsqueeze", "add_"}


def _param_names(module: nn.Module) -> Dict[int, str]:
    names: Dict[int, str] = {}
    for k, v in module.state_dict(keep_vars=True).items():
        names[id(v)] = k
    return names


def _arg(args, kwargs, pos, key, default=None):
    if len(args) > pos:
        return args[pos]
    return kwargs.get(key, default)


def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1] if len(v) > 1 else v[0])
    return int(v), int(v)


def record_ops(module: nn.Module, example: torch.Tensor):
    saved = copy.deepcopy({k: v.detach().clone() for k, v in module.state_dict().items()})
    rng = torch.get_rng_state()
    was_training = module.training
    try:
        with torch.no_grad(), _Recorder() as rec:
            out = module(example)
    finally:
        module.load_state_dict(saved)
        module.train(was_training)
        torch.set_rng_state(rng)
    return rec.ops, out


def extract_plan(module: nn.Module, role: str, in_shape: Tuple[int, ...]) -> NetPlan:
    """role "generator": in_shape = (z_dim, 1, 1); role "discriminator": in_shape = (C, H, W)."""
    if any(p.device.type != "cpu" for p in module.parameters()):
        raise UnsupportedModelError("extract_plan expects the module on the CPU (the engine owns the device copy)")
    names = _param_names(module)
    example = torch.zeros((2, *in_shape), dtype=torch.float32)
    ops, out = record_ops(module, example)
    plan = NetPlan(role=role, in_shape=tuple(in_shape), out_shape=tuple(out.shape[1:]))
    cur: Optional[ConvLayer] = None

    def key_of(t, what):
        if t is None:
            return None
        k = names.get(id(t))
        if k is None:
            raise UnsupportedModelError(f"{what}: tensor is not a parameter/buffer of the module")
        return k

    for name, args, kwargs, res in ops:
        if name in _IGNORED:
            continue
        if name in ("conv2d", "conv_transpose2d"):
            x, w = args[0], args[1]
            b = _arg(args, kwargs, 2, "bias")
            stride = _pair(_arg(args, kwargs, 3, "stride", 1))
            padding = _pair(_arg(args, kwargs, 4, "padding", 0))
            if name == "conv2d":
                dilation = _pair(_arg(args, kwargs, 5, "dilation", 1))
                groups = _arg(args, kwargs, 6, "groups", 1)
                out_pad = (0, 0)
            else:
                out_pad = _pair(_arg(args, kwargs, 5, "output_padding", 0))
                groups = _arg(args, kwargs, 6, "groups", 1)
                dilation = _pair(_arg(args, kwargs, 7, "dilation", 1))
            kh, kw = int(w.shape[2]), int(w.shape[3])
            if groups != 1 or dilation != (1, 1) or out_pad != (0, 0) or kh != kw or x.shape[2] != x.shape[3]:
                raise UnsupportedModelError(f"{name}: groups/dilation/output_padding/non-square not supported")
            h_in, h_out = int(x.shape[2]), int(res.shape[2])
            if name == "conv2d":
                c_out, c_in = int(w.shape[0]), int(w.shape[1])
                if (kh, stride, padding) == (4, (2, 2), (1, 1)):
                    kind = "down"
                elif stride == (1, 1) and padding == (0, 0) and h_in == kh and c_out == 1:
                    kind = "head"
                else:
                    raise UnsupportedModelError(f"conv2d k={kh} stride={stride} padding={padding} on {h_in}x{h_in}")
            else:
                c_in, c_out = int(w.shape[0]), int(w.shape[1])
                if (kh, stride, padding) == (4, (2, 2), (1, 1)):
                    kind = "up"
                elif stride == (1, 1) and padding == (0, 0) and h_in == 1:
                    kind = "dense_up"
                else:
                    raise UnsupportedModelError(f"conv_transpose2d k={kh} stride={stride} padding={padding}")
            cur = ConvLayer(kind=kind, weight=key_of(w, name), bias=key_of(b, name + ".bias"), c_in=c_in, c_out=c_out,
                            k=kh, h_in=h_in, h_out=h_out)
            plan.layers.append(cur)
        elif name == "batch_norm":
            if cur is None or cur.bn is not None or cur.act != ACT_NONE:
                raise UnsupportedModelError("batch_norm must directly follow a convolution")
            rm, rv = args[1], args[2]
            w, b = _arg(args, kwargs, 3, "weight"), _arg(args, kwargs, 4, "bias")
            training = _arg(args, kwargs, 5, "training", False)
            momentum = _arg(args, kwargs, 6, "momentum", 0.1)
            eps = _arg(args, kwargs, 7, "eps", 1e-5)
            if not training or w is None or b is None or rm is None:
                raise UnsupportedModelError("only affine train-mode BatchNorm2d with running stats is supported")
            rm_key = key_of(rm, "bn.running_mean")
            nbt_key = rm_key.replace("running_mean", "num_batches_tracked")
            if nbt_key not in module.state_dict():
                nbt_key = None
            cur.bn = BNSpec(weight=key_of(w, "bn.weight"), bias=key_of(b, "bn.bias"), running_mean=rm_key,
                            running_var=key_of(rv, "bn.running_var"), num_batches_tracked=nbt_key, eps=float(eps),
                            momentum=float(momentum if momentum is not None else 0.1))
        elif name in ("relu", "relu_", "leaky_relu", "leaky_relu_", "tanh", "sigmoid"):
            if cur is None or cur.act != ACT_NONE:
                raise UnsupportedModelError(f"{name}: activation must follow a convolution (+ BatchNorm)")
            if name.startswith("relu"):
                cur.act = ACT_RELU
            elif name.startswith("leaky_relu"):
                cur.act, cur.slope = ACT_LRELU, float(_arg(args, kwargs, 1, "negative_slope", 0.01))
            elif name == "tanh":
                cur.act = ACT_TANH
            else:
                cur.act = ACT_SIGMOID
        else:
            raise UnsupportedModelError(
                f"op `{name}` is not supported by the B200 MD-GAN engine (DCGAN-style conv models only; "
                "no fallback path exists)")
    _validate(plan)
    return plan


def _validate(plan: NetPlan) -> None:
    L = plan.layers
    if not L:
        raise UnsupportedModelError("no convolution layers found")
    if plan.role == "generator":
        if L[0].kind != "dense_up" or any(l.kind != "up" for l in L[1:]) or len(L) < 2:
            raise UnsupportedModelError("generator must be ConvT(k,1,0) on 1x1 followed by ConvT(4,2,1) layers")
        if L[-1].act != ACT_TANH or L[-1].bn is not None or L[-1].c_out not in (1, 3):
            raise UnsupportedModelError("generator must end with ConvT -> tanh producing 1 or 3 channels")
        for l in L[:-1]:
            if l.bn is None or l.act != ACT_RELU or l.bias is not None:
                raise UnsupportedModelError("generator hidden layers must be ConvT(bias=False) -> BatchNorm -> ReLU")
        if L[-1].bias is not None:
            raise UnsupportedModelError("generator output layer must not have a bias")
    else:
        # >= 3 layers: the LeakyReLU backward of layer 0 is the fused epilogue of layer 1's data-gradient GEMM
        # (DiscNet.backward), so a Conv -> head model has no kernel that would write dz[0]
        if any(l.kind != "down" for l in L[:-1]) or L[-1].kind != "head" or len(L) < 3:
            raise UnsupportedModelError("discriminator must be >= 2 Conv(4,2,1) layers followed by a kxk valid-conv head")
        if L[0].bn is not None or L[0].act != ACT_LRELU or L[0].bias is not None or L[0].c_in not in (1, 3):
            raise UnsupportedModelError("first discriminator layer must be Conv(bias=False) -> LeakyReLU on 1/3 channels")
        for l in L[1:-1]:
            if l.bn is None or l.act != ACT_LRELU:
                raise UnsupportedModelError("discriminator hidden layers must be Conv -> BatchNorm -> LeakyReLU")
        if L[-1].act != ACT_SIGMOID or L[-1].bias is not None or L[-1].bn is not None:
            raise UnsupportedModelError("discriminator head must be Conv(bias=False) -> sigmoid")


# ------------------------------------------------------------------------------------------------ MLP family
@dataclass
class LinearLayer:
    """One nn.Linear with the elementwise tail the reference applies to it (MNIST.py:86-96,114-120)."""
    weight: str                # state_dict keys
    bias: Optional[str]
    n_in: int
    n_out: int
    act: str = ACT_NONE        # lrelu | tanh | sigmoid | none
    slope: float = 0.0
    drop_p: float = 0.0        # F.dropout(p, training=True) after the activation (always active in the reference)


@dataclass
class MlpPlan:
    role: str
    layers: List[LinearLayer] = field(default_factory=list)
    in_shape: Tuple[int, ...] = ()
    out_shape: Tuple[int, ...] = ()


def is_mlp(module: nn.Module) -> bool:
    """True for models made of nn.Linear layers only (no convolution): the reference's MNIST plugin."""
    mods = list(module.modules())
    return any(isinstance(m, nn.Linear) for m in mods) and not any(
        isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d)) for m in mods)


def extract_mlp_plan(module: nn.Module, role: str, in_shape: Tuple[int, ...]) -> MlpPlan:
    """role "generator": in_shape = (z_dim, 1, 1), output reshaped to the image; role "discriminator": in_shape =
    (C, H, W), flattened, one sigmoid output per sample."""
    if any(p.device.type != "cpu" for p in module.parameters()):
        raise UnsupportedModelError("extract_mlp_plan expects the module on the CPU (the engine owns the device copy)")
    names = _param_names(module)
    ops, out = record_ops(module, torch.zeros((2, *in_shape), dtype=torch.float32))
    plan = MlpPlan(role=role, in_shape=tuple(in_shape), out_shape=tuple(out.shape[1:]))
    cur: Optional[LinearLayer] = None
    for name, args, kwargs, res in ops:
        if name in _IGNORED:
            continue
        if name == "linear":
            x, w = args[0], args[1]
            b = _arg(args, kwargs, 2, "bias")
            if x.dim() != 2 or names.get(id(w)) is None or (b is not None and names.get(id(b)) is None):
                raise UnsupportedModelError("linear: expected a 2-D input and module parameters")
            if cur is not None and cur.n_out != int(w.shape[1]):
                raise UnsupportedModelError("linear layers must be chained")
            cur = LinearLayer(weight=names[id(w)], bias=names[id(b)] if b is not None else None, n_in=int(w.shape[1]),
                              n_out=int(w.shape[0]))
            plan.layers.append(cur)
        elif name in ("leaky_relu", "leaky_relu_", "tanh", "sigmoid"):
            if cur is None or cur.act != ACT_NONE or cur.drop_p != 0.0:
                raise UnsupportedModelError(f"{name}: activation must directly follow a Linear layer")
            if name.startswith("leaky_relu"):
                cur.act, cur.slope = ACT_LRELU, float(_arg(args, kwargs, 1, "negative_slope", 0.01))
            else:
                cur.act = ACT_TANH if name == "tanh" else ACT_SIGMOID
        elif name == "dropout":
            p = float(_arg(args, kwargs, 1, "p", 0.5))
            training = bool(_arg(args, kwargs, 2, "training", True))
            if cur is None or cur.act != ACT_LRELU or cur.drop_p != 0.0 or not (0.0 < p < 1.0):
                raise UnsupportedModelError("dropout must follow Linear -> LeakyReLU with 0 < p < 1")
            if training:   # F.dropout(x, p) defaults to training=True whatever module.training says (MNIST.py:89)
                cur.drop_p = p
        else:
            raise UnsupportedModelError(f"op `{name}` is not supported by the B200 MD-GAN engine (MLP family: Linear, "
                                        "LeakyReLU, dropout, tanh / sigmoid outputs; no fallback path exists)")
    L = plan.layers
    if len(L) < 2:
        raise UnsupportedModelError("an MLP needs at least two Linear layers")
    n_in = 1
    for d in in_shape:
        n_in *= d
    if L[0].n_in != n_in:
        raise UnsupportedModelError("the first Linear layer must take the flattened input")
    for l in L[:-1]:
        if l.act != ACT_LRELU or l.slope <= 0.0:
            raise UnsupportedModelError("hidden layers must be Linear -> LeakyReLU(slope > 0) [-> dropout]")
    if role == "generator":
        n_out = 1
        for d in plan.out_shape:
            n_out *= d
        if L[-1].act != ACT_TANH or L[-1].n_out != n_out or any(l.drop_p for l in L):
            raise UnsupportedModelError("generator must end with Linear -> tanh reshaped to the image, without dropout")
    else:
        if L[-1].act != ACT_SIGMOID or L[-1].n_out != 1 or plan.out_shape != ():
            raise UnsupportedModelError("discriminator must end with Linear(., 1) -> sigmoid -> flatten")
    return plan
