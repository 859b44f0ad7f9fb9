import os, sys, torch, torch.nn.functional as F
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200"))
from mdgan_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def nhwc(x): return x.permute(0, 2, 3, 1).contiguous()
def nchw(x): return x.permute(0, 3, 1, 2).contiguous()
for (n, C, N, H, bn) in [(128, 256, 128, 4, 0), (128, 256, 128, 4, 64), (128, 256, 128, 4, 16), (16, 256, 128, 4, 32), (40, 256, 128, 4, 32), (128, 64, 128, 4, 32)]:
    x, W = torch.randn(n, C, H, H), torch.randn(C, N, 4, 4) * 0.05
    ref = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1)
    out = torch.zeros(n, 2 * H, 2 * H, N, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), ops.pack_up(W.to(dev), precision=1), ops.MODE_UP, N, out, (n, H, H), (H, H), precision=1, force_bn=bn)
    torch.cuda.synchronize()
    got = nchw(out).double().cpu()
    err = (got - ref).abs()
    print(f"n={n} C={C} N={N} H={H} bn={bn}: max relerr {err.max() / ref.abs().max():.3e}")
    # per image error
    per_img = err.amax(dim=(1, 2, 3)) / ref.abs().max()
    bad = (per_img > 1e-4).nonzero().flatten().tolist()
    print("   bad images:", bad[:40], "count", len(bad))
    per_ch = err.amax(dim=(0, 2, 3)) / ref.abs().max()
    badc = (per_ch > 1e-4).nonzero().flatten().tolist()
    print("   bad channels:", badc[:40], "count", len(badc))
    per_pos = err.amax(dim=(0, 1)) / ref.abs().max()
    print("   bad positions:", (per_pos > 1e-4).nonzero().tolist()[:20])
