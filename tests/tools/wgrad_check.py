"""TEST INFRASTRUCTURE: tensor-core weight gradient (dmst_conv3x3_wgrad) against float64 products on small and
real layer shapes, plus timing against the split-K batched library GEMMs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from diffmst_b200 import conv

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def wgrad_ours(x_pad, gz, Cin, Cout, tc):
    conv._TC_WGRAD = tc
    w = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
    xp = x_pad.clone()
    z = conv._Conv3x3Function.apply(xp, w, True)
    z.backward(gz)
    conv._TC_WGRAD = True
    return w.grad


for (B, H, W, Cin, Cout) in ((1, 6, 5, 1, 64), (2, 33, 17, 1, 64), (4, 1025, 257, 1, 64), (1, 6, 5, 32, 32), (2, 17, 9, 64, 64), (1, 33, 20, 32, 128), (2, 8, 8, 128, 256), (4, 1025, 257, 64, 64),
                              (4, 512, 128, 128, 128), (4, 8, 8, 1024, 1024)):
    x = torch.randn(B, Cin, H, W, device=dev)
    g = torch.randn(B, Cout, H, W, device=dev)
    x_pad = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
    gz = F.pad(g.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
    got = wgrad_ours(x_pad, gz, Cin, Cout, True)
    lib_ = wgrad_ours(x_pad, gz, Cin, Cout, False)
    if B * H * W <= 20000:
        want = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, 3, 3), g.double(), padding=1)
    else:
        want = lib_.double()
    err = float((got.double() - want).abs().max() / want.abs().max())
    def t(tc, n=5):
        for _ in range(2): wgrad_ours(x_pad, gz, Cin, Cout, tc)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): wgrad_ours(x_pad, gz, Cin, Cout, tc)
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    print(f"B{B} {H}x{W} {Cin}->{Cout}: rel-max err {err:.2e}   (call incl. forward+dgrad) tensor-core {t(True):.3f} ms, library {t(False):.3f} ms", flush=True)
