"""Diagnostic (GPU): where does the Cnn14 gradient on the reference-golden case (tests/golden/panns.npz) leave
float64?  Arms, all on the same generator-free weights and input, cosine of every conv weight gradient to float64:
  f64      OracleCnn14 in float64 (must reproduce the golden first-layer gradient)
  cudnn    float32 modules, cuDNN TF32 convolutions (the reference's own numerics on this GPU)
  fp32     float32 modules, TF32 off
  emu_rz   float32 modules, conv operands TRUNCATED to TF32 (what a tensor core does with raw fp32 bits)
  emu_rn   float32 modules, conv operands rounded to nearest TF32 (cvt.rna)
  ours     diffmst_b200.Cnn14
Usage: python tests/tools/cnn14_tf32_diag.py"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from panns_fill import fill_state  # noqa: E402
from oracle.panns import OracleCnn14  # noqa: E402


def rz(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def rn(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)   # round half away (cvt.rna), magnitudes


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, q):
        ctx.q = q
        ctx.save_for_backward(x, w)
        return F.conv2d(q(x), q(w), padding=1)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        q = ctx.q
        gx = torch.nn.grad.conv2d_input(x.shape, q(w), q(gy), padding=1)
        gw = torch.nn.grad.conv2d_weight(q(x), w.shape, q(gy), padding=1)
        return gx, gw, None


def patch(net, q):
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.forward = (lambda mm: (lambda x: QConv.apply(x, mm.weight, q)))(m)
    return net


def cos(a, b):
    return float(F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def main():
    from diffmst_b200 import Cnn14
    d = np.load(os.path.join(ROOT, "tests", "golden", "panns.npz"))
    g = torch.Generator().manual_seed(42)
    x = (torch.rand(1, 1, 1024, 128, generator=g) ** 2).cuda()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def make(dtype=torch.float32):
        n = OracleCnn14(num_classes=6)
        fill_state(n)
        return n.cuda().to(dtype).eval()

    arms = {}
    n64 = make(torch.float64)
    o64 = n64(x.double())
    o64.square().mean().backward()
    arms["f64"] = (n64, o64)
    print("f64 first-layer grad vs golden cosine:", cos(n64.conv_block1.conv1.weight.grad, torch.from_numpy(d["cnn14_g_first"]).cuda()))
    torch.backends.cudnn.allow_tf32 = True
    n = make(); o = n(x); o.square().mean().backward(); arms["cudnn"] = (n, o)
    torch.backends.cudnn.allow_tf32 = False
    n = make(); o = n(x); o.square().mean().backward(); arms["fp32"] = (n, o)
    n = patch(make(), rz); o = n(x); o.square().mean().backward(); arms["emu_rz"] = (n, o)
    n = patch(make(), rn); o = n(x); o.square().mean().backward(); arms["emu_rn"] = (n, o)
    ours = Cnn14(num_classes=6).cuda().eval()
    fill_state(ours)
    o = ours(x); o.square().mean().backward(); arms["ours"] = (ours, o)

    names = [f"conv_block{i}.conv{j}" for i in range(1, 7) for j in (1, 2)] + ["fc"]
    print("out relmax vs f64:", {k: float((v[1].double() - o64).abs().max() / o64.abs().max()) for k, v in arms.items()})
    print(f"{'layer':22s}" + "".join(f"{k:>12s}" for k in arms if k != "f64") + "   (1 - cosine to float64 weight gradient)")
    p64 = dict(n64.named_parameters())
    for nm in names:
        row = f"{nm:22s}"
        for k, (net, _) in arms.items():
            if k == "f64":
                continue
            row += f"{1.0 - cos(dict(net.named_parameters())[nm + '.weight'].grad, p64[nm + '.weight'].grad):12.2e}"
        print(row)


if __name__ == "__main__":
    main()
