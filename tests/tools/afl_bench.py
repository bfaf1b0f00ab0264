"""TEST INFRASTRUCTURE (uses the oracle as the comparison arm): AudioFeatureLoss forward+backward: ours vs the oracle's PyTorch composition run on the same GPU
(the reference's own GPU path is that composition of torch ops, mst/loss.py:62-260)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import AudioFeatureLoss
from oracle.loss import OracleAudioFeatureLoss
B, T = 8, 262144
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, 2, T, generator=g) * 0.1).cuda().requires_grad_(True)
y = (torch.randn(B, 2, T, generator=g) * 0.1).cuda()
w = [0.1, 0.001, 1.0, 1.0, 0.1]
ours = AudioFeatureLoss(w, 44100)
ref = OracleAudioFeatureLoss(w, 44100)
def run(f):
    x.grad = None
    d = f(x, y)
    sum(v for v in d.values()).backward()
def timeit(f, n=10):
    for _ in range(3): run(f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run(f)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"AudioFeatureLoss fwd+bwd B={B} T={T}: ours {timeit(ours):.3f} ms", flush=True)
try:
    print(f"  oracle torch composition on the GPU (float32): {timeit(ref):.3f} ms")
except Exception as e:
    print("  oracle on GPU failed:", repr(e)[:200])
