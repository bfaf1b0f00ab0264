"""A few MRSTFT loss evaluations (forward + gradient) at the headline shape, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffmst_b200 import MRSTFTLoss
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
x = (torch.randn(bench.B, 2, bench.T) * 0.1).cuda().requires_grad_(True)
y = (torch.randn(bench.B, 2, bench.T) * 0.1).cuda()
f = MRSTFTLoss(**bench.RES)
for _ in range(reps):
    x.grad = None
    f(x, y).backward()
torch.cuda.synchronize()
print("done")
