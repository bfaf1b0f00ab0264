"""Kernel-time breakdown (torch.profiler) of one Cnn14 training step on our path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from diffmst_b200 import Cnn14
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
x = torch.rand(B, 1, 1025, 257, device=dev) ** 3
m = Cnn14(num_classes=512).to(dev).train()
def step():
    for p in m.parameters(): p.grad = None
    m(x).square().mean().backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
