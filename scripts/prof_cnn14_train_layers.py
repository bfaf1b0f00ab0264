"""One Cnn14 training step at the DDP step's batch (40 items) for per-launch ncu timings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import Cnn14
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = Cnn14(num_classes=512).cuda().train()
x = (torch.rand(B, 1, 1025, 257) ** 2).cuda()
for _ in range(2):
    for p in net.parameters(): p.grad = None
    net(x).square().mean().backward()
torch.cuda.synchronize()
