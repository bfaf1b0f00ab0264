"""Two Cnn14 training steps (forward + backward, batch statistics, batch 4, encoder's real input size) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import Cnn14
dev = torch.device("cuda", 0)
m = Cnn14(num_classes=512).to(dev).train()
x = torch.rand(4, 1, 1025, 257, device=dev) ** 3
for _ in range(2):
    for p in m.parameters():
        p.grad = None
    m(x).square().mean().backward()
torch.cuda.synchronize()
print("done")
