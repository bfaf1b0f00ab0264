"""One Cnn14 forward (eval, batch 8, encoder's real input size) for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import Cnn14
dev = torch.device("cuda", 0)
m = Cnn14(num_classes=512).to(dev).eval()
x = torch.rand(8, 1, 1025, 257, device=dev) ** 3
with torch.no_grad():
    for _ in range(2):
        y = m(x)
torch.cuda.synchronize()
print("done")
