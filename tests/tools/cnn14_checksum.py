"""TEST INFRASTRUCTURE: checksums of one Cnn14 forward + backward (eval mode), to compare ACROSS processes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import Cnn14
torch.manual_seed(0)
g = torch.Generator().manual_seed(3)
m = Cnn14(num_classes=32).cuda().eval()
x = (torch.rand(2, 1, 1024, 128, generator=g) ** 2).cuda()
out = m(x); out.square().mean().backward()
cs = lambda t: f"{float(t.double().sum()):.10e}/{float(t.double().abs().sum()):.10e}"
names = ["conv_block6.conv2.weight", "conv_block6.conv1.weight", "conv_block5.conv2.weight", "conv_block3.conv1.weight", "conv_block1.conv2.weight", "conv_block1.conv1.weight", "conv_block6.bn2.weight"]
p = dict(m.named_parameters())
print("out", cs(out.detach()), " ".join(f"{n.replace('conv_block', 'b').replace('.weight', '')}={cs(p[n].grad)}" for n in names))
