"""TEST INFRASTRUCTURE: run-to-run and float64 comparison of the first-layer weight gradient after the whole Cnn14
(the quantity tests/test_conv_gpu.py::test_cnn14_backward_small checks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import Cnn14
from oracle.panns import OracleCnn14
g = torch.Generator().manual_seed(3)
ref = OracleCnn14(num_classes=32).cuda().eval()
ours = Cnn14(num_classes=32).cuda().eval()
ours.load_state_dict(ref.state_dict(), strict=True)
ref64 = OracleCnn14(num_classes=32).cuda().double().eval()
ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
x = (torch.rand(2, 1, 1024, 128, generator=g) ** 2).cuda()
def grad(m, inp):
    for p in m.parameters(): p.grad = None
    m(inp).square().mean().backward()
    return m.conv_block1.conv1.weight.grad.flatten().double().clone(), m.conv_block6.conv2.weight.grad.flatten().double().clone()
cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a, b, dim=0))
o = [grad(ours, x) for _ in range(4)]
r = [grad(ref, x) for _ in range(4)]
t = grad(ref64, x.double())
print("ours run-to-run (first layer):", [cos(o[0][0], o[i][0]) for i in range(1, 4)], "bitwise equal:", [bool(torch.equal(o[0][0], o[i][0])) for i in range(1, 4)])
print("ref  run-to-run (first layer):", [cos(r[0][0], r[i][0]) for i in range(1, 4)])
print("ours vs float64 (first layer):", [cos(o[i][0], t[0]) for i in range(4)], " last layer:", cos(o[0][1], t[1]))
print("ref  vs float64 (first layer):", [cos(r[i][0], t[0]) for i in range(4)], " last layer:", cos(r[0][1], t[1]))
print("ours vs ref     (first layer):", [cos(o[i][0], r[i][0]) for i in range(4)])
