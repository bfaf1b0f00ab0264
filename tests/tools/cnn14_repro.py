"""TEST INFRASTRUCTURE: repeat the Cnn14 forward + backward many times and report any run whose output or
gradients differ bitwise from the first (the kernels are deterministic: any difference is a race)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import Cnn14
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = sys.argv[2] if len(sys.argv) > 2 else "eval"
g = torch.Generator().manual_seed(3)
m = Cnn14(num_classes=32).cuda()
m = m.eval() if mode == "eval" else m.train()
x = (torch.rand(2, 1, 1024, 128, generator=g) ** 2).cuda()
first = None
bad = 0
for it in range(n):
    for p in m.parameters(): p.grad = None
    junk = torch.randn(int(torch.randint(1, 50, (1,))) * 1000003, device="cuda")  # perturb the allocator state
    out = m(x)
    out.square().mean().backward()
    cur = {"out": out.detach().clone(), **{k: p.grad.clone() for k, p in m.named_parameters()}}
    del junk
    if first is None:
        first = cur
        continue
    diff = [k for k in cur if not torch.equal(cur[k], first[k])]
    if diff:
        bad += 1
        print("iteration", it, "differs in", diff[:8], "max rel", max(float((cur[k] - first[k]).abs().max() / first[k].abs().max().clamp_min(1e-30)) for k in diff))
print(mode, "iterations", n, "runs that differ from the first:", bad)
