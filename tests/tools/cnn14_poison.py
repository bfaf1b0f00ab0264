"""TEST INFRASTRUCTURE: stale-memory probe.  Fills the caching allocator's blocks with NaN (or a large constant)
before running Cnn14 forward + backward: any read of a buffer region that was allocated with torch.empty and never
written shows up as NaN / a changed checksum, and names the first tensor that is affected."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import Cnn14
mode = sys.argv[1] if len(sys.argv) > 1 else "eval"
g = torch.Generator().manual_seed(3)
m = Cnn14(num_classes=32).cuda()
m = m.eval() if mode == "eval" else m.train()
x = (torch.rand(2, 1, 1024, 128, generator=g) ** 2).cuda()
def run():
    for p in m.parameters(): p.grad = None
    out = m(x); out.square().mean().backward()
    torch.cuda.synchronize()
    return {"out": out.detach().double().clone(), **{k: p.grad.double().clone() for k, p in m.named_parameters()}}
def poison(val):
    blocks = [torch.full((64 * 1024 * 1024,), val, device="cuda") for _ in range(12)]   # 3 GiB of poison
    small = [torch.full((n,), val, device="cuda") for n in (256, 4096, 65536, 1 << 20, 1 << 22) for _ in range(8)]
    del blocks, small
torch.cuda.empty_cache(); poison(0.0); base = run()
for val in (float("nan"), 1e30, -3.0):
    torch.cuda.empty_cache(); poison(val)
    cur = run()
    bad = [k for k in cur if not torch.equal(torch.nan_to_num(cur[k], nan=12345.0), torch.nan_to_num(base[k], nan=12345.0))]
    print(mode, "poison", val, "-> tensors that changed:", bad[:12], "non-finite:", [k for k in cur if not torch.isfinite(cur[k]).all()][:12])
