"""Debug aid: _Conv3x3Function gradients vs torch's conv2d on one layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from diffmst_b200.conv import _Conv3x3Function
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
for (B, Cin, Cout, H, W) in [(2, 64, 64, 9, 7), (2, 64, 128, 21, 18), (1, 8, 8, 5, 5)]:
    x = torch.randn(B, Cin, H, W, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.1
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    probe = torch.randn_like(yr)
    (yr * probe).sum().backward()
    xo = x.clone().requires_grad_(True); wo = w.clone().requires_grad_(True)
    xp = F.pad(xo.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
    z = _Conv3x3Function.apply(xp, wo)
    yo = z[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
    (yo * probe).sum().backward()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(B, Cin, Cout, H, W, "fwd", rel(yo, yr), "gx", rel(xo.grad, xr.grad), "gw", rel(wo.grad, wr.grad))
    e = (xo.grad - xr.grad).abs().amax(dim=(0, 1))
    print((e / xr.grad.abs().max() > 1e-2).int())
