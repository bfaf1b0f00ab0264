"""TEST INFRASTRUCTURE: probes of dmst_conv3x3_wgrad with impulse inputs."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import _lib
from diffmst_b200.conv import _ptr, _stream
lib = _lib.lib()
dev = torch.device("cuda", 0)
def run(x_pad, gz):
    B, Hp, Wp, Cin = x_pad.shape; Cout = gz.shape[-1]
    n = lib.dmst_conv3x3_wgrad_workspace_bytes(B, Hp - 2, Wp - 2, Cin, Cout)
    ws = torch.zeros(n, dtype=torch.uint8, device=dev)
    g9 = torch.full((9, Cout, Cin), -7.0, device=dev)
    rc = lib.dmst_conv3x3_wgrad(_ptr(x_pad), _ptr(gz), _ptr(g9), B, Hp - 2, Wp - 2, Cin, Cout, _ptr(ws), n, _stream(dev))
    torch.cuda.synchronize()
    return rc, g9, ws.view(torch.float32)
B, H, W, Cin, Cout = 1, 6, 5, 32, 32
Hp, Wp = H + 2, W + 2
x = torch.zeros(B, Hp, Wp, Cin, device=dev); gz = torch.zeros(B, Hp, Wp, Cout, device=dev)
x[0, 3, 3, 5] = 2.0      # pixel p = 3*7+3 = 24
gz[0, 3, 3, 9] = 3.0     # same pixel -> centre tap (4): g9[4][9][5] = 6
rc, g9, part = run(x, gz)
print("rc", rc, "nonzeros", torch.nonzero(g9).tolist()[:10], "values", g9[g9 != 0][:10].tolist(), "partial nnz", int((part != 0).sum()))
gz.zero_(); gz[0, 2, 3, 9] = 3.0   # dz at pixel above x's pixel: x[p + off] with off = +Wp -> tap 7 (ky=2,kx=1)
rc, g9, part = run(x, gz)
print("rc", rc, "nonzeros", torch.nonzero(g9).tolist()[:10], "values", g9[g9 != 0][:10].tolist())
x = torch.randn(B, Hp, Wp, Cin, device=dev); gz = torch.randn(B, Hp, Wp, Cout, device=dev)
gz[:, 0] = 0; gz[:, -1] = 0; gz[:, :, 0] = 0; gz[:, :, -1] = 0
rc, g9, part = run(x, gz)
P = B * Hp * Wp
xf, gf = x.view(P, Cin).double(), gz.view(P, Cout).double()
want = torch.zeros(9, Cout, Cin, dtype=torch.float64, device=dev)
for t in range(9):
    off = (t // 3 - 1) * Wp + (t % 3 - 1)
    lo, hi = max(0, -off), min(P, P - off)
    want[t] = gf[lo:hi].t() @ xf[lo + off:hi + off]
print("random: max|got|", float(g9.abs().max()), "max|want|", float(want.abs().max()), "rel err", float((g9.double() - want).abs().max() / want.abs().max()))
for t in range(9):
    print(" tap", t, "err", float((g9[t].double() - want[t]).abs().max() / want.abs().max()), "errT", float((g9[t].double() - want[t].t()).abs().max() / want.abs().max()) if Cin == Cout else "")
