"""TEST INFRASTRUCTURE: dumps of the wgrad kernel's shared-memory tiles and accumulator (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import _lib
from diffmst_b200.conv import _ptr, _stream
lib = _lib.lib()
dev = torch.device("cuda", 0)
B, H, W, Cin, Cout = 1, 6, 5, 32, 32
Hp, Wp = H + 2, W + 2
P = B * Hp * Wp
torch.manual_seed(1)
x = torch.randn(B, Hp, Wp, Cin, device=dev); gz = torch.randn(B, Hp, Wp, Cout, device=dev)
gz[:, 0] = 0; gz[:, -1] = 0; gz[:, :, 0] = 0; gz[:, :, -1] = 0
n = lib.dmst_conv3x3_wgrad_workspace_bytes(B, H, W, Cin, Cout)
ws = torch.zeros(n, dtype=torch.uint8, device=dev)
g9 = torch.full((9 * Cout * Cin,), -7.0, device=dev)
rc = lib.dmst_conv3x3_wgrad(_ptr(x), _ptr(gz), _ptr(g9), B, H, W, Cin, Cout, _ptr(ws), n, _stream(dev))
torch.cuda.synchronize()
mode = int(os.environ.get("DMST_WG_DEBUG", "0"))
print("mode", mode, "rc", rc)
xf, gf = x.view(P, Cin), gz.view(P, Cout)
if mode & 4:
    sa, sb, acc = g9[:1024].view(32, 32), g9[1024:2048].view(32, 32), g9[2048:2048 + 4096].view(128, 32)
    # un-swizzle: 32-byte chunk index ^= (row & 3)
    def unsw(t):
        o = torch.empty_like(t)
        for r in range(32):
            for c in range(4):
                o[r, 8 * c:8 * c + 8] = t[r, 8 * (c ^ (r & 3)):8 * (c ^ (r & 3)) + 8]
        return o
    print("dz slab matches rows 0..31 of dz:", float((unsw(sa) - gf[0:32, 0:32]).abs().max()))
    print("x tap4 matches rows 0..31 of x:", float((unsw(sb) - xf[0:32, 0:32]).abs().max()))
    want = gf[:, :32].double().t() @ xf[:, :32].double()
    print("acc max", float(acc.abs().max()), "acc[:32] vs want", float((acc[:32].double() - want).abs().max() / want.abs().max()),
          "vs want.T", float((acc[:32].double() - want.t()).abs().max() / want.abs().max()))
    print(acc[:4, :6]); print(want[:4, :6])
else:
    w4 = gf.double().t() @ xf.double()
    print("tap4 err", float((g9.view(9, Cout, Cin)[4].double() - w4).abs().max() / w4.abs().max()), "max got", float(g9.abs().max()))
