"""Diagnostic: separate the console backward from the MRSTFT gradient at T=262144."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_headline_gpu as h
import test_console_gpu as c
from oracle.auraloss.freq import MultiResolutionSTFTLoss as OracleMRSTFT
from diffmst_b200 import MRSTFTLoss
np.set_printoptions(precision=4, suppress=True, linewidth=200)
T = 262144
tracks, tp, fp, mp, tp2, mp2 = h._inputs(1, 16, T, seed=2026)
g = torch.Generator().manual_seed(5)
probe = torch.randn(1, 2, T, generator=g)
flags = dict(use_fx_bus=False)
ours = c.run_ours(tracks, tp, fp, mp, probe, flags)
o64 = c.run_oracle(tracks, tp, fp, mp, probe, flags, torch.float64)
o32 = c.run_oracle(tracks, tp, fp, mp, probe, flags, torch.float32)
print("linear probe, T=262144: mix relmax ours/f32", c.relmax(ours["mix"], o64["mix"]), c.relmax(o32["mix"], o64["mix"]))
print("  gtp rell2 ours/f32", c.rell2(ours["gtp"], o64["gtp"]), c.rell2(o32["gtp"], o64["gtp"]))
print("  gmp rell2 ours/f32", c.rell2(ours["gmp"], o64["gmp"]), c.rell2(o32["gmp"], o64["gmp"]))
print("  gmp f64 ", o64["gmp"][0]); print("  gmp ours", ours["gmp"][0]); print("  gmp f32 ", o32["gmp"][0])
# MRSTFT gradient alone: same inputs
x64 = torch.from_numpy(o64["mix"]); y64 = torch.randn(1, 2, T, generator=g, dtype=torch.float64) * 0.1
def og(dtype):
    x = x64.to(dtype).requires_grad_(True)
    l = OracleMRSTFT(**h.RES)(x, y64.to(dtype)); l.backward(); return float(l), x.grad.double().numpy()
l64, g64 = og(torch.float64); l32, g32 = og(torch.float32)
xc = x64.float().cuda().requires_grad_(True)
lo = MRSTFTLoss(**h.RES)(xc, y64.float().cuda()); lo.backward()
go = xc.grad.double().cpu().numpy()
print("MRSTFT alone: loss", float(lo), l64, l32)
print("  dL/dx rell2 ours/f32:", c.rell2(go, g64), c.rell2(g32, g64), " relmax:", c.relmax(go, g64), c.relmax(g32, g64))
# sensitivity: perturb x by 1e-6 relative-to-max noise in f64 and see how much the f64 gradient moves
xp = x64 + torch.randn(x64.shape, generator=g, dtype=torch.float64) * 1e-6 * x64.abs().max()
xp = xp.requires_grad_(True); lp = OracleMRSTFT(**h.RES)(xp, y64); lp.backward()
print("  f64 gradient moved by a 1e-6 (of max) input perturbation: rell2", c.rell2(xp.grad.numpy(), g64))
