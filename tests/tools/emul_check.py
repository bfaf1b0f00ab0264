"""TEST INFRASTRUCTURE: runs the host-emulated kernels (tests/emul) against the float64 oracle on a
small case; for debugging kernel logic in the GPU-less container."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import numpy as np, torch
import harness
from oracle.console import OracleAdvancedMixConsole

B, N, T = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 2, 20000)))
torch.manual_seed(0)
tracks = torch.randn(B, N, T) * 0.1
tp, fp, mp = torch.rand(B, N, 27), torch.rand(B, 25), torch.rand(B, 26)
probe = torch.randn(B, 2, T)
con = harness.EmulConsole()
flags = harness.flags_from()
t0 = time.time()
mix, mixed, status = con.forward(tracks.numpy(), tp.numpy(), mp.numpy(), flags, want_mixed=True)
gtp, gmp, gtr = con.backward(probe.numpy())
print(f"emulated fwd+bwd in {time.time()-t0:.1f}s")
orc = OracleAdvancedMixConsole(44100)
tpd = tp.double().requires_grad_(True); mpd = mp.double().requires_grad_(True); trd = tracks.double().requires_grad_(True)
omixed, omix, _, _, _ = orc(trd, tpd, fp.double(), mpd, use_fx_bus=False)
(omix * probe.double()).sum().backward()
def rel(a, b):
    a = torch.as_tensor(a).double(); return float((a - b).abs().max() / b.abs().max())
print("mix   rel-max", rel(mix, omix.detach()))
print("mixed rel-max", rel(mixed, omixed.detach()))
def rl2(a, b):
    a = torch.as_tensor(a).double(); return float((a - b).norm() / b.norm())
print("gtp rel-l2", rl2(gtp, tpd.grad), "gmp rel-l2", rl2(gmp, mpd.grad), "gtracks rel-l2", rl2(gtr, trd.grad))
if "-v" in sys.argv:
    np.set_printoptions(precision=3, linewidth=200)
    print("gmp ours  ", np.asarray(gmp)[0]); print("gmp oracle", mpd.grad.numpy()[0])
    print("gtp ours  ", np.asarray(gtp)[0, 0]); print("gtp oracle", tpd.grad.numpy()[0, 0])
