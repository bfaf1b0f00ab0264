"""Ad-hoc device timing of the console at BASELINE configs[1] (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import AdvancedMixConsole

B, N, T = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 16, 262144)))
torch.manual_seed(0)
x = (torch.randn(B, N, T) * 0.1).cuda()
tp = torch.rand(B, N, 27).cuda().requires_grad_(True)
fp = torch.rand(B, 25).cuda()
mp = torch.rand(B, 26).cuda().requires_grad_(True)
con = AdvancedMixConsole(44100).cuda()
con.materialize_tracks = False
con.check_ranges = False
probe = torch.randn(B, 2, T).cuda()

def fwd():
    with torch.no_grad():
        return con(x, tp, fp, mp, use_fx_bus=False)[1]

def fwdbwd():
    tp.grad = None; mp.grad = None
    mix = con(x, tp, fp, mp, use_fx_bus=False)[1]
    mix.backward(probe)

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    return min(ts), sorted(ts)[len(ts) // 2]

for name, fn in (("fwd", fwd), ("fwd+bwd", fwdbwd)):
    best, med = timeit(fn)
    tsec = B * N * T / 44100
    print(f"{name}: best {best:.3f} ms  median {med:.3f} ms  -> {tsec / (med / 1e3):.3e} track-s/s (B={B},N={N},T={T})")
con.materialize_tracks = True
best, med = timeit(fwd)
print(f"fwd (materialize mixed_tracks): median {med:.3f} ms")
