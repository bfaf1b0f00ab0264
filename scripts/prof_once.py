"""One forward+backward of the console at BASELINE configs[1] for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import AdvancedMixConsole
B, N, T = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 16, 262144)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(0)
x = (torch.randn(B, N, T) * 0.1).cuda()
tp = torch.rand(B, N, 27).cuda().requires_grad_(True)
fp = torch.rand(B, 25).cuda()
mp = torch.rand(B, 26).cuda().requires_grad_(True)
con = AdvancedMixConsole(44100).cuda()
con.materialize_tracks = False
con.check_ranges = "async"
probe = torch.randn(B, 2, T).cuda()
for _ in range(reps):
    tp.grad = None; mp.grad = None
    mix = con(x, tp, fp, mp, use_fx_bus=False)[1]
    mix.backward(probe)
torch.cuda.synchronize()
print("done")
