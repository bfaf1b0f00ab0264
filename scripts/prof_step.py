"""A few bench steps (console fwd -> MRSTFT -> bwd) for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffmst_b200 import AdvancedMixConsole, MRSTFTLoss, batch_stereo_peak_normalize
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
con = AdvancedMixConsole(bench.SR).to(dev); con.materialize_tracks = False; con.check_ranges = "async"
loss_fn = MRSTFTLoss(**bench.RES)
tracks, tp, fp, mp, tp2, mp2 = bench.make_inputs(torch, 0, bench.B, "cpu")
tracks = tracks.to(dev); fp = fp.to(dev)
tp = tp.to(dev).requires_grad_(True); mp = mp.to(dev).requires_grad_(True)
with torch.no_grad():
    target = batch_stereo_peak_normalize(con(tracks, tp2.to(dev), fp, mp2.to(dev), **bench.FLAGS)[1])
for _ in range(reps):
    tp.grad = None; mp.grad = None
    mix = con(tracks, tp, fp, mp, **bench.FLAGS)[1]
    loss = loss_fn(mix, target)
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
