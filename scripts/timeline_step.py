"""Device timeline of one headline step (console fwd -> MRSTFT -> bwd) replayed as a CUDA graph: start offset,
duration and stream of every kernel (torch.profiler / CUPTI), to see what overlaps and where the gaps are."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from diffmst_b200 import AdvancedMixConsole, MRSTFTLoss, GraphedStep, batch_stereo_peak_normalize
dev = torch.device("cuda", 0)
con = AdvancedMixConsole(bench.SR).to(dev); con.materialize_tracks = False; con.check_ranges = "async"
loss_fn = MRSTFTLoss(**bench.RES)
tracks, tp, fp, mp, tp2, mp2 = bench.make_inputs(torch, 0, bench.B, "cpu")
tracks = tracks.to(dev); fp = fp.to(dev)
tp = tp.to(dev).requires_grad_(True); mp = mp.to(dev).requires_grad_(True)
with torch.no_grad():
    target = batch_stereo_peak_normalize(con(tracks, tp2.to(dev), fp, mp2.to(dev), **bench.FLAGS)[1])
graphed = len(sys.argv) < 2 or sys.argv[1] != "eager"
if graphed:
    step = GraphedStep(lambda: loss_fn(con(tracks, tp, fp, mp, **bench.FLAGS)[1], target), params=[tp, mp], consoles=[con], warmup=2)
else:
    def step():
        tp.grad = None; mp.grad = None
        loss_fn(con(tracks, tp, fp, mp, **bench.FLAGS)[1], target).backward()
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time > 0]
evs.sort(key=lambda e: e.time_range.start)
# last step = last third of the events
n = len(evs) // 3
last = evs[-n:]
t0 = last[0].time_range.start
print(f"{'start us':>9} {'dur us':>8} {'end us':>8}  name")
for e in last:
    s = e.time_range.start - t0
    print(f"{s:9.1f} {e.device_time:8.1f} {s + e.device_time:8.1f}  {e.name[:90]}")
