"""TEST INFRASTRUCTURE (uses the oracle as the comparison arm): Context number (not a bench contract line): the reference ALGORITHM (oracle port of dasp-pytorch's
frequency-sampling path + auraloss MRSTFT, float32) run as PyTorch ops on the same GPU, i.e. what the
reference itself would execute on a B200, next to our step.  Same workload as bench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from oracle.auraloss.freq import MultiResolutionSTFTLoss as OracleMR
from oracle.console import OracleAdvancedMixConsole
from oracle.loss import batch_stereo_peak_normalize as oracle_norm
from diffmst_b200 import AdvancedMixConsole, MRSTFTLoss, batch_stereo_peak_normalize

dev = torch.device("cuda", 0)
tracks, tp, fp, mp, tp2, mp2 = bench.make_inputs(torch, 0, bench.B, "cpu")
tracks = tracks.to(dev); fp = fp.to(dev); tp2 = tp2.to(dev); mp2 = mp2.to(dev)
tp = tp.to(dev).requires_grad_(True); mp = mp.to(dev).requires_grad_(True)

def timeit(step, n):
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ocon, oloss = OracleAdvancedMixConsole(bench.SR), OracleMR(**bench.RES)
with torch.no_grad():
    otarget = oracle_norm(ocon(tracks, tp2, fp, mp2, **bench.FLAGS)[1])
def ref_step():
    tp.grad = None; mp.grad = None
    oloss(ocon(tracks, tp, fp, mp, **bench.FLAGS)[1], otarget).backward()
t_ref = timeit(ref_step, 3)
peak = torch.cuda.max_memory_allocated() / 2**30

con = AdvancedMixConsole(bench.SR).to(dev); con.materialize_tracks = False; con.check_ranges = False
loss = MRSTFTLoss(**bench.RES)
with torch.no_grad():
    target = batch_stereo_peak_normalize(con(tracks, tp2, fp, mp2, **bench.FLAGS)[1])
def our_step():
    tp.grad = None; mp.grad = None
    loss(con(tracks, tp, fp, mp, **bench.FLAGS)[1], target).backward()
t_ours = timeit(our_step, 20)
units = bench.B * bench.N * bench.T / bench.SR
print(f"reference algorithm as PyTorch ops on this GPU (float32, FFT path): {t_ref:.2f} ms/step = {units / t_ref * 1e3:,.0f} track-s/s "
      f"(peak {peak:.1f} GiB) | ours {t_ours:.3f} ms/step = {units / t_ours * 1e3:,.0f} track-s/s | ratio {t_ref / t_ours:.1f}x")
