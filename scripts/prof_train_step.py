"""Kernel-time breakdown of one DDP training step (single rank) with torch.profiler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from diffmst_b200 import AdvancedMixConsole, MRSTFTLoss, SpectrogramEncoder
from diffmst_b200.training import BucketedGradAllReduce, MixStyleTransferModel, TransformerController, training_step
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, N, T = 4, 8, 262144
model = MixStyleTransferModel(SpectrogramEncoder(embed_dim=512), SpectrogramEncoder(embed_dim=512),
                              TransformerController(512, 27, 25, 26, num_layers=12, nhead=8)).to(dev).train()
console = AdvancedMixConsole(44100).to(dev); console.materialize_tracks = False; console.check_ranges = "async"
loss_fn = MRSTFTLoss(fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096], win_lengths=[512, 2048, 8192])
reducer = BucketedGradAllReduce(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)
tracks = (torch.randn(B, N, T) * 0.1).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
for _ in range(3):
    training_step(model, console, loss_fn, tracks, reducer, opt, generator=gen)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        training_step(model, console, loss_fn, tracks, reducer, opt, generator=gen)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total kernel time per step {tot:.2f} ms")
for k, ms, c in rows[:28]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{c:<4d} {k[:110]}")
