"""Small console forward + backward (training path: parameter gradients only; and the audio-gradient path) for
compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import AdvancedMixConsole, MRSTFTLoss
torch.manual_seed(0)
B, N, T = 2, 3, 20011   # ragged: 3 forward tiles / 5 backward tiles, T % 4 != 0
con = AdvancedMixConsole(44100).cuda(); con.check_ranges = "async"
loss_fn = MRSTFTLoss(fft_sizes=[512, 2048, 8192, 600], hop_sizes=[256, 1024, 4096, 150], win_lengths=[512, 2048, 8192, 400])   # fused front end (three plans) + the library-FFT path
x = (torch.randn(B, N, T) * 0.1).cuda()
fp = torch.rand(B, 25).cuda()
target = (torch.randn(B, 2, T) * 0.1).cuda()
for want_audio_grad, materialize in ((False, False), (False, True), (True, True)):
    tp = torch.rand(B, N, 27).cuda().requires_grad_(True)
    mp = torch.rand(B, 26).cuda().requires_grad_(True)
    xx = x.clone().requires_grad_(want_audio_grad)
    con.materialize_tracks = materialize
    mixed, mix = con(xx, tp, fp, mp, use_fx_bus=False)[:2]
    loss = loss_fn(mix, target) + (mixed.square().mean() if materialize else 0.0)
    loss.backward()
    torch.cuda.synchronize()
    print("ok", want_audio_grad, materialize, float(loss), float(tp.grad.abs().sum()), float(mp.grad.abs().sum()))
with torch.no_grad():
    print("no_grad", float(con(x, tp, fp, mp, use_fx_bus=False)[1].abs().sum()))
con.check_pending_ranges()
