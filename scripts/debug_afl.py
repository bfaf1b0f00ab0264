import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from diffmst_b200 import AudioFeatureLoss, MRSTFTLoss
from oracle.auraloss.freq import MultiResolutionSTFTLoss as O
g = torch.Generator().manual_seed(0)
y = (torch.randn(8, 2, 262144, generator=g) * 0.1).cuda()
afl = AudioFeatureLoss([0.1, 0.001, 1.0, 1.0, 0.1], 44100)
for i in range(2):
    print({k: float(v) for k, v in afl(y, y).items()})
y2 = (torch.randn(2, 2, 40000, generator=g) * 0.1).cuda()
print({k: float(v) for k, v in afl(y2, y2).items()})
d = dict(np.load('tests/golden/mrstft_train.npz'))
RES = dict(fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096], win_lengths=[512, 2048, 8192])
o = O(**RES)
x = torch.from_numpy(d['x']).cuda().requires_grad_(True)
l = o(x, torch.from_numpy(d['y']).cuda()); l.backward()
gr = x.grad.cpu().numpy()
print('torch CUDA fp32 shim grad relL2 vs fp64 golden', np.linalg.norm(gr - d['grad_x']) / np.linalg.norm(d['grad_x']))
f = MRSTFTLoss(**RES)
x2 = torch.from_numpy(d['x']).cuda().requires_grad_(True)
l2 = f(x2, torch.from_numpy(d['y']).cuda()); l2.backward()
go = x2.grad.cpu().numpy()
print('ours grad relL2 vs golden', np.linalg.norm(go - d['grad_x']) / np.linalg.norm(d['grad_x']), ' ours vs torch-cuda', np.linalg.norm(go - gr) / np.linalg.norm(gr))
for r in range(3):
    ff = MRSTFTLoss([RES['fft_sizes'][r]], [RES['hop_sizes'][r]], [RES['win_lengths'][r]])
    oo = O([RES['fft_sizes'][r]], [RES['hop_sizes'][r]], [RES['win_lengths'][r]])
    xa = torch.from_numpy(d['x']).cuda().requires_grad_(True); ff(xa, torch.from_numpy(d['y']).cuda()).backward()
    xb = torch.from_numpy(d['x']).double().requires_grad_(True); oo(xb, torch.from_numpy(d['y']).double()).backward()
    xc = torch.from_numpy(d['x']).cuda().requires_grad_(True); oo(xc, torch.from_numpy(d['y']).cuda()).backward()
    gb = xb.grad.numpy()
    print(' res', RES['fft_sizes'][r], 'ours', np.linalg.norm(xa.grad.cpu().numpy() - gb) / np.linalg.norm(gb), 'torch-cuda-fp32', np.linalg.norm(xc.grad.cpu().numpy() - gb) / np.linalg.norm(gb))
