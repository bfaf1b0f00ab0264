"""TEST INFRASTRUCTURE (uses the oracle as the comparison arm): Cnn14 forward + backward (training mode, batch statistics) at the encoder's real input size: ours with the CUDA
BatchNorm/ReLU/pooling Functions, ours with those on PyTorch ops, and the reference modules on cuDNN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import Cnn14, conv
from oracle.panns import OracleCnn14

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
x = torch.rand(B, 1, 1025, 257, device=dev) ** 3


def timeit(model, n=5):
    def step():
        for p in model.parameters():
            p.grad = None
        model(x).square().mean().backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, torch.cuda.max_memory_allocated() / 2**30


ours = Cnn14(num_classes=512).to(dev).train()
ref = OracleCnn14(num_classes=512).to(dev).train()
ref.load_state_dict(ours.state_dict())
torch.cuda.reset_peak_memory_stats()
t, m = timeit(ours)
print(f"Cnn14 fwd+bwd (train), batch {B} x 1025 x 257: ours, CUDA BN/ReLU/pool Functions   {t:8.2f} ms   peak {m:.1f} GiB")
conv._CUDA_BN_POOL = False
torch.cuda.reset_peak_memory_stats()
t, m = timeit(ours)
print(f"                                              ours, BN/ReLU/pool on PyTorch ops    {t:8.2f} ms   peak {m:.1f} GiB")
conv._CUDA_BN_POOL = True
torch.cuda.reset_peak_memory_stats()
t, m = timeit(ref)
print(f"                                              reference modules (cuDNN TF32, NCHW) {t:8.2f} ms   peak {m:.1f} GiB")
