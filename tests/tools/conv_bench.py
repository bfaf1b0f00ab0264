"""TEST INFRASTRUCTURE (uses the oracle as the comparison arm): Per-layer timing of the Cnn14 conv trunk (tensor-core row): TFLOP/s per 3x3 convolution at the
encoder's real shapes (mst/modules.py:786-806: 1025 bins x 257 frames), batch of items = argv[1]."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from diffmst_b200 import _lib
from diffmst_b200.conv import _ptr, _stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = _lib.lib()
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
tf32_peak = peaks["bf16_tflops"] / 2.0   # dense TF32 is half the BF16 rate on this part
layers = [(64, 64, 1025, 257), (64, 128, 512, 128), (128, 128, 512, 128), (128, 256, 128, 32), (256, 256, 128, 32),
          (256, 512, 32, 16), (512, 512, 32, 16), (512, 1024, 8, 8), (1024, 1024, 8, 8), (1024, 2048, 2, 4), (2048, 2048, 2, 4)]
tot_flop = tot_ms = 0.0
for cin, cout, H, W in layers:
    x = torch.randn(B, H + 2, W + 2, cin, device=dev)
    w9 = torch.randn(9, cout, cin, device=dev) * 0.05
    sc = torch.rand(cout, device=dev); sh = torch.rand(cout, device=dev)
    y = torch.empty(B, H + 2, W + 2, cout, device=dev)
    nws = lib.dmst_conv3x3_workspace_bytes(B, H, W, cin, cout)
    ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=dev)
    def run():
        rc = lib.dmst_conv3x3_forward_ws(_ptr(x), _ptr(w9), _ptr(sc), _ptr(sh), _ptr(y), B, H, W, cin, cout, 1, _ptr(ws), nws,
                                         _stream(dev))
        assert rc == 0, rc
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    ev[0].record()
    for _ in range(n): run()
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    flop = 2.0 * 9 * cin * cout * B * H * W
    tot_flop += flop; tot_ms += ms
    # the reference's own path for this op: torch conv2d -> cuDNN, TF32 allowed (torch default), NCHW and channels_last
    torch.backends.cudnn.allow_tf32 = True; torch.backends.cudnn.benchmark = True
    xr = torch.randn(B, cin, H, W, device=dev); wr = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    best = 1e9
    for fmt in (torch.contiguous_format, torch.channels_last):
        xx, ww = xr.contiguous(memory_format=fmt), wr.contiguous(memory_format=fmt)
        with torch.no_grad():
            for _ in range(3): torch.nn.functional.conv2d(xx, ww, padding=1)
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(n): torch.nn.functional.conv2d(xx, ww, padding=1)
            ev[1].record(); torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]) / n)
    tot_ref = globals().get("tot_ref", 0.0) + best; globals()["tot_ref"] = tot_ref
    print(f"conv {cin:5d}->{cout:5d} @ {H:4d}x{W:3d} x{B}: ours {ms*1e3:8.1f} us {flop/ms/1e9:7.1f} TFLOP/s ({flop/ms/1e9/tf32_peak*100:5.1f}% of {tf32_peak:.0f} TF32 = measured bf16/2) | cuDNN(tf32) {best*1e3:8.1f} us {flop/best/1e9:7.1f} TFLOP/s", flush=True)
print(f"total ours {tot_ms:.3f} ms ({tot_flop/tot_ms/1e9:.1f} TFLOP/s) | cuDNN conv only {tot_ref:.3f} ms ({tot_flop/tot_ref/1e9:.1f} TFLOP/s) for {tot_flop/1e12:.2f} TFLOP (ours includes the BN+ReLU epilogue)")

# ---- whole trunk, apples to apples: the reference's path is conv2d (cuDNN, TF32 allowed) + BatchNorm2d + ReLU +
# avg_pool2d as separate kernels (mst/panns.py:49-85); ours fuses BN + ReLU into the convolution epilogue ----
from diffmst_b200 import Cnn14
from oracle.panns import OracleCnn14
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.benchmark = True
ref = OracleCnn14(num_classes=512).to(dev).eval()
ours = Cnn14(num_classes=512).to(dev).eval()
ours.load_state_dict(ref.state_dict())
xs = (torch.rand(B, 1, 1025, 257, device=dev) ** 3)
def timeit(fn, n=5):
    with torch.no_grad():
        for _ in range(2): fn()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(n): fn()
        ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n
t_ours = timeit(lambda: ours(xs))
t_ref = timeit(lambda: ref(xs))
t_ref_cl = timeit(lambda: ref.to(memory_format=torch.channels_last)(xs.contiguous(memory_format=torch.channels_last)))
print(f"Cnn14 forward (eval), batch {B} x 1025 x 257: ours {t_ours:.3f} ms | reference modules (cuDNN TF32 + separate BN/ReLU/pool) "
      f"{t_ref:.3f} ms NCHW, {t_ref_cl:.3f} ms channels_last", flush=True)
