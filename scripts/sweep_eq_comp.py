"""BASELINE.json configs[4]: ParametricEQ + Compressor kernel sweep, tracks in {1, 8, 64} x
samples in {65536, 262144, 1048576}: device time of the track chain kernels and achieved
algorithmic HBM GB/s (4 + 8/N bytes per track-sample per pass) against the measured roofline."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffmst_b200 import AdvancedMixConsole, _lib

peak = 6538.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
lib = _lib.lib()
con = AdvancedMixConsole(44100).cuda()
con.materialize_tracks = False
con.check_ranges = "async"
rows = []
NS = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 8, 64]
for N in NS:
    for T in (65536, 262144, 1048576):
        B = 1
        g = torch.Generator().manual_seed(N * 7 + T)
        x = (torch.randn(B, N, T, generator=g) * 0.1).cuda()
        tp = torch.rand(B, N, 27, generator=g).cuda().requires_grad_(True)
        fp = torch.rand(B, 25, generator=g).cuda()
        mp = torch.rand(B, 26, generator=g).cuda().requires_grad_(True)
        probe = torch.randn(B, 2, T, generator=g).cuda()
        kw = dict(use_master_bus=False, use_output_fader=False, use_fx_bus=False)
        def step():
            tp.grad = None; mp.grad = None
            con(x, tp, fp, mp, **kw)[1].backward(probe)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = 10
        lib.dmst_profile_enable(n)
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        buf = (ctypes.c_float * n)()
        ms = {}
        for kind, name in ((0, "track_fwd"), (3, "track_bwd")):
            k = lib.dmst_profile_read(kind, buf, n)
            ms[name] = sorted(buf[i] for i in range(k))[k // 2]
        lib.dmst_profile_enable(0)
        byts = (4 + 8 / N) * B * N * T
        rows.append(dict(tracks=N, samples=T, fwd_ms=ms["track_fwd"], bwd_ms=ms["track_bwd"],
                         fwd_gbs=byts / ms["track_fwd"] / 1e6, bwd_gbs=byts / ms["track_bwd"] / 1e6))
        r = rows[-1]
        print(f"N={N:3d} T={T:8d}  fwd {r['fwd_ms']*1e3:8.1f} us {r['fwd_gbs']:7.1f} GB/s ({r['fwd_gbs']/peak*100:5.2f}% of {peak:.0f})"
              f"   bwd {r['bwd_ms']*1e3:8.1f} us {r['bwd_gbs']:7.1f} GB/s ({r['bwd_gbs']/peak*100:5.2f}%)", flush=True)
print(json.dumps(rows))
