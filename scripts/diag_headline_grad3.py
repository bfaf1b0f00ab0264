import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_headline_gpu as h
import test_console_gpu as c
np.set_printoptions(precision=4, suppress=True, linewidth=220)
T = 262144
tracks, tp, fp, mp, tp2, mp2 = h._inputs(1, 16, T, seed=2026)
print("master comp (thr, ratio, attack, rel, knee, makeup):", (mp[0,18]*60-60).item(), (mp[0,19]*9+1).item(), (mp[0,20]*245+5).item(), 0, (mp[0,22]*9+3).item(), (mp[0,23]*6).item(), "out fader dB", (mp[0,24]*96-48).item(), "in fader dB", (mp[0,25]*96-48).item())
g = torch.Generator().manual_seed(5)
probe = torch.randn(1, 2, T, generator=g)
for name, flags in (("all", dict(use_fx_bus=False)), ("no master", dict(use_fx_bus=False, use_master_bus=False, use_output_fader=False)),
                    ("no master, no comp", dict(use_fx_bus=False, use_master_bus=False, use_output_fader=False, use_track_compressor=False)),
                    ("no master, no eq", dict(use_fx_bus=False, use_master_bus=False, use_output_fader=False, use_track_eq=False))):
    ours = c.run_ours(tracks, tp, fp, mp, probe, flags)
    o64 = c.run_oracle(tracks, tp, fp, mp, probe, flags, torch.float64)
    o32 = c.run_oracle(tracks, tp, fp, mp, probe, flags, torch.float32)
    sc = np.abs(o64["mix"]).max()
    print("==", name, "max|mix|", sc)
    print("  mix relmax ours/f32", c.relmax(ours["mix"], o64["mix"]), c.relmax(o32["mix"], o64["mix"]))
    print("  per 32768 segment ours:", ["%.1e" % (np.abs(ours["mix"][..., s:s+32768] - o64["mix"][..., s:s+32768]).max() / sc) for s in range(0, T, 32768)])
    print("  per 32768 segment f32 :", ["%.1e" % (np.abs(o32["mix"][..., s:s+32768] - o64["mix"][..., s:s+32768]).max() / sc) for s in range(0, T, 32768)])
    print("  gtp rell2 ours/f32", c.rell2(ours["gtp"], o64["gtp"]), c.rell2(o32["gtp"], o64["gtp"]), " gmp", c.rell2(ours["gmp"], o64["gmp"]), c.rell2(o32["gmp"], o64["gmp"]))
    eo = np.abs(ours["gtp"] - o64["gtp"])[0]; er = np.abs(o32["gtp"] - o64["gtp"])[0]
    print("  gtp col abs err ours:", eo.max(0)); print("  gtp col abs err f32 :", er.max(0)); print("  gtp col max f64     :", np.abs(o64["gtp"][0]).max(0))
    print("  worst track (ours):", eo.max(1))
