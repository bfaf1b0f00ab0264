"""Conditioning survey of the headline step (1 x 16 x 262144, training flags, MRSTFT): ours and the reference
algorithm's float32 evaluation, both against float64, over several parameter draws."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_headline_gpu as h
print("seed | mix relmax ours / ref-f32 | loss rel ours / ref-f32 | grad track_params rel-L2 ours / ref-f32 | grad master rel-L2 ours / ref-f32")
for seed in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5, 6, 7, 8, 2026]:
    inputs = h._inputs(1, 16, 262144, seed=seed)
    o64 = h._oracle_step(inputs, torch.float64); o32 = h._oracle_step(inputs, torch.float32)
    ours = h._our_step(inputs, False)
    r = lambda a, b: abs(a - b) / abs(b)
    print("%5d | %.1e / %.1e | %.1e / %.1e | %.1e / %.1e | %.1e / %.1e" % (
        seed, h.relmax(ours["mix"], o64["mix"]), h.relmax(o32["mix"], o64["mix"]), r(ours["loss"], o64["loss"]), r(o32["loss"], o64["loss"]),
        h.rell2(ours["gtp"], o64["gtp"]), h.rell2(o32["gtp"], o64["gtp"]), h.rell2(ours["gmp"], o64["gmp"]), h.rell2(o32["gmp"], o64["gmp"])), flush=True)
