"""Diagnostic: master-bus / track parameter gradients of the headline step, ours vs float64 vs float32 oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_headline_gpu as h
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for T in (65536, 262144):
    g = torch.Generator().manual_seed(2026)
    inputs = h._inputs(1, 16, 262144, seed=2026)
    inputs = tuple(t[..., :T].contiguous() if t.dim() == 3 and t.shape[-1] == 262144 else t for t in inputs)
    o64 = h._oracle_step(inputs, torch.float64); o32 = h._oracle_step(inputs, torch.float32)
    ours = h._our_step(inputs, False)
    print("T", T, "loss", ours["loss"], o64["loss"], o32["loss"])
    print("gmp f64 ", o64["gmp"][0]); print("gmp f32 ", o32["gmp"][0]); print("gmp ours", ours["gmp"][0])
    print("rell2 gmp ours/f32:", h.rell2(ours["gmp"], o64["gmp"]), h.rell2(o32["gmp"], o64["gmp"]))
    print("rell2 gtp ours/f32:", h.rell2(ours["gtp"], o64["gtp"]), h.rell2(o32["gtp"], o64["gtp"]))
    e_o = np.abs(ours["gtp"] - o64["gtp"])[0]; e_r = np.abs(o32["gtp"] - o64["gtp"])[0]
    print("gtp abs err per param column (ours):", e_o.max(0)); print("gtp abs err per param column (f32): ", e_r.max(0))
    print("gtp f64 col max:", np.abs(o64["gtp"][0]).max(0))
