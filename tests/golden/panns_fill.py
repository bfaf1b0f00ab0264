"""Deterministic, generator-free parameter fill shared by make_golden_panns.py (which applies it to the REFERENCE's
mst.panns classes) and the tests (which apply it to the oracle / product modules): every floating-point entry of the
state dict, visited in sorted-name order, gets a sine pattern scaled so that activations stay alive through the six
blocks.  (The reference initialises with the global RNG through Conv2d's constructor plus xavier_uniform_, which a
restatement cannot replay; weights as large as Cnn14's cannot be stored either.)"""
import math

import torch


def fill_state(module: torch.nn.Module) -> None:
    sd = module.state_dict()
    for k, name in enumerate(sorted(sd)):
        t = sd[name]
        if not t.is_floating_point():
            continue
        i = torch.arange(t.numel(), dtype=torch.float64)
        s = torch.sin(0.37 * i + 1.3 * k)
        if name.endswith("running_var"):
            v = 1.0 + 0.2 * s * s
        elif name.endswith("running_mean"):
            v = 0.02 * s
        elif ".bn" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * s
        elif name.endswith("bias"):
            v = 0.05 * s
        else:  # conv / linear weights: unit-gain scaling for ReLU networks
            fan_in = t[0].numel()
            v = s * math.sqrt(2.0 / fan_in) * 1.4
        t.copy_(v.reshape(t.shape).to(t.dtype))
