"""CPU check of the console KERNEL LOGIC: the same .cuh sources compiled for the host-side emulator
(tests/emul, g++ -DDMST_EMULATE) run a small forward + backward and are compared with the float64 oracle.
This is test infrastructure (it can never be the product path: the product library refuses to load unless it is a
device build); it lets the GPU-less CI see a broken scan, chain hand-off or adjoint before the GPU tier does."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))


@pytest.fixture(scope="module")
def emul_console():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    res = subprocess.run(["bash", os.path.join(ROOT, "tests", "emul", "build_emul.sh")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    import harness
    return harness


def rel_max(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# (lengths >= 32768: below that the oracle's own FFT method time-aliases; the second case has a
# ragged last tile and more than one item)
@pytest.mark.parametrize("shape", [(1, 2, 32768), (2, 3, 36871)])
def test_emulated_console_kernels_match_float64_oracle(emul_console, shape):
    from oracle.console import OracleAdvancedMixConsole
    B, N, T = shape
    g = torch.Generator().manual_seed(T)
    tracks = torch.randn(B, N, T, generator=g) * 0.1
    tp, fp, mp = torch.rand(B, N, 27, generator=g), torch.rand(B, 25, generator=g), torch.rand(B, 26, generator=g)
    probe = torch.randn(B, 2, T, generator=g)
    con = emul_console.EmulConsole()
    mix, mixed, status = con.forward(tracks.numpy(), tp.numpy(), mp.numpy(), emul_console.flags_from(), want_mixed=True)
    gtp, gmp, gtr = con.backward(probe.numpy())
    orc = OracleAdvancedMixConsole(44100)
    tpd, mpd, trd = (t.double().requires_grad_(True) for t in (tp, mp, tracks))
    omixed, omix, _, _, _ = orc(trd, tpd, fp.double(), mpd, use_fx_bus=False)
    (omix * probe.double()).sum().backward()
    assert rel_max(mix, omix.detach().numpy()) <= 1e-4
    assert rel_max(mixed, omixed.detach().numpy()) <= 1e-4
    assert rel_l2(gtp, tpd.grad.numpy()) <= 1e-3
    assert rel_l2(gmp, mpd.grad.numpy()) <= 1e-3
    assert rel_l2(gtr, trd.grad.numpy()) <= 1e-3
