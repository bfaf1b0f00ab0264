"""Analytic known-answer tests that pin the oracle shims independently of upstream
(the reference's own tests assert nothing; these are the assertions its
scripts tests/test_comp.py, test_peq.py, test_panner.py imply)."""
import math

import numpy as np
import pytest
import torch

from oracle.dasp_pytorch import functional as F
from oracle.dasp_pytorch import signal as S
from oracle import timedomain as td
from oracle.auraloss.freq import MultiResolutionSTFTLoss, STFTLoss

SR = 44100


def _freq_gain_db(b, a, f):
    z = np.exp(-1j * 2 * math.pi * f / SR)
    num = b[0] + b[1] * z + b[2] * z * z
    den = a[0] + a[1] * z + a[2] * z * z
    return 20 * math.log10(abs(num / den))


@pytest.mark.parametrize("g", [-12.0, -3.0, 6.0, 12.0])
def test_biquad_gains_at_reference_points(g):
    t = lambda v: torch.tensor([v], dtype=torch.float64)
    b, a = S.biquad(t(g), t(1000.0), t(0.707), SR, "peaking")
    assert abs(_freq_gain_db(b[0].numpy(), a[0].numpy(), 1000.0) - g) < 1e-9
    b, a = S.biquad(t(g), t(200.0), t(0.707), SR, "low_shelf")
    assert abs(_freq_gain_db(b[0].numpy(), a[0].numpy(), 0.0) - g) < 1e-9
    b, a = S.biquad(t(g), t(8000.0), t(0.707), SR, "high_shelf")
    assert abs(_freq_gain_db(b[0].numpy(), a[0].numpy(), SR / 2) - g) < 1e-9
    # time-domain restatement uses the same formulas
    bb, aa = td.rbj(g, 1000.0, 0.707, SR, "peaking")
    b, a = S.biquad(t(g), t(1000.0), t(0.707), SR, "peaking")
    assert np.allclose(bb, b[0].numpy(), atol=1e-15) and np.allclose(aa, a[0].numpy(), atol=1e-15)


def test_panner_law_matches_reference_test_values():
    # tests/test_panner.py:4-8: ones x pan [0, .5, 1, 0] -> L [1, .5946, 0, 1], R [0, .5946, 1, 0]
    x = torch.ones(1, 4, 1, dtype=torch.float64)
    y = F.stereo_panner(x, SR, torch.tensor([[0.0, 0.5, 1.0, 0.0]], dtype=torch.float64))
    assert y.shape == (1, 2, 4, 1)
    c = math.sqrt(0.5 * math.cos(math.pi / 4))
    assert torch.allclose(y[0, 0, :, 0], torch.tensor([1.0, c, 0.0, 1.0], dtype=torch.float64), atol=1e-12)
    assert torch.allclose(y[0, 1, :, 0], torch.tensor([0.0, c, 1.0, 0.0], dtype=torch.float64), atol=1e-12)
    assert abs(c - 0.5946) < 1e-4
    assert abs(float(y.sum(2)[0, 0, 0]) - 2.5946) < 1e-4


def test_compressor_step_response():
    # tests/test_comp.py:18-36 made analytic: unit step, thr -12 dB, ratio 4, knee 6, attack 100 ms
    T = 65536
    x = torch.zeros(1, 1, T, dtype=torch.float64)
    x[..., 4096:] = 1.0
    p = lambda v: torch.tensor([v], dtype=torch.float64)
    y = F.compressor(x, SR, p(-12.0), p(4.0), p(100.0), p(0.0), p(6.0), p(0.0))
    alpha = math.exp(-math.log(9.0) / (SR * 0.1))
    gc = (1 / 4 - 1) * (0.0 - (-12.0))  # above the knee: -9 dB
    # input below eps before the step: x_db = -160 dB -> no gain reduction
    n = np.arange(T - 4096)
    expected = 10 ** (gc * (1 - alpha ** (n + 1)) / 20.0)
    assert np.allclose(y[0, 0, 4096:].numpy(), expected, rtol=0, atol=1e-9)
    assert float(y[0, 0, :4096].abs().max()) == 0.0
    assert abs(20 * math.log10(float(y[0, 0, -1])) - (-9.0)) < 1e-3
    # release_ms is a dummy parameter (tests/test_comp.py:27, mst/modules.py:377)
    y2 = F.compressor(x, SR, p(-12.0), p(4.0), p(100.0), p(250.0), p(6.0), p(0.0))
    assert torch.equal(y, y2)
    # lookahead delays the signal path only and zeroes the head
    y3 = F.compressor(x, SR, p(-12.0), p(4.0), p(100.0), p(0.0), p(6.0), p(0.0), lookahead_samples=1024)
    g = 10 ** (gc * (1 - alpha ** (n + 1)) / 20.0)
    assert float(y3[0, 0, : 4096 + 1024].abs().max()) == 0.0
    assert np.allclose(y3[0, 0, 4096 + 1024:].numpy(), g[1024:], atol=1e-9)


def test_knee_region_is_continuous_and_quadratic():
    x_db = torch.linspace(-30, 0, 3001, dtype=torch.float64)
    g = F.compressor_gain_computer(x_db, torch.tensor(-12.0), torch.tensor(4.0), torch.tensor(6.0))
    assert float(g[x_db < -15.0].abs().max()) == 0.0
    mid = x_db == -12.0
    assert abs(float(g[mid]) - (1 / 4 - 1) * 9 / 12) < 1e-12
    assert float((g[1:] - g[:-1]).abs().max()) < 0.01
    gt = td.gain_computer(x_db.numpy(), -12.0, 4.0, 6.0)
    assert np.allclose(gt, g.numpy(), atol=1e-13)


def test_fsm_equals_time_domain_recursion_fp64():
    g = torch.Generator().manual_seed(0)
    T = 65536
    x = torch.randn(2, 1, T, generator=g, dtype=torch.float64) * 0.1
    eq = [3.0, 120.0, 0.9, -6.0, 400.0, 2.0, 9.0, 3000.0, 0.5, -4.0, 9000.0, 3.0, 5.0, 15000.0, 1.0,
          -8.0, 10000.0, 0.8]
    args = [torch.full((2,), v, dtype=torch.float64) for v in eq]
    y = F.parametric_eq(x, SR, *args)
    yt = td.parametric_eq(x[:, 0].numpy(), SR, eq)
    assert np.max(np.abs(y[:, 0].numpy() - yt)) / np.max(np.abs(yt)) < 1e-10
    p = lambda v: torch.full((2,), v, dtype=torch.float64)
    c = F.compressor(y, SR, p(-30.0), p(5.0), p(20.0), p(50.0), p(6.0), p(2.0), lookahead_samples=2048)
    ct = np.stack([td.compressor(yt[i:i + 1], SR, -30.0, 5.0, 20.0, 6.0, 2.0, 2048)[0] for i in range(2)])
    assert np.max(np.abs(c[:, 0].numpy() - ct)) / np.max(np.abs(ct)) < 1e-10


def test_stft_loss_closed_forms():
    g = torch.Generator().manual_seed(1)
    y = torch.randn(2, 2, 8192, generator=g, dtype=torch.float64)
    f = MultiResolutionSTFTLoss([512, 2048], [256, 1024], [512, 2048])
    assert float(f(y, y)) == 0.0
    # x = c*y: |X| = c|Y| -> SC = |1-c|, log-L1 = |ln c| at every resolution
    c = 0.5
    assert abs(float(f(c * y, y)) - (abs(1 - c) + abs(math.log(c)))) < 1e-9
    f2 = MultiResolutionSTFTLoss([512], [256], [512], w_sc=0.0, w_log_mag=0.0, w_lin_mag=1.0)
    one = STFTLoss(512, 256, 512)
    assert abs(float(f2(c * y, y)) - (1 - c) * float(one.stft(y.reshape(-1, 8192)).mean())) < 1e-9
    # silence hits the 1e-8 clamp inside the sqrt: |X| = 1e-4
    z = torch.zeros(1, 2, 4096, dtype=torch.float64)
    assert torch.allclose(one.stft(z.reshape(-1, 4096)), torch.full((2, 257, 17), 1e-4, dtype=torch.float64))
