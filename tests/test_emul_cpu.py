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
@pytest.mark.parametrize("shape", [(1, 2, 32768), (2, 3, 36871), (1, 1, 33001)])
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


@pytest.mark.parametrize("flags", [dict(use_track_compressor=False, use_master_bus=False),          # EQ only
                                   dict(use_track_eq=False, use_master_bus=False),                  # compressor only
                                   dict(use_track_eq=False, use_track_compressor=False),            # gain / pan + master bus
                                   dict(use_track_input_fader=False, use_output_fader=False)])      # faders off
def test_emulated_console_flag_combinations(emul_console, flags):
    """The chain's stages are switched by flags (mst/modules.py:192-198): each combination takes different branches of
    the forward and adjoint kernels (no EQ adjoint, no compressor halo, neutral master row)."""
    from oracle.console import OracleAdvancedMixConsole
    B, N, T = 1, 3, 32768
    g = torch.Generator().manual_seed(77)
    tracks = torch.randn(B, N, T, generator=g) * 0.1
    tp, fp, mp = torch.rand(B, N, 27, generator=g), torch.rand(B, 25, generator=g), torch.rand(B, 26, generator=g)
    probe = torch.randn(B, 2, T, generator=g)
    con = emul_console.EmulConsole()
    mix, mixed, status = con.forward(tracks.numpy(), tp.numpy(), mp.numpy(), emul_console.flags_from(**flags), want_mixed=True)
    gtp, gmp, gtr = con.backward(probe.numpy())
    def oracle(dtype):
        tpd, mpd, trd = (t.to(dtype).requires_grad_(True) for t in (tp, mp, tracks))
        omixed, omix, _, _, _ = OracleAdvancedMixConsole(44100)(trd, tpd, fp.to(dtype), mpd, use_fx_bus=False, **flags)
        (omix * probe.to(dtype)).sum().backward()
        gm = mpd.grad if mpd.grad is not None else torch.zeros_like(mpd)
        return [t.detach().double().numpy() for t in (omix, omixed, tpd.grad, trd.grad, gm)]

    o64, o32 = oracle(torch.float64), oracle(torch.float32)
    # tolerance protocol of the GPU tests: 1e-4 / 1e-3, or the reference algorithm's own float32
    # distance from float64 on the same inputs (x1.5) where that is larger
    for ours, w64, w32, err, tol in ((mix, o64[0], o32[0], rel_max, 1e-4), (mixed, o64[1], o32[1], rel_max, 1e-4),
                                     (gtp, o64[2], o32[2], rel_l2, 1e-3), (gtr, o64[3], o32[3], rel_l2, 1e-3)):
        assert err(ours, w64) <= max(tol, 1.5 * err(w32, w64)), (err(ours, w64), err(w32, w64))
    if float(np.abs(o64[4]).max()) > 0:
        assert rel_l2(gmp, o64[4]) <= max(1e-3, 1.5 * rel_l2(o32[4], o64[4]))
    else:
        assert float(np.abs(np.nan_to_num(gmp)).max()) == 0.0


def test_emulated_basic_console_config0_shape(emul_console):
    """BASELINE configs[0]: BasicMixConsole (gain + pan + bus sum), 4 tracks x 44100 samples, batch 1.  Gain and pan are
    per-track scalars and the bus sum is an indexing contract: float32-exact against the oracle evaluated in float32
    is not required of a different summation order, so the bound is 1e-6 relative (observed ~1e-7)."""
    from oracle.console import OracleBasicMixConsole
    B, N, T = 1, 4, 44100
    g = torch.Generator().manual_seed(44100)
    tracks = torch.randn(B, N, T, generator=g) * 0.1
    tp = torch.rand(B, N, 2, generator=g)
    probe = torch.randn(B, 2, T, generator=g)
    con = emul_console.EmulConsole()
    flags = emul_console.BASIC_CONSOLE | emul_console.USE_TRACK_INPUT_FADER | emul_console.USE_TRACK_PANNER
    mix, mixed, _ = con.forward(tracks.numpy(), tp.numpy(), None, flags, la_t=0, la_m=0, want_mixed=True)
    gtp, _, gtr = con.backward(probe.numpy())
    tpd, trd = tp.double().requires_grad_(True), tracks.double().requires_grad_(True)
    omixed, omix, _, _, _ = OracleBasicMixConsole(44100)(trd, tpd)
    (omix * probe.double()).sum().backward()
    assert mix.shape == (B, 2, T) and mixed.shape == (B, 2, N, T)
    assert rel_max(mix, omix.detach().numpy()) <= 1e-6
    assert rel_max(mixed, omixed.detach().numpy()) <= 1e-6
    assert rel_l2(gtp, tpd.grad.numpy()) <= 1e-5
    assert rel_l2(gtr, trd.grad.numpy()) <= 1e-6


@pytest.mark.parametrize("shape,flagkw", [((1, 2, 32768), {}), ((2, 2, 36871), {}), ((1, 1, 33001), dict(use_track_compressor=False)),
                                          ((1, 2, 32768), dict(use_track_eq=False))])
def test_emulated_parameter_only_backward_matches_float64_oracle(emul_console, shape, flagkw):
    """The training-mode track backward (console_bwd2.cuh: commuting-sections EQ gradients, delta-form recursions,
    no section checkpoints) - taken when no gradient w.r.t. the audio is requested - against float64 autograd."""
    from oracle.console import OracleAdvancedMixConsole
    B, N, T = shape
    g = torch.Generator().manual_seed(T + 1)
    tracks = torch.randn(B, N, T, generator=g) * 0.1
    tp, fp, mp = torch.rand(B, N, 27, generator=g), torch.rand(B, 25, generator=g), torch.rand(B, 26, generator=g)
    probe = torch.randn(B, 2, T, generator=g)
    con = emul_console.EmulConsole()
    mix, mixed, status = con.forward(tracks.numpy(), tp.numpy(), mp.numpy(), emul_console.flags_from(**flagkw), want_mixed=False,
                                     want_grad_tracks=False)
    gtp, gmp, gtr = con.backward(probe.numpy())
    assert gtr is None
    orc = OracleAdvancedMixConsole(44100)
    tpd, mpd = (t.double().requires_grad_(True) for t in (tp, mp))
    kw = dict(use_fx_bus=False); kw.update(flagkw)
    omix = orc(tracks.double(), tpd, fp.double(), mpd, **kw)[1]
    (omix * probe.double()).sum().backward()
    assert rel_max(mix, omix.detach().numpy()) <= 1e-4
    assert np.isfinite(gtp).all()
    assert rel_l2(gtp, tpd.grad.numpy()) <= 1e-3, (rel_l2(gtp, tpd.grad.numpy()), np.abs(gtp - tpd.grad.numpy()).max(axis=(0, 1)))
    assert rel_l2(gmp, mpd.grad.numpy()) <= 1e-3


# (fft size, hop, window, rows, samples): every radix plan of the fused STFT front end (complex lengths 32 .. 4096:
# 16x2, 16x4, 16x8, 16x16, 16x16x2, 16x16x4, 16x16x8, 16x16x16), half-overlapping and auraloss-default framing, windows
# shorter than the FFT, odd lengths (reflect padding and the scalar load path), several frames per block
@pytest.mark.parametrize("cfg", [(512, 256, 512, 2, 5000), (2048, 1024, 2048, 2, 9000), (8192, 4096, 8192, 1, 20000),
                                 (1024, 120, 600, 1, 3001), (64, 16, 64, 1, 700), (128, 64, 100, 1, 1000),
                                 (256, 128, 256, 1, 1500), (4096, 1024, 4096, 1, 9001)])
def test_emulated_fused_stft_front_end_matches_numpy(emul_console, cfg):
    """stft_fused.cuh (framing + real FFT of both signals + loss sums) on the host emulator against numpy float64:
    spectrum of the prediction, clamped power of the target, the four sums the loss terms are made of."""
    import ctypes
    n, hop, win, rows, T = cfg
    lib = emul_console.load()
    rng = np.random.default_rng(n + T)
    x = (rng.standard_normal((rows, T)) * 0.1).astype(np.float32)
    y = (rng.standard_normal((rows, T)) * 0.1 + 0.3 * x).astype(np.float32)
    M = n // 2
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win) / win)).astype(np.float32)   # torch.hann_window (periodic)
    j = np.arange(M)
    twm = np.stack([np.cos(2 * np.pi * j / M), -np.sin(2 * np.pi * j / M)], -1).astype(np.float32)
    k = np.arange(M // 2 + 1)
    twn = np.stack([np.cos(2 * np.pi * k / n), -np.sin(2 * np.pi * k / n)], -1).astype(np.float32)
    frames, bins = 1 + T // hop, M + 1
    X = np.zeros((rows, frames, bins, 2), np.float32)
    PY = np.zeros((rows, frames, bins), np.float32)
    fpb = 4096 // M
    partial = np.zeros((rows, (frames + fpb - 1) // fpb, 4), np.float32)
    eps = 1e-8
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.dmst_emul_stft_loss(p(x), p(y), rows, T, n, hop, win, p(w), p(twm), p(twn), p(X), p(PY),
                                 ctypes.c_float(eps), p(partial))
    assert rc == 0
    wp = np.zeros(n)
    wl = (n - win) // 2
    wp[wl:wl + win] = w

    def stft(s):   # torch.stft(center=True, pad_mode="reflect", onesided) as auraloss calls it
        sp = np.pad(s.astype(np.float64), ((0, 0), (n // 2, n // 2)), mode="reflect")
        return np.fft.rfft(np.stack([sp[:, f * hop:f * hop + n] * wp for f in range(frames)], 1), axis=-1)

    Xr, Yr = stft(x), stft(y)
    assert np.abs((X[..., 0] + 1j * X[..., 1]) - Xr).max() <= 5e-7 * np.abs(Xr).max()
    pyr, pxr = np.maximum(np.abs(Yr) ** 2, eps), np.maximum(np.abs(Xr) ** 2, eps)
    assert np.abs(PY - pyr).max() <= 2e-6 * pyr.max()
    d = np.sqrt(pyr) - np.sqrt(pxr)
    ref = np.array([(d ** 2).sum(), pyr.sum(), np.abs(0.5 * (np.log(pxr) - np.log(pyr))).sum(), np.abs(d).sum()])
    got = partial.astype(np.float64).sum((0, 1))
    assert (np.abs(got - ref) / ref).max() <= 2e-6


@pytest.mark.parametrize("cfg", [(512, 2, 20), (2048, 1, 7), (8192, 1, 3), (64, 1, 70), (128, 1, 40), (256, 2, 17),
                                 (1024, 1, 9), (4096, 1, 2)])
@pytest.mark.parametrize("terms", [(1, 1), (0, 0)])
def test_emulated_fused_gradient_inverse_fft_matches_closed_form(emul_console, cfg, terms):
    """stft_fused.cuh::istft_grad_kernel (spectrum + clamped target power -> half-spectrum gradient -> inverse real FFT
    of every frame) on the host emulator against the closed form r[n] = Re sum_k g[k] exp(+2 pi i k n / N) in float64,
    for every radix plan, with and without the log / linear magnitude terms, including a clamped bin."""
    import ctypes
    n, rows, frames = cfg
    use_log, use_lin = terms
    lib = emul_console.load()
    rng = np.random.default_rng(n + frames)
    M = n // 2
    bins = M + 1
    X = rng.standard_normal((rows, frames, bins, 2)).astype(np.float32)
    X[0, 0, 3] = 0   # |X|^2 below eps: the clamp passes no gradient
    Y = rng.standard_normal((rows, frames, bins, 2))
    eps = 1e-8
    PY = np.maximum((Y ** 2).sum(-1), eps).astype(np.float32)
    rc = (rng.random(rows) * 0.01).astype(np.float32)
    scal = np.array([3e-4, 2e-4], np.float32)
    j = np.arange(M)
    twm = np.stack([np.cos(2 * np.pi * j / M), -np.sin(2 * np.pi * j / M)], -1).astype(np.float32)
    k = np.arange(M // 2 + 1)
    twn = np.stack([np.cos(2 * np.pi * k / n), -np.sin(2 * np.pi * k / n)], -1).astype(np.float32)
    out = np.zeros((rows, frames, n), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rcode = lib.dmst_emul_istft_grad(p(X), p(PY), p(out), rows, n, frames, p(twm), p(twn), ctypes.c_float(eps), p(rc),
                                     p(scal), use_log, use_lin)
    assert rcode == 0
    Xc = X[..., 0].astype(np.float64) + 1j * X[..., 1].astype(np.float64)
    px, py = np.abs(Xc) ** 2, PY.astype(np.float64)
    mx, my = np.sqrt(np.maximum(px, 1e-300)), np.sqrt(py)
    gsc = rc.astype(np.float64)[:, None, None] * (mx - my)
    if use_log:
        gsc = gsc + scal[0] * np.sign(px - py) / mx
    if use_lin:
        gsc = gsc + scal[1] * np.sign(mx - my)
    g = np.where(px >= eps, gsc / mx, 0.0) * Xc
    E = np.exp(2j * np.pi * np.outer(np.arange(bins), np.arange(n)) / n)
    ref = np.real(g @ E)
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
