"""GPU parity tests of the loss stack (through the C ABI) against the float64 oracle and the
golden fixtures.  Tolerances: 1e-4 relative on loss values (north_star's float tolerance),
1e-3 relative L2 on gradients (float32 FFTs on both sides of the comparison)."""
import math

import numpy as np
import pytest
import torch

from oracle.auraloss.freq import MultiResolutionSTFTLoss as OracleMRSTFT
from oracle.loss import OracleAudioFeatureLoss

pytestmark = pytest.mark.gpu
SR = 44100
RES = dict(fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096], win_lengths=[512, 2048, 8192])


def rell2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("tag", ["train", "eval"])
def test_mrstft_golden(golden, tag):
    from diffmst_b200 import MultiResolutionSTFTLoss
    d = golden(f"mrstft_{tag}")
    w = d["weights"]
    f = MultiResolutionSTFTLoss(**RES, w_sc=float(w[0]), w_log_mag=float(w[1]), w_lin_mag=float(w[2]))
    x = torch.from_numpy(d["x"]).cuda().requires_grad_(True)
    loss = f(x, torch.from_numpy(d["y"]).cuda())
    assert loss.dim() == 0
    assert abs(float(loss) - float(d["loss"])) <= 1e-4 * abs(float(d["loss"]))
    loss.backward()
    # float32 FFTs bound the gradient's distance from float64 (1/|X| weights amplify the FFT's
    # rounding in weak bins): calibrate with the reference algorithm itself run in float32 on the
    # same device (the shim is plain torch, so this is torch.stft -> cuFFT), and require that
    # we sit on top of that evaluation.
    o = OracleMRSTFT(**RES, w_sc=float(w[0]), w_log_mag=float(w[1]), w_lin_mag=float(w[2]))
    xr = torch.from_numpy(d["x"]).cuda().requires_grad_(True)
    o(xr, torch.from_numpy(d["y"]).cuda()).backward()
    ref32 = rell2(xr.grad.cpu().numpy(), d["grad_x"])
    ours = rell2(x.grad.cpu().numpy(), d["grad_x"])
    print(f"mrstft_{tag}: gradient distance from float64: ours {ours:.3e}, torch float32 {ref32:.3e}")
    assert ours <= max(1e-3, 1.2 * ref32)
    # (the fused front end computes its own FFT: two float32 evaluations differ from each other by about as much as
    # each differs from float64, so the comparison with torch's float32 result is bounded by the sum of the two)
    assert rell2(x.grad.cpu().numpy(), xr.grad.cpu().numpy()) <= ours + ref32 + 1e-6


def test_mrstft_library_fft_path():
    """FFT sizes the fused front end does not serve (not a power of two) go through framing + cuFFT + the loss
    kernel: same parity bar against the float64 oracle."""
    from diffmst_b200 import MRSTFTLoss
    res = dict(fft_sizes=[1000, 600], hop_sizes=[250, 150], win_lengths=[1000, 400])
    g = torch.Generator().manual_seed(24)
    x = torch.randn(2, 2, 30000, generator=g) * 0.1
    y = torch.randn(2, 2, 30000, generator=g) * 0.1 + 0.4 * x
    f, o = MRSTFTLoss(**res), OracleMRSTFT(**res)
    xc = x.cuda().requires_grad_(True)
    loss = f(xc, y.cuda())
    x64 = x.double().requires_grad_(True)
    ref = o(x64, y.double())
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    loss.backward(); ref.backward()
    assert rell2(xc.grad.cpu().numpy(), x64.grad.numpy()) <= 3e-3


def test_mrstft_vs_oracle_strided_and_scaled():
    from diffmst_b200 import MRSTFTLoss
    g = torch.Generator().manual_seed(21)
    full_x = torch.randn(3, 2, 50000, generator=g) * 0.1
    full_y = torch.randn(3, 2, 50000, generator=g) * 0.1 + 0.3 * full_x
    x, y = full_x[..., 10000:], full_y[..., 10000:]  # strided views, as mst/system.py:256-258 makes
    f = MRSTFTLoss(**RES)
    o = OracleMRSTFT(**RES)
    xc = full_x.cuda()[..., 10000:].requires_grad_(True)
    loss = f(xc, full_y.cuda()[..., 10000:])
    x64 = x.double().requires_grad_(True)
    ref = o(x64, y.double())
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    (3.0 * loss).backward()
    ref.backward()
    assert rell2(xc.grad.cpu().numpy(), 3.0 * x64.grad.numpy()) <= 3e-3  # float32 FFT floor, see above
    # non-default window lengths (win < fft) and the default auraloss resolutions
    f2, o2 = MRSTFTLoss(), OracleMRSTFT()
    a = float(f2(x.contiguous().cuda(), y.contiguous().cuda()))
    b = float(o2(x.double(), y.double()))
    assert abs(a - b) <= 1e-4 * abs(b)


def test_mrstft_gradient_general_layouts():
    """Gradient parity where the fast paths do not apply: window shorter than the FFT and hops that are not
    half the FFT (auraloss defaults 1024/2048/512, hops 120/240/50, windows 600/1200/240), an odd length, and
    views that are only 4-byte aligned."""
    from diffmst_b200 import MRSTFTLoss
    g = torch.Generator().manual_seed(23)
    full_x = torch.randn(2, 2, 30003, generator=g) * 0.1
    full_y = torch.randn(2, 2, 30003, generator=g) * 0.1 + 0.5 * full_x
    for res in (dict(), RES):
        f, o = MRSTFTLoss(**res), OracleMRSTFT(**res)
        xc = full_x.cuda()[..., 10001:].requires_grad_(True)      # odd offset, odd length (20002 samples)
        loss = f(xc, full_y.cuda()[..., 10001:])
        x64 = full_x[..., 10001:].double().requires_grad_(True)
        ref = o(x64, full_y[..., 10001:].double())
        assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
        loss.backward(); ref.backward()
        assert rell2(xc.grad.cpu().numpy(), x64.grad.numpy()) <= 3e-3


def test_mrstft_closed_forms_and_errors():
    from diffmst_b200 import MRSTFTLoss
    g = torch.Generator().manual_seed(22)
    y = (torch.randn(2, 2, 20000, generator=g) * 0.1).cuda()
    f = MRSTFTLoss(**RES)
    assert float(f(y, y)) == 0.0
    c = 0.5
    assert abs(float(f(c * y, y)) - (abs(1 - c) + abs(math.log(c)))) < 1e-4
    # both terms are invariant to a common gain
    assert abs(float(f(4 * (0.7 * y), 4 * y)) - float(f(0.7 * y, y))) < 1e-5
    z = torch.zeros(1, 2, 20000).cuda()
    assert abs(float(MRSTFTLoss(**RES, w_sc=0.0, w_log_mag=0.0, w_lin_mag=1.0)(z, z))) == 0.0
    with pytest.raises(NotImplementedError):
        MRSTFTLoss(perceptual_weighting=True)
    with pytest.raises(ValueError):
        f(y[..., :3000], y[..., :3000])  # reflect padding of the 8192 resolution needs > 4096 samples
    with pytest.raises(RuntimeError, match="no CPU path"):
        f(y.cpu(), y.cpu())


def test_afl_golden(golden):
    from diffmst_b200 import AudioFeatureLoss
    d = golden("afl")
    f = AudioFeatureLoss([float(v) for v in d["weights"]], SR)
    x = torch.from_numpy(d["input"]).cuda().requires_grad_(True)
    out = f(x, torch.from_numpy(d["target"]).cuda())
    assert list(out.keys()) == list(d["keys"])
    got = np.array([float(v) for v in out.values()])
    assert np.allclose(got, d["values"], rtol=2e-4), (got, d["values"])
    total = sum(v.mean() for v in out.values())  # mst/system.py:334-336
    assert abs(float(total) - float(d["total"])) <= 2e-4 * abs(float(d["total"]))
    total.backward()
    assert rell2(x.grad.cpu().numpy(), d["grad_input"]) <= 2e-3


def test_afl_vs_oracle_per_term_gradients():
    from diffmst_b200 import AudioFeatureLoss
    g = torch.Generator().manual_seed(23)
    a = torch.randn(3, 2, 36000, generator=g) * 0.1
    a[:, 1] = 0.5 * a[:, 0] + 0.5 * a[:, 1]
    b = torch.randn(3, 2, 36000, generator=g) * 0.07
    weights = [0.1, 0.001, 1.0, 1.0, 0.1]
    ours, orc = AudioFeatureLoss(weights, SR), OracleAudioFeatureLoss(weights, SR)
    keys = ["mix-rms", "mix-crest_factor", "mix-stereo_width", "mix-stereo_imbalance", "mix-barkspectrum"]
    for k in keys:
        xc = a.cuda().requires_grad_(True)
        x64 = a.double().requires_grad_(True)
        vo, vr = ours(xc, b.cuda())[k], orc(x64, b.double())[k]
        assert abs(float(vo) - float(vr)) <= 2e-4 * abs(float(vr)) + 1e-12, k
        vo.backward(); vr.backward()
        assert rell2(xc.grad.cpu().numpy(), x64.grad.numpy()) <= 2e-3, k


def test_peak_normalize_golden(golden):
    from diffmst_b200 import batch_stereo_peak_normalize
    d = golden("peaknorm")
    y = batch_stereo_peak_normalize(torch.from_numpy(d["x"]).cuda())
    assert np.array_equal(y.cpu().numpy(), d["y"])


def test_loss_full_size_properties():
    """BASELINE configs[1] loss size (8, 2, 262144): finite, deterministic, zero at equality."""
    from diffmst_b200 import MRSTFTLoss, AudioFeatureLoss
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(8, 2, 262144, generator=g) * 0.1).cuda().requires_grad_(True)
    y = (torch.randn(8, 2, 262144, generator=g) * 0.1).cuda()
    f = MRSTFTLoss(**RES)
    l1 = f(x, y); l1.backward(); g1 = x.grad.clone(); x.grad = None
    l2 = f(x, y); l2.backward()
    assert torch.isfinite(l1) and torch.equal(l1, l2) and torch.equal(g1, x.grad)
    assert float(f(y, y)) == 0.0
    afl = AudioFeatureLoss([0.1, 0.001, 1.0, 1.0, 0.1], SR)
    out = afl(x, y)
    assert all(torch.isfinite(v) for v in out.values())
    same = afl(y, y)
    assert all(float(v) == 0.0 for v in same.values())


def test_mrstft_one_call_form_of_the_c_abi_matches_the_autograd_node():
    """dmst_mrstft_forward with a gradient buffer (loss and d loss / d x in ONE call, what a non-autograd host would
    bind) against the two-call form the autograd node uses (dmst_mrstft_forward_keep + dmst_mrstft_backward)."""
    import ctypes
    from diffmst_b200 import MRSTFTLoss, _lib
    from diffmst_b200.console import _ptr
    g = torch.Generator().manual_seed(25)
    x = (torch.randn(2, 2, 30000, generator=g) * 0.1).cuda()
    y = (torch.randn(2, 2, 30000, generator=g) * 0.1 + 0.3 * x.cpu()).cuda()
    f = MRSTFTLoss(**RES)
    xr = x.clone().requires_grad_(True)
    loss = f(xr, y)
    loss.backward()
    lib, cfg, win = _lib.lib(), f._cfg(), f._windows_on(x.device)
    rows, T = 4, 30000
    n = lib.dmst_mrstft_workspace_bytes(ctypes.byref(cfg), rows, T)
    ws = torch.empty(n, dtype=torch.uint8, device=x.device)
    out = torch.empty(1 + 3 * cfg.n_res, device=x.device)
    gx = torch.empty(rows, T, device=x.device)
    rc = lib.dmst_mrstft_forward(_ptr(x), T, _ptr(y), T, _ptr(win), ctypes.byref(cfg), rows, T, _ptr(out), _ptr(gx),
                                 _ptr(ws), n, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    assert float(out[0]) == float(loss.detach())
    assert torch.equal(gx.view_as(x), xr.grad)
    assert torch.equal(out[1:].view(-1, 3), f.last_terms)
