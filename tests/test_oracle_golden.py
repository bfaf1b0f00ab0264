"""The oracle restatements against the golden vectors written by
tests/golden/make_golden.py (which ran the reference's own classes, imported unmodified
from /root/reference, in float64)."""
import numpy as np
import torch

from oracle.console import OracleAdvancedMixConsole, OracleBasicMixConsole
from oracle.loss import OracleAudioFeatureLoss, barkscale_fbanks, batch_stereo_peak_normalize
from oracle.auraloss.freq import MultiResolutionSTFTLoss
from oracle import timedomain as td

SR = 44100
FLAG_NAMES = ["use_track_input_fader", "use_track_eq", "use_track_compressor",
              "use_track_panner", "use_master_bus", "use_fx_bus", "use_output_fader"]
CASES = ["console_adv_all", "console_adv_b2", "console_adv_train_flags", "console_adv_eq_only",
         "console_adv_comp_only", "console_adv_gainpan_only"]


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def test_console_oracle_matches_golden(golden):
    for name in CASES:
        d = golden(name)
        flags = {k: bool(v) for k, v in zip(FLAG_NAMES, d["flags"])}
        tp = torch.from_numpy(d["track_params"]).double().requires_grad_(True)
        mp = torch.from_numpy(d["master_bus_params"]).double().requires_grad_(True)
        con = OracleAdvancedMixConsole(SR)
        mixed, mix, tpd, _, _ = con(torch.from_numpy(d["tracks"]).double(), tp,
                                    torch.from_numpy(d["fx_bus_params"]).double(), mp, **flags)
        assert _relmax(mix.detach().numpy(), d["mix"]) < 2e-7, name  # fixture stored as fp32
        assert np.allclose((mixed.detach() ** 2).sum(-1).numpy(), d["mixed_energy"], rtol=1e-10)
        (mix * torch.from_numpy(d["probe"]).double()).sum().backward()
        assert _relmax(tp.grad.numpy(), d["grad_track_params"]) < 1e-6, name
        g = mp.grad.numpy() if mp.grad is not None else np.zeros_like(d["grad_master_bus_params"])
        assert _relmax(g, d["grad_master_bus_params"]) < 1e-6 or np.abs(d["grad_master_bus_params"]).max() == 0
        assert np.array_equal(tpd["compressor"]["threshold_db"].detach().numpy(), d["denorm_threshold_db"])


def test_time_domain_oracle_matches_fsm_golden(golden):
    # independent direction: exact recursion (scipy) vs the FSM golden, T = 32768
    d = golden("console_adv_all")
    con = OracleAdvancedMixConsole(SR)
    _, _, tpd, _, mpd = con(torch.from_numpy(d["tracks"]).double(),
                            torch.from_numpy(d["track_params"]).double(),
                            torch.from_numpy(d["fx_bus_params"]).double(),
                            torch.from_numpy(d["master_bus_params"]).double(), use_fx_bus=False)
    from oracle.console import EQ_KEYS, COMP_KEYS
    tp = np.zeros((3, 27))
    tp[:, 0] = tpd["input_fader"]["gain_db"][0].numpy()
    for i, k in enumerate(EQ_KEYS):
        tp[:, 1 + i] = tpd["parametric_eq"][k][0].numpy()
    for i, k in enumerate(COMP_KEYS):
        tp[:, 19 + i] = tpd["compressor"][k][0].numpy()
    tp[:, 25] = tpd["stereo_panner"]["pan"][0].numpy()
    mpv = np.zeros(26)
    for i, k in enumerate(EQ_KEYS):
        mpv[i] = mpd["parametric_eq"][k][0].item()
    for i, k in enumerate(COMP_KEYS):
        mpv[18 + i] = mpd["compressor"][k][0].item()
    mpv[24] = mpd["output_fader"]["gain_db"][0].item()
    mpv[25] = mpd["input_fader"]["gain_db"][0].item()
    mixed, mix = td.console(d["tracks"][0].astype(np.float64), tp, mpv, SR)
    # FSM time-aliasing at T=32768 bounds the agreement
    assert _relmax(mix, d["mix"][0]) < 5e-5


def test_basic_console_golden(golden):
    d = golden("console_basic")
    con = OracleBasicMixConsole(SR)
    mixed, mix, tpd, _, _ = con(torch.from_numpy(d["tracks"]), torch.from_numpy(d["track_params"]))
    assert np.array_equal(mix.numpy(), d["mix_f32"])
    assert np.array_equal(mixed[..., :2048].numpy(), d["mixed_f32_head"])
    assert con.num_track_control_params == 2


def test_mrstft_oracle_golden(golden):
    for tag in ("train", "eval"):
        d = golden(f"mrstft_{tag}")
        w = d["weights"]
        f = MultiResolutionSTFTLoss([512, 2048, 8192], [256, 1024, 4096], [512, 2048, 8192],
                                    w_sc=float(w[0]), w_log_mag=float(w[1]), w_lin_mag=float(w[2]))
        x = torch.from_numpy(d["x"]).double().requires_grad_(True)
        val = f(x, torch.from_numpy(d["y"]).double())
        assert abs(float(val) - float(d["loss"])) < 1e-12
        val.backward()
        assert _relmax(x.grad.numpy(), d["grad_x"]) < 1e-6


def test_afl_oracle_golden(golden):
    d = golden("afl")
    f = OracleAudioFeatureLoss(list(d["weights"]), SR)
    vals = f(torch.from_numpy(d["input"]).double(), torch.from_numpy(d["target"]).double())
    assert list(vals.keys()) == list(d["keys"])
    assert list(vals.keys()) == ["mix-rms", "mix-crest_factor", "mix-stereo_width",
                                 "mix-stereo_imbalance", "mix-barkspectrum"]
    got = np.array([float(v) for v in vals.values()])
    assert np.allclose(got, d["values"], rtol=1e-12)
    # the reference class itself (fp32) agrees with the fp64 oracle to fp32 rounding
    assert np.allclose(d["ref_values_f32"], d["values"], rtol=2e-4)


def test_bark_and_peaknorm_golden(golden):
    d = golden("bark_fb")
    fb = barkscale_fbanks(16385, 20.0, 20000.0, 24, SR)
    assert tuple(d["shape"]) == tuple(fb.shape) == (16385, 24)
    assert np.array_equal(fb.sum(0).numpy(), d["col_sums"])
    assert np.array_equal(fb.argmax(0).numpy(), d["argmax"])
    d = golden("peaknorm")
    assert np.array_equal(batch_stereo_peak_normalize(torch.from_numpy(d["x"])).numpy(), d["y"])


def test_panns_oracle_matches_reference_golden(golden):
    """oracle/panns.py against the reference's own mst.panns.ConvBlock / Cnn14 (tests/golden/make_golden_panns.py,
    float64, generator-free parameter fill of tests/golden/panns_fill.py): the a11 row's oracle is pinned."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from panns_fill import fill_state
    from oracle.panns import OracleCnn14, OracleConvBlock
    d = golden("panns")
    blk = OracleConvBlock(3, 8).double().train()
    fill_state(blk)
    x = torch.from_numpy(d["block_x"]).requires_grad_(True)
    y = blk(x, (2, 2))
    (y * torch.from_numpy(d["block_probe"])).sum().backward()
    assert _relmax(y.detach().numpy(), d["block_y"]) <= 1e-12
    assert _relmax(x.grad.numpy(), d["block_gx"]) <= 1e-11
    for n, p in blk.named_parameters():
        assert _relmax(p.grad.numpy(), d["block_g_" + n.replace(".", "_")]) <= 1e-10, n
    assert _relmax(blk.bn1.running_mean.numpy(), d["block_running_mean1"]) <= 1e-12   # training-mode bookkeeping
    assert _relmax(blk.bn1.running_var.numpy(), d["block_running_var1"]) <= 1e-12
    net = OracleCnn14(num_classes=6).double().eval()
    fill_state(net)
    g = torch.Generator().manual_seed(42)
    xin = (torch.rand(1, 1, 1024, 128, generator=g) ** 2).double()
    out = net(xin)
    out.square().mean().backward()
    assert _relmax(out.detach().numpy(), d["cnn14_out"]) <= 1e-10
    grads = dict(net.named_parameters())
    for i, n in enumerate(d["cnn14_grad_names"]):
        gsum, gabs = float(grads[str(n)].grad.sum()), float(grads[str(n)].grad.abs().sum())
        assert abs(gabs - d["cnn14_grad_abs"][i]) <= 1e-9 * d["cnn14_grad_abs"][i], n
        assert abs(gsum - d["cnn14_grad_sum"][i]) <= 1e-9 * d["cnn14_grad_abs"][i], n
    assert _relmax(grads["conv_block1.conv1.weight"].grad.numpy(), d["cnn14_g_first"]) <= 1e-9
    assert _relmax(grads["fc.weight"].grad.numpy(), d["cnn14_g_fc"]) <= 1e-10


def test_spectrogram_front_end_formula_golden(golden):
    """The spectrogram lines of the reference's SpectrogramEncoder.forward (mst/modules.py:787-800; golden written by
    the reference class with its trunk replaced by an identity): STFT 2048 / 512, periodic Hann, centred with reflect
    padding, (|X| + 1e-8)^0.3, layout (bs, chs, bins, frames).  This is the formula dmst_spectrogram_frontend
    implements (GPU test: test_conv_gpu.py::test_spectrogram_frontend_matches_torch_stft_composition)."""
    d = golden("panns")
    w = torch.from_numpy(d["spec_wave"])
    X = torch.stft(w.view(-1, w.shape[-1]), n_fft=2048, hop_length=512, window=torch.hann_window(2048), return_complex=True)
    S = torch.pow(X.abs().view(w.shape[0], w.shape[1], 1025, -1) + 1e-8, 0.3)
    assert S.shape == d["spec_out"].shape == (2, 1, 1025, 33)
    assert _relmax(S.numpy(), d["spec_out"]) <= 1e-6
