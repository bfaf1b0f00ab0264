"""Generate the golden fixtures in this directory.

Runs ONLY in the build container (needs /root/reference).  It imports the reference's own
modules *unmodified* from /root/reference — ``mst.modules.AdvancedMixConsole``,
``mst.loss.AudioFeatureLoss``, ``mst.filter.barkscale_fbanks`` — with the oracle shims
(oracle/dasp_pytorch, oracle/auraloss, oracle/stubs/librosa) standing in for the
third-party packages that are not installed, evaluates them in float64 on seeded inputs and
stores inputs + outputs as small .npz files.  It also cross-checks the oracle's own
restatements (oracle/console.py, oracle/loss.py) against those reference classes so that the
fixtures pin both.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
# shims first, then the reference tree
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # exposes dasp_pytorch, auraloss
sys.path.insert(0, "/root/reference")

import mst.modules as ref_modules  # noqa: E402  (unmodified reference)
import mst.loss as ref_loss  # noqa: E402
import mst.filter as ref_filter  # noqa: E402
import auraloss  # noqa: E402  (oracle shim)

from oracle.console import OracleAdvancedMixConsole, OracleBasicMixConsole  # noqa: E402
from oracle.loss import OracleAudioFeatureLoss, barkscale_fbanks, batch_stereo_peak_normalize  # noqa: E402

SR = 44100
FLAG_NAMES = ["use_track_input_fader", "use_track_eq", "use_track_compressor",
              "use_track_panner", "use_master_bus", "use_fx_bus", "use_output_fader"]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def console_case(name, bs, n, T, seed, flags):
    g = torch.Generator().manual_seed(seed)
    tracks = (torch.randn(bs, n, T, generator=g) * 0.1).float()
    tp = torch.rand(bs, n, 27, generator=g).float()
    fp = torch.rand(bs, 25, generator=g).float()
    mp = torch.rand(bs, 26, generator=g).float()
    probe = torch.randn(bs, 2, T, generator=g).double()

    tp64 = tp.double().requires_grad_(True)
    mp64 = mp.double().requires_grad_(True)
    x64 = tracks.double().requires_grad_(True)
    ref = ref_modules.AdvancedMixConsole(SR)
    mixed, mix, tpd, fpd, mpd = ref(x64, tp64, fp.double(), mp64, **flags)
    loss = (mix * probe).sum()
    loss.backward()

    # the oracle's own restatement must agree with the reference class to rounding
    orc = OracleAdvancedMixConsole(SR)
    omixed, omix, otpd, _, ompd = orc(tracks.double(), tp.double(), fp.double(), mp.double(), **flags)
    assert torch.allclose(omix, mix.detach(), rtol=0, atol=1e-12 * float(mix.abs().max())), name
    assert torch.allclose(omixed, mixed.detach(), rtol=0, atol=1e-12 * float(mixed.abs().max()))
    for eff in tpd:
        for k in tpd[eff]:
            assert torch.equal(otpd[eff][k], tpd[eff][k].detach())

    save(name,
         tracks=tracks.numpy(), track_params=tp.numpy(), fx_bus_params=fp.numpy(),
         master_bus_params=mp.numpy(), probe=probe.float().numpy(),
         flags=np.array([int(flags[k]) for k in FLAG_NAMES], dtype=np.int32),
         mix=mix.detach().numpy().astype(np.float32),
         mixed_energy=(mixed.detach() ** 2).sum(-1).numpy(),
         mixed_head=mixed.detach()[..., :1024].numpy().astype(np.float32),
         grad_track_params=tp64.grad.numpy(),
         grad_master_bus_params=(mp64.grad if mp64.grad is not None
                                 else torch.zeros_like(mp64)).detach().numpy(),
         grad_tracks_head=x64.grad[..., :2048].numpy().astype(np.float32),
         grad_tracks_energy=(x64.grad ** 2).sum(-1).numpy(),
         denorm_threshold_db=tpd["compressor"]["threshold_db"].detach().numpy(),
         denorm_band3_cutoff=tpd["parametric_eq"]["band3_cutoff_freq"].detach().numpy(),
         loss=np.float64(loss.item()))


def main():
    all_on = dict(use_track_input_fader=True, use_track_eq=True, use_track_compressor=True,
                  use_track_panner=True, use_master_bus=True, use_fx_bus=False,
                  use_output_fader=True)
    console_case("console_adv_all", 1, 3, 32768, 1, all_on)
    console_case("console_adv_b2", 2, 2, 16384 + 8192, 2, all_on)  # T not a power of two
    console_case("console_adv_train_flags", 1, 2, 32768, 3,
                 dict(all_on, use_track_input_fader=False, use_output_fader=False))
    console_case("console_adv_eq_only", 1, 2, 32768, 4,
                 dict(all_on, use_track_compressor=False, use_master_bus=False))
    console_case("console_adv_comp_only", 1, 2, 32768, 5,
                 dict(all_on, use_track_eq=False, use_master_bus=False, use_output_fader=False))
    console_case("console_adv_gainpan_only", 1, 4, 44100, 6,
                 dict(all_on, use_track_eq=False, use_track_compressor=False,
                      use_master_bus=False, use_output_fader=False))

    # ---- BasicMixConsole (reconstruction; BASELINE configs[0] shape 1 x 4 x 44100) ----
    g = torch.Generator().manual_seed(7)
    tracks = (torch.randn(1, 4, 44100, generator=g) * 0.1).float()
    tp = torch.rand(1, 4, 2, generator=g).float()
    basic = OracleBasicMixConsole(SR)
    # float32 on purpose: gain/pan/bus is the "bit-exact" tier, evaluated in the
    # reference's own dtype and op order.
    mixed32, mix32, tpd32, _, _ = basic(tracks, tp)
    mixed64, mix64, _, _, _ = basic(tracks.double(), tp.double())
    save("console_basic", tracks=tracks.numpy(), track_params=tp.numpy(),
         mix_f32=mix32.numpy(), mixed_f32_head=mixed32[..., :2048].numpy(), mix_f64=mix64.numpy(),
         gain_db=tpd32["input_gain"]["gain_db"].numpy(), pan=tpd32["stereo_panner"]["pan"].numpy())

    # ---- MRSTFT (auraloss shim; configs/models/naive.yaml:54-68) ----
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(2, 2, 20000, generator=g) * 0.1).float()
    y = (torch.randn(2, 2, 20000, generator=g) * 0.1 + 0.5 * x).float()
    for tag, kw in (("train", dict(w_sc=1.0, w_log_mag=1.0, w_lin_mag=0.0)),
                    ("eval", dict(w_sc=0.0, w_log_mag=1.0, w_lin_mag=1.0))):
        fn = auraloss.freq.MultiResolutionSTFTLoss(
            fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096],
            win_lengths=[512, 2048, 8192], **kw)
        x64 = x.double().requires_grad_(True)
        val = fn(x64, y.double())
        val.backward()
        save(f"mrstft_{tag}", x=x.numpy(), y=y.numpy(), loss=np.float64(val.item()),
             grad_x=x64.grad.numpy().astype(np.float32),
             weights=np.array([kw["w_sc"], kw["w_log_mag"], kw["w_lin_mag"]]))

    # ---- AudioFeatureLoss: the reference class itself (fp32, its native dtype) and the
    #      oracle restatement in fp64 ----
    g = torch.Generator().manual_seed(9)
    a = (torch.randn(2, 2, 40000, generator=g) * 0.1).float()
    a[:, 1] = 0.6 * a[:, 0] + 0.4 * a[:, 1]
    b = (torch.randn(2, 2, 40000, generator=g) * 0.05).float()
    b[:, 0] = 0.3 * b[:, 1] + 0.7 * b[:, 0]
    weights = [0.1, 0.001, 1.0, 1.0, 0.1]
    ref_afl = ref_loss.AudioFeatureLoss(weights, SR)
    ref_vals = ref_afl(a, b)
    orc_afl = OracleAudioFeatureLoss(weights, SR)
    a64 = a.double().requires_grad_(True)
    orc_vals = orc_afl(a64, b.double())
    total = sum(v.mean() for v in orc_vals.values())
    total.backward()
    assert list(ref_vals.keys()) == list(orc_vals.keys()), (ref_vals.keys(), orc_vals.keys())
    for k in ref_vals:
        r, o = float(ref_vals[k]), float(orc_vals[k])
        assert abs(r - o) <= 2e-4 * max(abs(o), 1e-12) + 1e-9, (k, r, o)
    save("afl", input=a.numpy(), target=b.numpy(), weights=np.array(weights),
         keys=np.array(list(orc_vals.keys())),
         values=np.array([float(v) for v in orc_vals.values()]),
         ref_values_f32=np.array([float(v) for v in ref_vals.values()]),
         grad_input=a64.grad.numpy().astype(np.float32), total=np.float64(total.item()))

    # ---- bark filterbank: reference function vs oracle restatement ----
    fb_ref = ref_filter.barkscale_fbanks(16385, 20.0, 20000.0, 24, SR)
    fb_orc = barkscale_fbanks(16385, 20.0, 20000.0, 24, SR)
    assert torch.equal(fb_ref, fb_orc)
    save("bark_fb", col_sums=fb_ref.sum(0).numpy(), row_sums_head=fb_ref.sum(1)[:4096].numpy(),
         argmax=fb_ref.argmax(0).numpy(), shape=np.array(fb_ref.shape))

    # ---- peak normalise (mst/utils.py:14-29; pyloudnorm blocks importing mst.utils here) ----
    g = torch.Generator().manual_seed(10)
    m = torch.randn(3, 2, 4096, generator=g).float()
    m[2] = 0
    save("peaknorm", x=m.numpy(), y=batch_stereo_peak_normalize(m).numpy())
    print("golden fixtures written")


if __name__ == "__main__":
    main()
