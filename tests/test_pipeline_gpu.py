"""GPU parity of whole training steps in the shapes of BASELINE.json's configs[2] and configs[3] (at a
reduced length the float64 oracle finishes in seconds): console forward -> loss -> backward to the console
parameters, CUDA path (through the C ABI) against the float64 oracle of the same step.  These cover the
composition the per-component tests do not: the loss gradients entering the console's backward kernels.

Tolerances: 1e-4 relative on the loss, 1e-3 relative L2 on the gradients, or the distance of the reference
algorithm's own float32 evaluation from float64 on the same inputs where that is larger."""
import numpy as np
import pytest
import torch

from oracle.auraloss.freq import MultiResolutionSTFTLoss as OracleMRSTFT
from oracle.console import OracleAdvancedMixConsole
from oracle.loss import OracleAudioFeatureLoss, batch_stereo_peak_normalize as oracle_peak_normalize

pytestmark = pytest.mark.gpu
SR = 44100
RES = dict(fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096], win_lengths=[512, 2048, 8192])
# training flags of the shipped configs (configs/models/naive.yaml:5-8, mst/system.py:83-89)
FLAGS = dict(use_track_input_fader=True, use_track_eq=True, use_track_compressor=True, use_track_panner=True,
             use_master_bus=True, use_fx_bus=False, use_output_fader=True)
AFL_WEIGHTS = [0.1, 0.001, 1.0, 1.0, 0.1]  # configs/losses/feat.yaml order: rms, crest, width, imbalance, bark


def rell2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_step_inputs(B, N, T, seed):
    g = torch.Generator().manual_seed(seed)
    tracks = torch.randn(B, N, T, generator=g) * 0.1
    tp, fp, mp = torch.rand(B, N, 27, generator=g), torch.rand(B, 25, generator=g), torch.rand(B, 26, generator=g)
    tp2, mp2 = torch.rand(B, N, 27, generator=g), torch.rand(B, 26, generator=g)  # the reference mix's parameters
    return tracks, tp, fp, mp, tp2, mp2


def oracle_step(inputs, loss_fn, dtype):
    """mst/system.py:232-253 (random reference mix, peak-normalised, no grad) + :280-338 (predicted mix, loss)."""
    tracks, tp, fp, mp, tp2, mp2 = (t.to(dtype) for t in inputs)
    con = OracleAdvancedMixConsole(SR)
    with torch.no_grad():
        target = oracle_peak_normalize(con(tracks, tp2, fp, mp2, **FLAGS)[1])
    tp = tp.clone().requires_grad_(True); mp = mp.clone().requires_grad_(True)
    loss = loss_fn(con(tracks, tp, fp, mp, **FLAGS)[1], target)
    loss.backward()
    return float(loss.detach()), tp.grad.numpy(), mp.grad.numpy()


def our_step(inputs, loss_fn):
    from diffmst_b200 import AdvancedMixConsole, batch_stereo_peak_normalize
    tracks, tp, fp, mp, tp2, mp2 = (t.cuda() for t in inputs)
    con = AdvancedMixConsole(SR).cuda()
    con.materialize_tracks = False
    with torch.no_grad():
        target = batch_stereo_peak_normalize(con(tracks, tp2, fp, mp2, **FLAGS)[1])
    tp = tp.clone().requires_grad_(True); mp = mp.clone().requires_grad_(True)
    loss = loss_fn(con(tracks, tp, fp, mp, **FLAGS)[1], target)
    loss.backward()
    return float(loss.detach()), tp.grad.cpu().numpy(), mp.grad.cpu().numpy()


def check_step(ours, o64, o32):
    assert abs(ours[0] - o64[0]) <= max(1e-4, 1.5 * abs(o32[0] - o64[0]) / abs(o64[0])) * abs(o64[0]), (ours[0], o64[0], o32[0])
    for i, name in ((1, "track_params"), (2, "master_bus_params")):
        bound = max(1e-3, 1.5 * rell2(o32[i], o64[i]))
        assert np.isfinite(ours[i]).all()
        assert rell2(ours[i], o64[i]) <= bound, (name, rell2(ours[i], o64[i]), bound)


def test_naive_mix_mrstft_step_matches_float64_oracle():
    """configs[3]: Method-1 step (configs/models/naive.yaml) - 8 tracks, random reference mix, MRSTFT loss."""
    from diffmst_b200 import MRSTFTLoss
    inputs = make_step_inputs(2, 8, 65536, seed=31)
    ours = our_step(inputs, MRSTFTLoss(**RES))
    o64 = oracle_step(inputs, OracleMRSTFT(**RES), torch.float64)
    o32 = oracle_step(inputs, OracleMRSTFT(**RES), torch.float32)
    check_step(ours, o64, o32)


def test_audio_feature_loss_step_32_tracks_matches_float64_oracle():
    """configs[2]: 32 tracks + AudioFeatureLoss; the dict is reduced as System.common_step does
    (mst/system.py:334-338: sum of the terms' means)."""
    from diffmst_b200 import AudioFeatureLoss
    inputs = make_step_inputs(1, 32, 65536, seed=32)
    reduce_dict = lambda f: (lambda a, b: sum(v.mean() for v in f(a, b).values()))
    ours = our_step(inputs, reduce_dict(AudioFeatureLoss(AFL_WEIGHTS, SR)))
    o64 = oracle_step(inputs, reduce_dict(OracleAudioFeatureLoss(AFL_WEIGHTS, SR)), torch.float64)
    o32 = oracle_step(inputs, reduce_dict(OracleAudioFeatureLoss(AFL_WEIGHTS, SR)), torch.float32)
    check_step(ours, o64, o32)
