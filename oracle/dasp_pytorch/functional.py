"""ORACLE (test infrastructure, not product code).

CPU restatement of the ``dasp_pytorch.functional`` entry points the reference imports at
mst/modules.py:7-14 (third-party ``dasp-pytorch==0.0.1``, requirements.txt:15; not
vendored, not installable here).  Formulas follow the published upstream algorithm as
recorded in SURVEY.md Appendix A; in-tree corroboration: release time unused
(mst/modules.py:377, tests/test_comp.py:27), panner output (bs, 2, N, T)
(mst/modules.py:272), six 3-parameter EQ sections (mst/modules.py:124-143).
PARITY UNPINNED by upstream tests; pinned by analytic KATs (tests/test_oracle_kats.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package.
"""
import math

import torch

from . import signal as _signal


def gain(x, sample_rate: float, gain_db):
    """x (bs, chs, T); gain_db (bs,) in dB."""
    bs = x.shape[0]
    return x * (10 ** (gain_db.reshape(bs, 1, 1) / 20.0))


def stereo_panner(x, sample_rate: float, pan):
    """x (bs, N, T), pan (bs, N) in [0, 1] -> (bs, 2, N, T); -4.5 dB centre law."""
    bs, num_tracks, seq_len = x.shape
    theta = pan * (math.pi / 2)
    left_gain = torch.sqrt(((math.pi / 2) - theta) * (2 / math.pi) * torch.cos(theta))
    right_gain = torch.sqrt(theta * (2 / math.pi) * torch.sin(theta))
    x = x.unsqueeze(1).repeat(1, 2, 1, 1)
    gains = torch.cat(
        (left_gain.view(bs, 1, num_tracks, 1), right_gain.view(bs, 1, num_tracks, 1)), dim=1
    )
    x = x * gains
    return x


def stereo_bus(x, sample_rate: float, send_db):
    """x (bs, 2, N, T), send_db (bs, N) -> (bs, 2, T)."""
    bs, chs, num_tracks, seq_len = x.shape
    sends = 10 ** (send_db.view(bs, 1, num_tracks, 1) / 20.0)
    return (x * sends).sum(dim=2)


def parametric_eq(
    x,
    sample_rate: float,
    low_shelf_gain_db,
    low_shelf_cutoff_freq,
    low_shelf_q_factor,
    band0_gain_db,
    band0_cutoff_freq,
    band0_q_factor,
    band1_gain_db,
    band1_cutoff_freq,
    band1_q_factor,
    band2_gain_db,
    band2_cutoff_freq,
    band2_q_factor,
    band3_gain_db,
    band3_cutoff_freq,
    band3_q_factor,
    high_shelf_gain_db,
    high_shelf_cutoff_freq,
    high_shelf_q_factor,
):
    """Six-section RBJ cascade [low_shelf, peaking x4, high_shelf] applied by FSM."""
    bs = x.shape[0]
    spec = [
        (low_shelf_gain_db, low_shelf_cutoff_freq, low_shelf_q_factor, "low_shelf"),
        (band0_gain_db, band0_cutoff_freq, band0_q_factor, "peaking"),
        (band1_gain_db, band1_cutoff_freq, band1_q_factor, "peaking"),
        (band2_gain_db, band2_cutoff_freq, band2_q_factor, "peaking"),
        (band3_gain_db, band3_cutoff_freq, band3_q_factor, "peaking"),
        (high_shelf_gain_db, high_shelf_cutoff_freq, high_shelf_q_factor, "high_shelf"),
    ]
    sections = []
    for g, f, q, kind in spec:
        b, a = _signal.biquad(
            g.reshape(bs).type_as(x), f.reshape(bs).type_as(x), q.reshape(bs).type_as(x),
            sample_rate, kind,
        )
        sections.append(torch.cat([b, a], dim=-1))
    sos = torch.stack(sections, dim=1)  # (bs, 6, 6)
    return _signal.sosfilt_via_fsm(sos, x)


def compressor_gain_computer(x_db, threshold_db, ratio, knee_db):
    """Soft-knee static curve minus input level (Appendix A)."""
    x_sc = x_db.clone()
    lo = threshold_db - knee_db / 2
    hi = threshold_db + knee_db / 2
    in_knee = torch.logical_and(x_db >= lo, x_db <= hi)
    knee_val = x_db + ((1 / ratio) - 1) * ((x_db - threshold_db + knee_db / 2) ** 2) / (2 * knee_db)
    x_sc = torch.where(in_knee, knee_val, x_sc)
    above = x_db > hi
    above_val = threshold_db + (x_db - threshold_db) / ratio
    x_sc = torch.where(above, above_val, x_sc)
    return x_sc - x_db


def compressor(
    x,
    sample_rate: float,
    threshold_db,
    ratio,
    attack_ms,
    release_ms,
    knee_db,
    makeup_gain_db,
    eps: float = 1e-8,
    lookahead_samples: int = 0,
):
    """Feed-forward compressor with one-pole (attack-only) gain smoothing by FSM.
    x (bs, chs, T); parameters (bs,).  ``release_ms`` is accepted and ignored, as in the
    pinned upstream release."""
    bs, chs, seq_len = x.shape
    x_side = x.sum(dim=1, keepdim=True)
    threshold_db = threshold_db.reshape(bs, 1, 1).type_as(x)
    ratio = ratio.reshape(bs, 1, 1).type_as(x)
    attack_ms = attack_ms.reshape(bs, 1, 1).type_as(x)
    knee_db = knee_db.reshape(bs, 1, 1).type_as(x)
    makeup_gain_db = makeup_gain_db.reshape(bs, 1, 1).type_as(x)

    normalized_attack_time = sample_rate * (attack_ms / 1e3)
    alpha_A = torch.exp(-math.log(9.0) / normalized_attack_time)

    x_db = 20 * torch.log10(torch.abs(x_side).clamp(eps))
    g_c = compressor_gain_computer(x_db, threshold_db, ratio, knee_db)

    b = torch.cat([1 - alpha_A, torch.zeros_like(alpha_A)], dim=-1).squeeze(1)  # (bs, 2)
    a = torch.cat([torch.ones_like(alpha_A), -alpha_A], dim=-1).squeeze(1)
    g_s = _signal.lfilter_via_fsm(g_c, b, a)

    if lookahead_samples > 0:
        x = torch.roll(x, lookahead_samples, dims=-1)
        x = torch.cat(
            [torch.zeros_like(x[..., :lookahead_samples]), x[..., lookahead_samples:]], dim=-1
        )

    g_lin = 10 ** ((g_s + makeup_gain_db) / 20.0)
    return x * g_lin


def noise_shaped_reverberation(*args, **kwargs):
    # fx bus: out of scope for this hot path (disabled in every
    # shipped config: configs/models/naive.yaml:7).
    raise NotImplementedError("fx bus reverberation is out of the oracle's scope")
