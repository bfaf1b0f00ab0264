"""ORACLE (test infrastructure, not product code).

CPU restatement of the parts of ``dasp_pytorch.signal`` (third-party, pinned
``dasp-pytorch==0.0.1`` by the reference: requirements.txt:15, setup.py:35) that the
reference mix console reaches through ``dasp_pytorch.functional`` (call sites
mst/modules.py:237,246,293,300).  The package is not vendored in /root/reference and
cannot be installed here (no network), so this file restates its *published algorithm*
(SURVEY.md Appendix A): RBJ-cookbook biquads and IIR filtering by the
frequency-sampling method (FSM).  PARITY UNPINNED by upstream tests (the reference ships
no golden vectors); it is pinned instead by the analytic known-answer tests
in tests/test_oracle_kats.py and by an exact float64 time-domain recursion
(oracle/timedomain.py).

Everything is dtype-generic: run it in float64 for the parity oracle and in float32 for
the "reference-style" CPU baseline.
"""
import math

import torch


def _next_pow2_fsm(seq_len: int) -> int:
    # n_fft = 2 ** ceil(log2(2T - 1)) (Appendix A, sosfilt_via_fsm / lfilter_via_fsm)
    return 1 << max(0, math.ceil(math.log2(max(1, 2 * seq_len - 1))))


def biquad(gain_db, cutoff_freq, q_factor, sample_rate: float, filter_type: str):
    """RBJ cookbook biquad.  Inputs broadcastable tensors; returns (b, a) each (..., 3),
    normalised by a0 (Appendix A)."""
    A = 10 ** (gain_db / 40.0)
    w0 = 2 * math.pi * (cutoff_freq / sample_rate)
    alpha = torch.sin(w0) / (2 * q_factor)
    cos_w0 = torch.cos(w0)
    sqrt_A = torch.sqrt(A)

    if filter_type == "high_shelf":
        b0 = A * ((A + 1) + (A - 1) * cos_w0 + 2 * sqrt_A * alpha)
        b1 = -2 * A * ((A - 1) + (A + 1) * cos_w0)
        b2 = A * ((A + 1) + (A - 1) * cos_w0 - 2 * sqrt_A * alpha)
        a0 = (A + 1) - (A - 1) * cos_w0 + 2 * sqrt_A * alpha
        a1 = 2 * ((A - 1) - (A + 1) * cos_w0)
        a2 = (A + 1) - (A - 1) * cos_w0 - 2 * sqrt_A * alpha
    elif filter_type == "low_shelf":
        b0 = A * ((A + 1) - (A - 1) * cos_w0 + 2 * sqrt_A * alpha)
        b1 = 2 * A * ((A - 1) - (A + 1) * cos_w0)
        b2 = A * ((A + 1) - (A - 1) * cos_w0 - 2 * sqrt_A * alpha)
        a0 = (A + 1) + (A - 1) * cos_w0 + 2 * sqrt_A * alpha
        a1 = -2 * ((A - 1) + (A + 1) * cos_w0)
        a2 = (A + 1) + (A - 1) * cos_w0 - 2 * sqrt_A * alpha
    elif filter_type == "peaking":
        b0 = 1 + alpha * A
        b1 = -2 * cos_w0
        b2 = 1 - alpha * A
        a0 = 1 + (alpha / A)
        a1 = -2 * cos_w0
        a2 = 1 - (alpha / A)
    else:
        raise ValueError(f"Invalid filter_type: {filter_type}.")

    b = torch.stack([b0, b1, b2], dim=-1)
    a = torch.stack([a0, a1, a2], dim=-1)
    b = b / a0.unsqueeze(-1)
    a = a / a0.unsqueeze(-1)
    return b, a


def fft_freqz(b, a, n_fft: int):
    """H = rfft(b, n) / rfft(a, n) along the last dim."""
    return torch.fft.rfft(b, n_fft, dim=-1) / torch.fft.rfft(a, n_fft, dim=-1)


def fft_sosfreqz(sos, n_fft: int):
    """sos (bs, n_sections, 6) = [b0 b1 b2 a0 a1 a2] -> product of section responses."""
    n_sections = sos.shape[1]
    H = None
    for k in range(n_sections):
        Hk = fft_freqz(sos[:, k, :3], sos[:, k, 3:], n_fft)
        H = Hk if H is None else H * Hk
    return H


def freqdomain_fir(x, H, n_fft: int):
    X = torch.fft.rfft(x, n_fft, dim=-1)
    return torch.fft.irfft(X * H, n_fft, dim=-1)


def lfilter_via_fsm(x, b, a=None):
    """x (bs, 1, T), b/a (bs, M): IIR by frequency sampling, cropped to T."""
    bs, chs, seq_len = x.shape
    assert chs == 1
    n_fft = _next_pow2_fsm(seq_len)
    b = b.type_as(x)
    if a is None:
        H = torch.fft.rfft(b, n_fft, dim=-1)
    else:
        H = fft_freqz(b, a.type_as(x), n_fft)
    y = freqdomain_fir(x, H.unsqueeze(1), n_fft)
    return y[..., :seq_len]


def sosfilt_via_fsm(sos, x):
    """sos (bs, n_sections, 6), x (bs, chs, T): same H for every channel of a row."""
    bs, chs, seq_len = x.shape
    n_fft = _next_pow2_fsm(seq_len)
    H = fft_sosfreqz(sos.type_as(x), n_fft)
    y = freqdomain_fir(x, H.unsqueeze(1), n_fft)
    return y[..., :seq_len]
