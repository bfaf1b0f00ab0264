"""ORACLE (test infrastructure, not product code).

Exact float64 *time-domain* evaluation of the same console chain, used to pin the FSM
shim (oracle/dasp_pytorch) from an independent direction: the reference applies its IIR
filters by frequency sampling; for stable filters and the lengths
the reference uses this equals the direct recursion to ~1e-12 in float64.  The CUDA
kernels compute the recursion, so this file is also the "what the kernel should produce"
statement at short lengths where FSM time-aliasing is not negligible.

numpy / scipy only.
"""
import math

import numpy as np
from scipy.signal import lfilter


def rbj(gain_db, f, q, sr, kind):
    A = 10 ** (gain_db / 40.0)
    w0 = 2 * math.pi * f / sr
    al = math.sin(w0) / (2 * q)
    c = math.cos(w0)
    sA = math.sqrt(A)
    if kind == "peaking":
        b = [1 + al * A, -2 * c, 1 - al * A]
        a = [1 + al / A, -2 * c, 1 - al / A]
    elif kind == "low_shelf":
        b = [A * ((A + 1) - (A - 1) * c + 2 * sA * al), 2 * A * ((A - 1) - (A + 1) * c),
             A * ((A + 1) - (A - 1) * c - 2 * sA * al)]
        a = [(A + 1) + (A - 1) * c + 2 * sA * al, -2 * ((A - 1) + (A + 1) * c),
             (A + 1) + (A - 1) * c - 2 * sA * al]
    elif kind == "high_shelf":
        b = [A * ((A + 1) + (A - 1) * c + 2 * sA * al), -2 * A * ((A - 1) + (A + 1) * c),
             A * ((A + 1) + (A - 1) * c - 2 * sA * al)]
        a = [(A + 1) - (A - 1) * c + 2 * sA * al, 2 * ((A - 1) - (A + 1) * c),
             (A + 1) - (A - 1) * c - 2 * sA * al]
    else:
        raise ValueError(kind)
    b = np.asarray(b, dtype=np.float64) / a[0]
    a = np.asarray(a, dtype=np.float64) / a[0]
    return b, a


KINDS = ["low_shelf", "peaking", "peaking", "peaking", "peaking", "high_shelf"]


def parametric_eq(x, sr, eq18):
    """x (..., T) float64; eq18 = 18 scalars [gain, freq, q] x 6 sections."""
    y = np.asarray(x, dtype=np.float64)
    for k, kind in enumerate(KINDS):
        b, a = rbj(eq18[3 * k], eq18[3 * k + 1], eq18[3 * k + 2], sr, kind)
        y = lfilter(b, a, y, axis=-1)
    return y


def gain_computer(x_db, thr, ratio, knee):
    lo, hi = thr - knee / 2, thr + knee / 2
    x_sc = np.where((x_db >= lo) & (x_db <= hi),
                    x_db + (1 / ratio - 1) * (x_db - thr + knee / 2) ** 2 / (2 * knee), x_db)
    x_sc = np.where(x_db > hi, thr + (x_db - thr) / ratio, x_sc)
    return x_sc - x_db


def compressor(x, sr, thr, ratio, attack_ms, knee, makeup, lookahead, eps=1e-8):
    """x (chs, T); linked side-chain = channel sum."""
    x = np.asarray(x, dtype=np.float64)
    side = x.sum(axis=0)
    alpha = math.exp(-math.log(9.0) / (sr * attack_ms / 1e3))
    x_db = 20 * np.log10(np.maximum(np.abs(side), eps))
    g_c = gain_computer(x_db, thr, ratio, knee)
    g_s = lfilter([1 - alpha], [1.0, -alpha], g_c)
    xd = np.zeros_like(x)
    if lookahead > 0:
        xd[:, lookahead:] = x[:, :-lookahead] if lookahead < x.shape[-1] else 0
    else:
        xd = x
    return xd * 10 ** ((g_s + makeup) / 20.0)


def pan_gains(pan):
    th = pan * math.pi / 2
    return (math.sqrt((math.pi / 2 - th) * (2 / math.pi) * math.cos(th)),
            math.sqrt(th * (2 / math.pi) * math.sin(th)))


def console(tracks, track_p, master_p, sr, use_input_fader=True, use_eq=True, use_comp=True,
            use_master=True, use_output_fader=True):
    """tracks (N, T); track_p (N, 27) and master_p (26,) DENORMALISED, in the reference's
    index order (mst/modules.py:353-460).  Returns (mixed (2, N, T), mix (2, T))."""
    N, T = tracks.shape
    mixed = np.zeros((2, N, T))
    for n in range(N):
        p = track_p[n]
        y = np.asarray(tracks[n], dtype=np.float64)
        if use_input_fader:
            y = y * 10 ** (p[0] / 20.0)
        if use_eq:
            y = parametric_eq(y, sr, p[1:19])
        if use_comp:
            y = compressor(y[None], sr, p[19], p[20], p[21], p[23], p[24], 2048)[0]
        gl, gr = pan_gains(p[25])
        mixed[0, n], mixed[1, n] = gl * y, gr * y
    mix = mixed.sum(axis=1)
    if use_master:
        mix = mix * 10 ** (master_p[25] / 20.0)
        mix = parametric_eq(mix, sr, master_p[0:18])
        mix = compressor(mix, sr, master_p[18], master_p[19], master_p[20], master_p[22],
                         master_p[23], 1024)
    if use_output_fader:
        mix = mix * 10 ** (master_p[24] / 20.0)
    return mixed, mix
