"""TEST INFRASTRUCTURE: drives the host-emulated build of the kernels
(build/libdiffmst_emul.so, see cuda_emul.h) with numpy buffers.  Used only to debug kernel
logic in the GPU-less build container; never imported by the product."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

USE_TRACK_INPUT_FADER, USE_TRACK_EQ, USE_TRACK_COMPRESSOR, USE_TRACK_PANNER = 1, 2, 4, 8
USE_MASTER_BUS, USE_FX_BUS, USE_OUTPUT_FADER = 16, 32, 64
WANT_MIXED_TRACKS, WANT_GRAD_TRACKS, BASIC_CONSOLE = 128, 256, 512


class Ranges(ctypes.Structure):
    _fields_ = [("track_lo", ctypes.c_float * 27), ("track_hi", ctypes.c_float * 27),
                ("master_lo", ctypes.c_float * 26), ("master_hi", ctypes.c_float * 26)]


def default_ranges(sr=44100):
    import sys
    sys.path.insert(0, ROOT)
    from oracle.console import param_ranges, EQ_KEYS, COMP_KEYS
    pr = param_ranges(sr)
    r = Ranges()
    t = [pr["input_fader"]["gain_db"]] + [pr["parametric_eq"][k] for k in EQ_KEYS] + \
        [pr["compressor"][k] for k in COMP_KEYS] + [pr["stereo_panner"]["pan"], pr["fx_bus"]["send_db"]]
    m = [pr["parametric_eq"][k] for k in EQ_KEYS] + [pr["compressor"][k] for k in COMP_KEYS] + \
        [pr["output_fader"]["gain_db"], pr["input_fader"]["gain_db"]]
    for i, (lo, hi) in enumerate(t):
        r.track_lo[i], r.track_hi[i] = lo, hi
    for i, (lo, hi) in enumerate(m):
        r.master_lo[i], r.master_hi[i] = lo, hi
    return r


def load():
    lib = ctypes.CDLL(os.path.join(ROOT, "build", "libdiffmst_emul.so"))
    lib.dmst_console_workspace_bytes.restype = ctypes.c_size_t
    lib.dmst_console_workspace_bytes.argtypes = [ctypes.c_int] * 3 + [ctypes.c_uint]
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def flags_from(use_track_input_fader=True, use_track_eq=True, use_track_compressor=True,
               use_track_panner=True, use_master_bus=True, use_fx_bus=False, use_output_fader=True):
    f = 0
    f |= USE_TRACK_INPUT_FADER if use_track_input_fader else 0
    f |= USE_TRACK_EQ if use_track_eq else 0
    f |= USE_TRACK_COMPRESSOR if use_track_compressor else 0
    f |= USE_TRACK_PANNER if use_track_panner else 0
    f |= USE_MASTER_BUS if use_master_bus else 0
    f |= USE_FX_BUS if use_fx_bus else 0
    f |= USE_OUTPUT_FADER if use_output_fader else 0
    return f


class EmulConsole:
    def __init__(self, sr=44100):
        self.lib = load()
        self.sr = sr
        self.ranges = default_ranges(sr)

    def forward(self, tracks, tp, mp, flags, la_t=2048, la_m=1024, want_mixed=True, want_grad_tracks=True):
        B, N, T = tracks.shape
        tracks = np.ascontiguousarray(tracks, dtype=np.float32)
        tp = np.ascontiguousarray(tp, dtype=np.float32)
        mp = np.ascontiguousarray(mp, dtype=np.float32) if mp is not None else None
        if want_mixed:
            flags |= WANT_MIXED_TRACKS
        if want_grad_tracks:   # shapes the checkpoints forward keeps: must be known at forward time
            flags |= WANT_GRAD_TRACKS
        nbytes = self.lib.dmst_console_workspace_bytes(B, N, T, flags)
        ws = np.zeros(nbytes // 4 + 64, dtype=np.float32)
        off = (-ws.ctypes.data) % 256
        wsp = ctypes.c_void_p(ws.ctypes.data + off)
        mix = np.full((B, 2, T), np.nan, dtype=np.float32)
        mixed = np.full((B, 2, N, T), np.nan, dtype=np.float32) if want_mixed else None
        status = np.zeros(4, dtype=np.int32)
        rc = self.lib.dmst_console_forward(
            _p(tracks), ctypes.c_longlong(N * T), ctypes.c_longlong(T), _p(tp), _p(mp),
            ctypes.byref(self.ranges), ctypes.c_float(self.sr), B, N, T, ctypes.c_uint(flags), la_t, la_m,
            _p(mix), _p(mixed), _p(status), wsp, ctypes.c_size_t(nbytes), None)
        assert rc == 0, rc
        self._saved = (tracks, tp, mp, flags, la_t, la_m, ws, wsp, nbytes)
        return mix, mixed, status

    def backward(self, gmix, gmixed=None):
        tracks, tp, mp, flags, la_t, la_m, ws, wsp, nbytes = self._saved
        B, N, T = tracks.shape
        want_grad_tracks = bool(flags & WANT_GRAD_TRACKS)
        gmix = np.ascontiguousarray(gmix, dtype=np.float32)
        gtp = np.full(tp.shape, np.nan, dtype=np.float32)
        gmp = np.full((B, 26), np.nan, dtype=np.float32) if mp is not None else None
        gtr = np.full((B, N, T), np.nan, dtype=np.float32) if want_grad_tracks else None
        rc = self.lib.dmst_console_backward(
            _p(tracks), ctypes.c_longlong(N * T), ctypes.c_longlong(T), _p(tp), _p(mp),
            ctypes.byref(self.ranges), ctypes.c_float(self.sr), B, N, T, ctypes.c_uint(flags), la_t, la_m,
            _p(gmix), _p(gmixed), _p(gtp), _p(gmp), _p(gtr), wsp, ctypes.c_size_t(nbytes), None)
        assert rc == 0, rc
        return gtp, gmp, gtr
