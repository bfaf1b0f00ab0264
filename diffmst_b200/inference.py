"""Whole-song inference on the device (mst/utils.py:32-173 ``run_diffmst``).

The reference predicts the parameters once on a 262144-sample excerpt and then renders the song window by window:
windows of 262144 samples every 131072, the console on each window (the last one shorter), a Hann weight per window
(the first half of the first window forced to 1) and an overlap-add into a host tensor (mst/utils.py:121-166).
``sliding_window_mix`` is that loop with everything on the device: each window is a strided view of the tracks that
the console kernels consume in place (no copy), the weighting + overlap-add is one kernel per window
(``dmst_ola_hann_add``), nothing crosses to the host.

``run_diffmst`` keeps the reference's signature and return tuple.  Its loudness normalisation (pyloudnorm, host
side, -48 LUFS per track; mst/utils.py:85-101) is host I/O outside the accelerated path: it is applied when
``pyloudnorm`` is importable, exactly as upstream; otherwise the caller passes tracks that are already normalised
(``loudness_normalize=False``)."""
import ctypes
from typing import Optional

import torch

from . import _lib
from .console import _ptr, _require_cuda

ANALYSIS_LEN = 262144   # mst/utils.py:64


def sliding_window_mix(tracks: torch.Tensor, track_params: torch.Tensor, fx_bus_params: torch.Tensor,
                       master_bus_params: torch.Tensor, mix_console: torch.nn.Module, window: int = ANALYSIS_LEN, **use_flags):
    """mst/utils.py:121-166 on the device.  tracks (bs, num_tracks, seq_len) -> (pred_mix (bs, 2, seq_len), and the
    three parameter dictionaries of the last window, as upstream returns them)."""
    lib = _lib.lib()
    _require_cuda(tracks, "tracks")
    if tracks.dim() != 3:
        raise ValueError("tracks must be (bs, num_tracks, seq_len)")
    flags = dict(use_track_input_fader=True, use_track_panner=True, use_track_eq=True, use_track_compressor=True,
                 use_fx_bus=False, use_master_bus=True, use_output_fader=True)   # mst/utils.py:56-62
    flags.update(use_flags)
    bs, _, total = tracks.shape
    dev = tracks.device
    out = torch.zeros(bs, 2, total, dtype=torch.float32, device=dev)
    keep = getattr(mix_console, "materialize_tracks", None)
    if keep is not None:
        mix_console.materialize_tracks = False   # the per-track tensor of a window is never looked at
    dicts = (None, None, None)
    try:
        with torch.no_grad(), torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for i in range(0, total, window // 2):
                view = tracks[..., i:i + window]          # strided view, ragged at the end: consumed in place
                _, mix_w, *dicts = mix_console(view, track_params, fx_bus_params, master_bus_params, **flags)
                n = mix_w.shape[-1]
                # upstream zero-pads a short window to `window` before weighting: only its first n samples land in the
                # output, so the weight is the full-length Hann evaluated on those n samples
                _lib.check(lib.dmst_ola_hann_add(_ptr(mix_w), mix_w.stride(1), ctypes.c_void_p(out.data_ptr() + 4 * i),
                                                 out.stride(1), bs * 2, min(n, total - i), window, 1 if i == 0 else 0, stream),
                           "dmst_ola_hann_add")
    finally:
        if keep is not None:
            mix_console.materialize_tracks = keep
    return (out, *dicts)


def run_diffmst(tracks: torch.Tensor, ref: torch.Tensor, model: torch.nn.Module, mix_console: torch.nn.Module,
                track_start_idx: int = 0, ref_start_idx: int = 0, loudness_normalize: Optional[bool] = None):
    """mst/utils.py:32-173.  tracks (bs, num_tracks, seq_len), ref (bs, 2, seq_len), both on the device; returns
    (pred_mix, pred_track_param_dict, pred_fx_bus_param_dict, pred_master_bus_param_dict)."""
    _require_cuda(tracks, "tracks")
    _require_cuda(ref, "ref")
    analysis_tracks = tracks[..., track_start_idx:track_start_idx + ANALYSIS_LEN] if tracks.shape[-1] >= ANALYSIS_LEN else tracks
    analysis_ref = ref[..., ref_start_idx:ref_start_idx + ANALYSIS_LEN] if ref.shape[-1] >= ANALYSIS_LEN else ref
    if loudness_normalize is None:
        try:
            import pyloudnorm  # noqa: F401
            loudness_normalize = True
        except ImportError:
            loudness_normalize = False
    if loudness_normalize:
        import pyloudnorm as pyln
        meter = pyln.Meter(44100)
        gains, kept = [], []
        host = analysis_tracks.detach().cpu()
        for n in range(host.shape[1]):
            lufs = meter.integrated_loudness(host[:, n:n + 1].squeeze(0).permute(1, 0).numpy())
            if lufs < -80.0:
                continue                               # mst/utils.py:92-94: silent tracks are dropped
            kept.append(n)
            gains.append(10 ** ((-48.0 - lufs) / 20))
        g = torch.tensor(gains, dtype=torch.float32, device=tracks.device).view(1, -1, 1)
        tracks = tracks[:, kept] * g
        analysis_tracks = analysis_tracks[:, kept] * g
    with torch.no_grad():
        track_params, fx_params, master_params = model(analysis_tracks.contiguous(), analysis_ref.contiguous())
    return sliding_window_mix(tracks.contiguous(), track_params, fx_params, master_params, mix_console)
