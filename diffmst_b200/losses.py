"""Drop-in losses backed by the sm_100a kernels.

``MultiResolutionSTFTLoss`` (alias ``MRSTFTLoss``) takes the constructor arguments of
``auraloss.freq.MultiResolutionSTFTLoss`` (third-party, pinned 0.4.0 by the reference; the
reference instantiates it at configs/models/naive.yaml:54-68 and mst/system.py:61-69 and
calls ``loss(pred_mix, ref_mix)`` at mst/system.py:332).  ``AudioFeatureLoss`` mirrors
mst/loss.py:198-260: same constructor, same dictionary keys and ordering, each value
``weight * mse`` so that mst/system.py:334-338 can sum ``val.mean()``.
``batch_stereo_peak_normalize`` mirrors mst/utils.py:14-29.

Framing, spectra (cuFFT), reductions and gradients run in libdiffmst_b200.so; torch carries
memory, streams and the autograd graph.  No CPU fallback.
"""
import ctypes
from typing import List, Optional

import torch

from . import _lib
from .bark import barkscale_fbanks
from .console import _ptr, _require_cuda


def _rows_view(t: torch.Tensor):
    """(B, C, T) -> (tensor, rows, T, row_stride) without copying when rows are equidistant."""
    if t.dim() != 3:
        raise ValueError("expected a (batch, channels, samples) tensor")
    B, C, T = t.shape
    if t.stride(2) != 1 or (B > 1 and t.stride(0) != C * t.stride(1)) or t.stride(1) < T:
        t = t.contiguous()
    return t, B * C, T, t.stride(1)


class _MrstftFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, windows, cfg, grad_enabled=True):
        lib = _lib.lib()
        _require_cuda(x, "input")
        _require_cuda(y, "target")
        if x.shape != y.shape:
            raise ValueError(f"input {tuple(x.shape)} and target {tuple(y.shape)} differ in shape")
        xv, rows, T, xs = _rows_view(x)
        yv, _, _, ys = _rows_view(y)
        if grad_enabled and ctx.needs_input_grad[1]:
            # auraloss differentiates both arguments; the reference only ever passes a detached target
            # (mst/system.py:232-258: the reference mix comes out of torch.no_grad()).  Refuse rather than return a
            # silent zero gradient.
            raise NotImplementedError("MultiResolutionSTFTLoss: a target that requires grad is not supported; "
                                      "detach it (the reference's target is a no-grad reference mix)")
        need_grad = bool(grad_enabled) and ctx.needs_input_grad[0]   # (no_grad: the gradient pass is skipped)
        dev = x.device
        with torch.cuda.device(dev):
            nbytes = lib.dmst_mrstft_workspace_bytes(ctypes.byref(cfg), rows, T)
            if nbytes == 0:
                raise ValueError("MultiResolutionSTFTLoss: invalid configuration for this input length "
                                 "(each fft_size must be even, >= win_length and < 2 * samples)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            terms = torch.empty(3 * cfg.n_res, dtype=torch.float32, device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if need_grad:
                # the per-frame gradients stay in the workspace; backward overlap-adds them scaled by the upstream
                # gradient of the loss (no separate multiply over the (rows, T) gradient)
                rc = lib.dmst_mrstft_forward_keep(_ptr(xv), xs, _ptr(yv), ys, _ptr(windows), ctypes.byref(cfg), rows, T,
                                                  _ptr(loss), _ptr(terms), _ptr(ws), nbytes, stream)
            else:
                out = torch.empty(1 + 3 * cfg.n_res, dtype=torch.float32, device=dev)
                rc = lib.dmst_mrstft_forward(_ptr(xv), xs, _ptr(yv), ys, _ptr(windows), ctypes.byref(cfg), rows, T,
                                             _ptr(out), None, _ptr(ws), nbytes, stream)
                loss, terms = out[0], out[1:]
        _lib.check(rc, "dmst_mrstft_forward")
        if need_grad:
            ctx.ws, ctx.nbytes, ctx.cfg, ctx.windows, ctx.rows, ctx.T = ws, nbytes, cfg, windows, rows, T
        else:
            ctx.ws = None
        ctx.shape = x.shape
        ctx.mark_non_differentiable(terms)
        return loss, terms

    @staticmethod
    def backward(ctx, gloss, _gterms):
        if ctx.ws is None:
            return None, None, None, None, None
        lib = _lib.lib()
        dev = ctx.ws.device
        g = gloss.detach().to(dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            gx = torch.empty(ctx.rows, ctx.T, dtype=torch.float32, device=dev)
            rc = lib.dmst_mrstft_backward(_ptr(ctx.windows), ctypes.byref(ctx.cfg), ctx.rows, ctx.T, _ptr(g), _ptr(gx),
                                          _ptr(ctx.ws), ctx.nbytes,
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "dmst_mrstft_backward")
        return gx.view(ctx.shape), None, None, None, None


class MultiResolutionSTFTLoss(torch.nn.Module):
    def __init__(
        self,
        fft_sizes: List[int] = [1024, 2048, 512],
        hop_sizes: List[int] = [120, 240, 50],
        win_lengths: List[int] = [600, 1200, 240],
        window: str = "hann_window",
        w_sc: float = 1.0,
        w_log_mag: float = 1.0,
        w_lin_mag: float = 0.0,
        w_phs: float = 0.0,
        sample_rate: Optional[float] = None,
        scale: Optional[str] = None,
        n_bins: Optional[int] = None,
        perceptual_weighting: bool = False,
        scale_invariance: bool = False,
        **kwargs,
    ):
        super().__init__()
        if not (len(fft_sizes) == len(hop_sizes) == len(win_lengths)):
            raise AssertionError("fft_sizes, hop_sizes and win_lengths must have equal length")
        if len(fft_sizes) > _lib.MRSTFT_MAX_RES:
            raise ValueError(f"at most {_lib.MRSTFT_MAX_RES} resolutions")
        unsupported = {"w_phs": w_phs, "scale": scale, "perceptual_weighting": perceptual_weighting,
                       "scale_invariance": scale_invariance}
        for k, v in unsupported.items():
            if v:
                raise NotImplementedError(
                    f"MultiResolutionSTFTLoss({k}={v!r}) is not used by any reference config "
                    "(configs/models/naive.yaml:54-68, mst/system.py:61-69) and is not implemented")
        extra = {k: v for k, v in kwargs.items() if k not in ("eps", "output", "reduction", "mag_distance", "device")}
        if extra:
            raise TypeError(f"unexpected arguments {sorted(extra)}")
        if kwargs.get("output", "loss") != "loss" or kwargs.get("reduction", "mean") != "mean" \
                or kwargs.get("mag_distance", "L1") != "L1":
            raise NotImplementedError("only output='loss', reduction='mean', mag_distance='L1'")
        self.fft_sizes, self.hop_sizes, self.win_lengths = list(fft_sizes), list(hop_sizes), list(win_lengths)
        self.window = window
        self.w_sc, self.w_log_mag, self.w_lin_mag, self.w_phs = w_sc, w_log_mag, w_lin_mag, w_phs
        self.eps = float(kwargs.get("eps", 1e-8))
        self.sample_rate = sample_rate
        self._windows = {}
        self.last_terms = None  # (n_res, 3) tensor of (L_sc, L_log, L_lin) from the last call

    def _cfg(self):
        c = _lib.MrstftCfg()
        c.n_res = len(self.fft_sizes)
        for i, (n, h, w) in enumerate(zip(self.fft_sizes, self.hop_sizes, self.win_lengths)):
            c.fft_size[i], c.hop_size[i], c.win_length[i] = int(n), int(h), int(w)
        c.w_sc, c.w_log_mag, c.w_lin_mag, c.eps = self.w_sc, self.w_log_mag, self.w_lin_mag, self.eps
        return c

    def _windows_on(self, device):
        if device not in self._windows:
            self._windows[device] = torch.cat(
                [getattr(torch, self.window)(int(w)).float() for w in self.win_lengths]).to(device)
        return self._windows[device]

    def forward(self, x: torch.Tensor, y: torch.Tensor):
        loss, terms = _MrstftFunction.apply(x, y, self._windows_on(x.device), self._cfg(), torch.is_grad_enabled())
        self.last_terms = terms.view(-1, 3)
        return loss


MRSTFTLoss = MultiResolutionSTFTLoss


class _AflFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target, fb, window, weights, fft_size, n_bands):
        lib = _lib.lib()
        _require_cuda(inp, "input")
        _require_cuda(target, "target")
        if inp.shape != target.shape or inp.dim() != 3 or inp.shape[1] != 2:
            raise AssertionError("Input must be stereo")  # mst/loss.py:158
        if inp.stride(2) != 1 or inp.stride(1) < inp.shape[2]:
            inp = inp.contiguous()
        if target.stride() != inp.stride():
            target = target.contiguous()
            inp = inp.contiguous()
        B, _, T = inp.shape
        dev = inp.device
        w = (ctypes.c_float * 5)(*[float(v) for v in weights])
        with torch.cuda.device(dev):
            nbytes = lib.dmst_afl_workspace_bytes(B, T, fft_size, n_bands)
            if nbytes == 0:
                raise ValueError(f"AudioFeatureLoss needs more than {fft_size // 2} samples (reflect padding of "
                                 "the bark-spectrum STFT, mst/loss.py:106-112)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            losses = torch.empty(5, dtype=torch.float32, device=dev)
            rc = lib.dmst_afl_forward(_ptr(inp), _ptr(target), inp.stride(0), inp.stride(1), _ptr(fb), _ptr(window),
                                      w, B, T, fft_size, n_bands, _ptr(losses), _ptr(ws), nbytes,
                                      ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "dmst_afl_forward")
        ctx.save_for_backward(inp, fb, window)
        ctx.ws, ctx.nbytes, ctx.w, ctx.cfg = ws, nbytes, w, (fft_size, n_bands)
        return losses

    @staticmethod
    def backward(ctx, glosses):
        lib = _lib.lib()
        inp, fb, window = ctx.saved_tensors
        fft_size, n_bands = ctx.cfg
        B, _, T = inp.shape
        dev = inp.device
        gw = glosses.contiguous().float()
        with torch.cuda.device(dev):
            gx = torch.empty(B, 2, T, dtype=torch.float32, device=dev)
            rc = lib.dmst_afl_backward(_ptr(inp), inp.stride(0), inp.stride(1), _ptr(fb), _ptr(window), ctx.w,
                                       _ptr(gw), B, T, fft_size, n_bands, _ptr(gx), _ptr(ctx.ws), ctx.nbytes,
                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "dmst_afl_backward")
        return gx, None, None, None, None, None, None


class AudioFeatureLoss(torch.nn.Module):
    """mst/loss.py:198-260.  Keys: mix-rms, mix-crest_factor, mix-stereo_width,
    mix-stereo_imbalance, mix-barkspectrum."""

    FEATURES = ["rms", "crest_factor", "stereo_width", "stereo_imbalance", "barkspectrum"]
    FFT_SIZE, N_BANDS, F_MIN, F_MAX = 32768, 24, 20.0, 20000.0  # compute_barkspectrum defaults

    def __init__(self, weights: List[float], sample_rate: int, stem_separation: bool = False,
                 use_clap: bool = False) -> None:
        super().__init__()
        self.weights = weights
        self.sample_rate = sample_rate
        self.stem_separation = stem_separation
        self.sources_list = ["mix"]
        self.source_weights = [1.0]
        self.use_clap = use_clap
        assert len(self.FEATURES) == len(weights)
        self._cache = {}

    def _tables(self, device):
        if device not in self._cache:
            fb = barkscale_fbanks(self.FFT_SIZE // 2 + 1, self.F_MIN, self.F_MAX, self.N_BANDS, self.sample_rate)
            self._cache[device] = (fb.t().contiguous().to(device), torch.hann_window(self.FFT_SIZE).to(device))
        return self._cache[device]

    def forward(self, input: torch.Tensor, target: torch.Tensor):
        fb, window = self._tables(input.device)
        vals = _AflFunction.apply(input, target, fb, window, self.weights, self.FFT_SIZE, self.N_BANDS)
        return {f"{self.sources_list[0]}-{name}": vals[i] * self.source_weights[0]
                for i, name in enumerate(self.FEATURES)}


def batch_stereo_peak_normalize(x: torch.Tensor) -> torch.Tensor:
    """mst/utils.py:14-29; used on the no-grad reference mix (mst/system.py:249).  Returns a new
    tensor; no gradient is propagated."""
    lib = _lib.lib()
    _require_cuda(x, "x")
    if x.dim() != 3 or x.shape[1] != 2:
        raise ValueError("expected (bs, 2, seq_len)")
    x = x.detach()
    if x.stride(2) != 1:
        x = x.contiguous()
    B, _, T = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty(B, 2, T, dtype=torch.float32, device=x.device)
        rc = lib.dmst_peak_normalize(_ptr(x), x.stride(0), x.stride(1), _ptr(y), B, T,
                                     ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc, "dmst_peak_normalize")
    return y
