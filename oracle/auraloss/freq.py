"""ORACLE (test infrastructure, not product code).

CPU restatement of ``auraloss.freq`` (third-party ``auraloss==0.4.0``, pinned by the
reference at requirements.txt:7 / setup.py:34; instantiated at
configs/models/naive.yaml:54-68 and mst/system.py:61-69, called at mst/system.py:332).
Not vendored in /root/reference and not installable here; this restates the published
algorithm.  PARITY UNPINNED by upstream tests; pinned by
closed-form known-answer tests (tests/test_oracle_kats.py).
"""
from typing import List, Optional

import torch


class SpectralConvergenceLoss(torch.nn.Module):
    """mean over rows of ||Y|-|X||_F / ||Y||_F (Frobenius over (bins, frames))."""

    def forward(self, x_mag, y_mag):
        num = torch.norm(y_mag - x_mag, p="fro", dim=[-1, -2])
        den = torch.norm(y_mag, p="fro", dim=[-1, -2])
        return (num / den).mean()


class STFTMagnitudeLoss(torch.nn.Module):
    def __init__(self, log: bool = True, log_eps: float = 0.0, log_fac: float = 1.0,
                 distance: str = "L1", reduction: str = "mean"):
        super().__init__()
        self.log, self.log_eps, self.log_fac = log, log_eps, log_fac
        if distance == "L1":
            self.distance = torch.nn.L1Loss(reduction=reduction)
        elif distance == "L2":
            self.distance = torch.nn.MSELoss(reduction=reduction)
        else:
            raise ValueError(f"Invalid distance: '{distance}'.")

    def forward(self, x_mag, y_mag):
        if self.log:
            x_mag = torch.log(self.log_fac * x_mag + self.log_eps)
            y_mag = torch.log(self.log_fac * y_mag + self.log_eps)
        return self.distance(x_mag, y_mag)


class STFTLoss(torch.nn.Module):
    def __init__(self, fft_size: int = 1024, hop_size: int = 256, win_length: int = 1024,
                 window: str = "hann_window", w_sc: float = 1.0, w_log_mag: float = 1.0,
                 w_lin_mag: float = 0.0, w_phs: float = 0.0, sample_rate: Optional[float] = None,
                 scale: Optional[str] = None, n_bins: Optional[int] = None,
                 perceptual_weighting: bool = False, scale_invariance: bool = False,
                 eps: float = 1e-8, output: str = "loss", reduction: str = "mean",
                 mag_distance: str = "L1", device=None, **kwargs):
        super().__init__()
        if scale is not None or perceptual_weighting or scale_invariance or w_phs:
            raise NotImplementedError("oracle covers the options the reference uses")
        self.fft_size, self.hop_size, self.win_length = fft_size, hop_size, win_length
        self.window = getattr(torch, window)(win_length)
        self.w_sc, self.w_log_mag, self.w_lin_mag = w_sc, w_log_mag, w_lin_mag
        self.eps, self.output, self.reduction = eps, output, reduction
        self.spectralconv = SpectralConvergenceLoss()
        self.logstft = STFTMagnitudeLoss(log=True, reduction=reduction, distance=mag_distance)
        self.linstft = STFTMagnitudeLoss(log=False, reduction=reduction, distance=mag_distance)

    def stft(self, x):
        x_stft = torch.stft(x, self.fft_size, self.hop_size, self.win_length,
                            self.window.to(x.device).type_as(x), return_complex=True)
        return torch.sqrt(torch.clamp(x_stft.real ** 2 + x_stft.imag ** 2, min=self.eps))

    def forward(self, input, target):
        x_mag = self.stft(input.reshape(-1, input.size(-1)))
        y_mag = self.stft(target.reshape(-1, target.size(-1)))
        sc = self.spectralconv(x_mag, y_mag) if self.w_sc else 0.0
        lg = self.logstft(x_mag, y_mag) if self.w_log_mag else 0.0
        ln = self.linstft(x_mag, y_mag) if self.w_lin_mag else 0.0
        return self.w_sc * sc + self.w_log_mag * lg + self.w_lin_mag * ln


class MultiResolutionSTFTLoss(torch.nn.Module):
    def __init__(self, fft_sizes: List[int] = [1024, 2048, 512],
                 hop_sizes: List[int] = [120, 240, 50],
                 win_lengths: List[int] = [600, 1200, 240],
                 window: str = "hann_window", w_sc: float = 1.0, w_log_mag: float = 1.0,
                 w_lin_mag: float = 0.0, w_phs: float = 0.0, sample_rate: Optional[float] = None,
                 scale: Optional[str] = None, n_bins: Optional[int] = None,
                 perceptual_weighting: bool = False, scale_invariance: bool = False, **kwargs):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = torch.nn.ModuleList(
            STFTLoss(fs, ss, wl, window, w_sc, w_log_mag, w_lin_mag, w_phs, sample_rate, scale,
                     n_bins, perceptual_weighting, scale_invariance, **kwargs)
            for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths)
        )

    def forward(self, x, y):
        total = 0.0
        for f in self.stft_losses:
            total = total + f(x, y)
        return total / len(self.stft_losses)
