"""ORACLE (test infrastructure, not product code).

CPU restatement of the reference's in-tree loss stack, dtype-generic:

* bark filterbank (Traunmueller)            mst/filter.py:38-161
* feature transforms                         mst/loss.py:62-195
* AudioFeatureLoss                           mst/loss.py:198-260
* batch_stereo_peak_normalize                mst/utils.py:14-29

Validated against the unmodified reference modules imported from /root/reference by
tests/golden/make_golden.py, and pinned by the golden vectors that script writes.
"""
import torch


def _hz_to_bark(f: float) -> float:
    # mst/filter.py:57-65 (traunmuller branch)
    b = (26.81 * f) / (1960.0 + f) - 0.53
    if b < 2:
        b += 0.15 * (2 - b)
    elif b > 20.1:
        b += 0.22 * (b - 20.1)
    return b


def _bark_to_hz(barks: torch.Tensor) -> torch.Tensor:
    # mst/filter.py:90-100.  NB the reference uses if/elif on *any()*: when some point is
    # below 2 bark (always true for f_min = 20 Hz) the >20.1 correction is never applied.
    barks = barks.clone()
    if bool((barks < 2).any()):
        idx = barks < 2
        barks[idx] = (barks[idx] - 0.3) / 0.85
    elif bool((barks > 20.1).any()):
        idx = barks > 20.1
        barks[idx] = (barks[idx] + 4.422) / 1.22
    return 1960 * ((barks + 0.53) / (26.28 - barks))


def barkscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_barks: int, sample_rate: int):
    """(n_freqs, n_barks) triangular filters, float32 like the reference (filter.py:107-161)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_bark(f_min), _hz_to_bark(f_max), n_barks + 2)
    f_pts = _bark_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def compute_barkspectrum(x, fft_size=32768, n_bands=24, sample_rate=44100, f_min=20.0,
                         f_max=20000.0, **kwargs):
    """mid/side bark spectrum, (bs, 24, 2); mst/loss.py:62-124 (mode="mid-side")."""
    fb = barkscale_fbanks(fft_size // 2 + 1, f_min, f_max, n_bands, sample_rate)
    fb = fb.unsqueeze(0).type_as(x).permute(0, 2, 1)
    outs = []
    for sig in (x[:, 0, :] + x[:, 1, :], x[:, 0, :] - x[:, 1, :]):
        X = torch.stft(sig, n_fft=fft_size, hop_length=fft_size // 4, return_complex=True,
                       window=torch.hann_window(fft_size).to(x.device).type_as(x))
        X = torch.abs(X).mean(dim=-1, keepdim=True)
        outs.append(torch.log(torch.matmul(fb, X) + 1e-8))
    return torch.cat(outs, dim=-1)


def compute_rms(x, **kwargs):
    return torch.sqrt(torch.mean(x ** 2, dim=-1).clamp(min=1e-8))


def compute_crest_factor(x, **kwargs):
    num = torch.max(torch.abs(x), dim=-1)[0]
    den = compute_rms(x).clamp(min=1e-8)
    return 20 * torch.log10((num / den).clamp(min=1e-8))


def compute_stereo_width(x, **kwargs):
    s = torch.mean((x[:, 0, :] + x[:, 1, :]) ** 2, dim=-1)
    d = torch.mean((x[:, 0, :] - x[:, 1, :]) ** 2, dim=-1)
    return d / s.clamp(min=1e-8)


def compute_stereo_imbalance(x, **kwargs):
    l = torch.mean(x[:, 0, :] ** 2, dim=-1)
    r = torch.mean(x[:, 1, :] ** 2, dim=-1)
    return (r - l) / (r + l).clamp(min=1e-8)


class OracleAudioFeatureLoss(torch.nn.Module):
    """mst/loss.py:198-260; keys ``mix-<transform>``."""

    NAMES = ["rms", "crest_factor", "stereo_width", "stereo_imbalance", "barkspectrum"]
    FNS = [compute_rms, compute_crest_factor, compute_stereo_width, compute_stereo_imbalance,
           compute_barkspectrum]

    def __init__(self, weights, sample_rate, stem_separation=False, use_clap=False):
        super().__init__()
        assert len(weights) == len(self.FNS)
        self.weights, self.sample_rate = weights, sample_rate

    def forward(self, input, target):
        losses = {}
        for name, fn, w in zip(self.NAMES, self.FNS, self.weights):
            a = fn(input, sample_rate=self.sample_rate)
            b = fn(target, sample_rate=self.sample_rate)
            losses[f"mix-{name}"] = w * torch.nn.functional.mse_loss(a, b) * 1.0
        return losses


def batch_stereo_peak_normalize(x):
    """mst/utils.py:14-29."""
    g = x.abs().max(dim=-1, keepdim=True)[0].max(dim=-2, keepdim=True)[0]
    return x / g.clamp(1e-8)
