"""Bark-scale triangular filterbank, host-side, built once and cached (the reference rebuilds
it on the CPU inside every loss call, mst/loss.py:88-90).  Restates the construction of
mst/filter.py:38-161 (Traunmueller scale) including its quirk that the >20.1-bark inverse
correction is skipped whenever some point lies below 2 bark (filter.py:90-96 uses if/elif on
any())."""
import torch


def _hz_to_bark(f: float) -> float:
    b = (26.81 * f) / (1960.0 + f) - 0.53
    if b < 2:
        b += 0.15 * (2 - b)
    elif b > 20.1:
        b += 0.22 * (b - 20.1)
    return b


def _bark_to_hz(b: torch.Tensor) -> torch.Tensor:
    b = b.clone()
    low, high = b < 2, b > 20.1
    if bool(low.any()):
        b[low] = (b[low] - 0.3) / 0.85
    elif bool(high.any()):
        b[high] = (b[high] + 4.422) / 1.22
    return 1960 * ((b + 0.53) / (26.28 - b))


def barkscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_barks: int, sample_rate: int) -> torch.Tensor:
    """(n_freqs, n_barks) float32, as mst.filter.barkscale_fbanks returns it."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    edges = _bark_to_hz(torch.linspace(_hz_to_bark(f_min), _hz_to_bark(f_max), n_barks + 2))
    width = edges[1:] - edges[:-1]
    slopes = edges.unsqueeze(0) - freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / width[:-1]
    up = slopes[:, 2:] / width[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))
