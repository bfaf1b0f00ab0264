"""Reference-mix generation on the device.

``naive_random_mix`` keeps the signature, keyword spelling (``use_ouput_fader``) and 8-tuple of
mst/mixing.py:35-94, so ``mix_fn: diffmst_b200.mixing.naive_random_mix`` is a config change.  The
difference is where the random parameters live: upstream draws them with the CPU generator and
copies them to the device every call (mst/mixing.py:61-69: three H2D copies and a stream
stall per step, twice per step in ``System.common_step``); here they are drawn on the tracks'
device, optionally from a caller-supplied ``torch.Generator`` for reproducibility.

``random_reference_mix`` is the whole target pipeline of mst/system.py:232-253 in one call:
random mix under no_grad -> peak normalise (mst/utils.py:14-29) -> NaN flag kept on the device
(upstream raises from the host after a blocking ``isnan().any()``; callers that want that
behaviour check the returned flag when they next synchronise).
"""
from typing import Optional

import torch

from .losses import batch_stereo_peak_normalize


def naive_random_mix(
    tracks: torch.Tensor,
    mix_console: torch.nn.Module,
    use_track_input_fader: bool = True,
    use_track_eq: bool = True,
    use_track_compressor: bool = True,
    use_track_panner: bool = True,
    use_fx_bus: bool = True,
    use_master_bus: bool = True,
    use_ouput_fader: bool = True,
    generator: Optional[torch.Generator] = None,
    **kwargs,
):
    """Random mix with parameters sampled uniformly on the console's ranges (mst/mixing.py:35-94)."""
    bs, num_tracks, seq_len = tracks.size()
    opts = dict(device=tracks.device, dtype=tracks.dtype, generator=generator)
    mix_params = torch.rand(bs, num_tracks, mix_console.num_track_control_params, **opts)
    fx_bus_params = torch.rand(bs, mix_console.num_fx_bus_control_params, **opts)
    master_bus_params = torch.rand(bs, mix_console.num_master_bus_control_params, **opts)
    with torch.no_grad():
        mixed_tracks, mix, track_param_dict, fx_bus_param_dict, master_bus_param_dict = mix_console(
            tracks, mix_params, fx_bus_params, master_bus_params,
            use_track_input_fader=use_track_input_fader, use_track_eq=use_track_eq,
            use_track_compressor=use_track_compressor, use_track_panner=use_track_panner,
            use_master_bus=use_master_bus, use_fx_bus=use_fx_bus, use_output_fader=use_ouput_fader)
    return (mixed_tracks, mix, track_param_dict, fx_bus_param_dict, master_bus_param_dict, mix_params,
            fx_bus_params, master_bus_params)


def random_reference_mix(tracks: torch.Tensor, mix_console: torch.nn.Module,
                         generator: Optional[torch.Generator] = None, **use_flags):
    """Target mix of a training step (mst/system.py:232-253): random mix -> peak normalise.

    Returns ``(ref_mix, has_nan, params)``: the normalised mix ``(bs, 2, seq_len)``, a 0-dim bool
    tensor on the device (upstream: ``raise ValueError("Found nan in ref_mix")``), and the three
    normalised parameter tensors that produced it."""
    out = naive_random_mix(tracks, mix_console, generator=generator, **use_flags)
    ref_mix = batch_stereo_peak_normalize(out[1])
    return ref_mix, torch.isnan(ref_mix).any(), out[5:]
