"""ORACLE (test infrastructure, not product code).

CPU restatement of the reference mix console, dtype-generic (run in float64 for parity,
float32 for the CPU baseline).  Follows:

* parameter index map and ranges      mst/modules.py:121-181, 353-460
* denormalisation + range ValueError   mst/modules.py:71-97
* processing chain and flags           mst/modules.py:186-314 (gain -> EQ -> compressor
  (lookahead 2048) -> pan -> bus sum -> master gain -> EQ -> compressor (lookahead 1024)
  -> output fader)
* BasicMixConsole: absent at the reference commit; reconstructed from README.md:14,
  mst/mixing.py:122-164, 935-945 (two per-track parameters: gain_db, pan).

The DSP arithmetic comes from the dasp_pytorch shim next to this file (SURVEY.md
Appendix A).  This module is validated against the *unmodified* reference class imported
from /root/reference (tests/golden/make_golden.py) and pinned by the golden vectors that
script writes.
"""
import torch

from .dasp_pytorch import functional as F

EQ_BANDS = ["low_shelf", "band0", "band1", "band2", "band3", "high_shelf"]
EQ_KEYS = [f"{b}_{p}" for b in EQ_BANDS for p in ("gain_db", "cutoff_freq", "q_factor")]
COMP_KEYS = ["threshold_db", "ratio", "attack_ms", "release_ms", "knee_db", "makeup_gain_db"]


def param_ranges(sample_rate, input_min_gain_db=-48.0, input_max_gain_db=48.0,
                 output_min_gain_db=-48.0, output_max_gain_db=48.0, min_send_db=-80.0,
                 max_send_db=12.0, eq_min_gain_db=-12.0, eq_max_gain_db=12.0, min_pan=0.0,
                 max_pan=1.0, reverb_min_band_gain=0.0, reverb_max_band_gain=1.0,
                 reverb_min_band_decay=0.0, reverb_max_band_decay=1.0):
    """Range table of mst/modules.py:121-181."""
    hi = (sample_rate // 2) - 1000
    cut = {"low_shelf": (20, 2000), "band0": (80, 2000), "band1": (2000, 8000),
           "band2": (8000, 12000), "band3": (12000, hi), "high_shelf": (6000, hi)}
    eq = {}
    for b in EQ_BANDS:
        eq[f"{b}_gain_db"] = (eq_min_gain_db, eq_max_gain_db)
        eq[f"{b}_cutoff_freq"] = cut[b]
        eq[f"{b}_q_factor"] = (0.1, 5.0)
    rev = {f"band{i}_gain": (reverb_min_band_gain, reverb_max_band_gain) for i in range(12)}
    rev.update({f"band{i}_decay": (reverb_min_band_decay, reverb_max_band_decay) for i in range(12)})
    rev["mix"] = (0.0, 1.0)
    return {
        "input_fader": {"gain_db": (input_min_gain_db, input_max_gain_db)},
        "output_fader": {"gain_db": (output_min_gain_db, output_max_gain_db)},
        "parametric_eq": eq,
        "compressor": {"threshold_db": (-60.0, 0.0), "ratio": (1.0, 10.0),
                       "attack_ms": (5.0, 250.0), "release_ms": (10.0, 250.0),
                       "knee_db": (3.0, 12.0), "makeup_gain_db": (0.0, 6.0)},
        "reverberation": rev,
        "fx_bus": {"send_db": (min_send_db, max_send_db)},
        "stereo_panner": {"pan": (min_pan, max_pan)},
    }


def denormalize_parameters(param_dict, ranges):
    """mst/modules.py:79-97."""
    out = {}
    for effect, params in param_dict.items():
        out[effect] = {}
        for name, t in params.items():
            if t.min() < 0 or t.max() > 1:
                raise ValueError(f"Parameter {name} of effect {effect} is out of range.")
            lo, hi = ranges[effect][name]
            out[effect][name] = t * (hi - lo) + lo
    return out


def split_track_params(p):
    """(…, 27) -> nested dict, mst/modules.py:353-392."""
    d = {"input_fader": {"gain_db": p[..., 0]}, "parametric_eq": {}, "compressor": {}}
    for i, k in enumerate(EQ_KEYS):
        d["parametric_eq"][k] = p[..., 1 + i]
    for i, k in enumerate(COMP_KEYS):
        d["compressor"][k] = p[..., 19 + i]
    d["stereo_panner"] = {"pan": p[..., 25]}
    d["fx_bus"] = {"send_db": p[..., 26]}
    return d


def split_fx_params(p):
    """(…, 25) -> nested dict, mst/modules.py:394-422 (mix forced to ones)."""
    rev = {f"band{i}_gain": p[..., i] for i in range(12)}
    rev.update({f"band{i}_decay": p[..., 12 + i] for i in range(12)})
    rev["mix"] = torch.ones_like(p[..., 24])
    return {"reverberation": rev}


def split_master_params(p):
    """(…, 26) -> nested dict, mst/modules.py:424-460."""
    d = {"parametric_eq": {}, "compressor": {}}
    for i, k in enumerate(EQ_KEYS):
        d["parametric_eq"][k] = p[..., i]
    for i, k in enumerate(COMP_KEYS):
        d["compressor"][k] = p[..., 18 + i]
    d["output_fader"] = {"gain_db": p[..., 24]}
    d["input_fader"] = {"gain_db": p[..., 25]}
    return d


class OracleAdvancedMixConsole(torch.nn.Module):
    def __init__(self, sample_rate, **range_kwargs):
        super().__init__()
        self.sample_rate = sample_rate
        self.param_ranges = param_ranges(sample_rate, **range_kwargs)
        self.num_track_control_params = 27
        self.num_fx_bus_control_params = 25
        self.num_master_bus_control_params = 26

    def forward_mix_console(self, tracks, track_param_dict, fx_bus_param_dict,
                            master_bus_param_dict, use_track_input_fader=True,
                            use_track_eq=True, use_track_compressor=True,
                            use_track_panner=True, use_fx_bus=True, use_master_bus=True,
                            use_output_fader=True):
        bs, num_tracks, seq_len = tracks.shape
        sr = self.sample_rate
        x = tracks.reshape(-1, 1, seq_len)
        flat = lambda d: {k: v.reshape(-1) for k, v in d.items()}
        if use_track_input_fader:
            x = F.gain(x, sr, **flat(track_param_dict["input_fader"]))
        if use_track_eq:
            x = F.parametric_eq(x, sr, **flat(track_param_dict["parametric_eq"]))
        if use_track_compressor:
            x = F.compressor(x, sr, **flat(track_param_dict["compressor"]), lookahead_samples=2048)
        x = x.reshape(bs, num_tracks, seq_len)
        if not use_track_panner:
            # mst/modules.py:269 calls .repeat(1, 2, 1) on a 4-D tensor, which raises.
            raise RuntimeError("use_track_panner=False is broken in the reference (modules.py:269)")
        x = F.stereo_panner(x, sr, **track_param_dict["stereo_panner"])
        master = x.sum(dim=2)
        if use_fx_bus:
            raise NotImplementedError("fx bus is out of the oracle's scope")
        if use_master_bus:
            master = F.gain(master, sr, **master_bus_param_dict["input_fader"])
            master = F.parametric_eq(master, sr, **master_bus_param_dict["parametric_eq"])
            master = F.compressor(master, sr, **master_bus_param_dict["compressor"],
                                  lookahead_samples=1024)
        if use_output_fader:
            master = F.gain(master, sr, **master_bus_param_dict["output_fader"])
        return x, master

    def forward(self, tracks, track_params, fx_bus_params, master_bus_params,
                use_track_input_fader=True, use_track_eq=True, use_track_compressor=True,
                use_track_panner=True, use_master_bus=True, use_fx_bus=True,
                use_output_fader=True):
        tpd = denormalize_parameters(split_track_params(track_params), self.param_ranges)
        fpd = denormalize_parameters(split_fx_params(fx_bus_params), self.param_ranges)
        mpd = denormalize_parameters(split_master_params(master_bus_params), self.param_ranges)
        mixed, mix = self.forward_mix_console(
            tracks, tpd, fpd, mpd, use_track_input_fader=use_track_input_fader,
            use_track_eq=use_track_eq, use_track_compressor=use_track_compressor,
            use_track_panner=use_track_panner, use_fx_bus=use_fx_bus,
            use_master_bus=use_master_bus, use_output_fader=use_output_fader)
        return mixed, mix, tpd, fpd, mpd


class OracleBasicMixConsole(torch.nn.Module):
    """Reconstruction (the class is absent at the reference commit): per-track gain + pan,
    bus sum.  Two track parameters [gain_db, pan]; no fx / master parameters
    (mst/mixing.py:935-945)."""

    def __init__(self, sample_rate, min_gain_db=-48.0, max_gain_db=48.0, min_pan=0.0, max_pan=1.0):
        super().__init__()
        self.sample_rate = sample_rate
        self.param_ranges = {"input_gain": {"gain_db": (min_gain_db, max_gain_db)},
                             "stereo_panner": {"pan": (min_pan, max_pan)}}
        self.num_track_control_params = 2
        self.num_fx_bus_control_params = 0
        self.num_master_bus_control_params = 0

    def forward(self, tracks, track_params, fx_bus_params=None, master_bus_params=None, **flags):
        bs, num_tracks, seq_len = tracks.shape
        tpd = denormalize_parameters(
            {"input_gain": {"gain_db": track_params[..., 0]},
             "stereo_panner": {"pan": track_params[..., 1]}}, self.param_ranges)
        x = F.gain(tracks.reshape(-1, 1, seq_len), self.sample_rate,
                   tpd["input_gain"]["gain_db"].reshape(-1)).reshape(bs, num_tracks, seq_len)
        x = F.stereo_panner(x, self.sample_rate, tpd["stereo_panner"]["pan"])
        return x, x.sum(dim=2), tpd, {}, {}
