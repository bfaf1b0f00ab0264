"""Drop-in mix consoles backed by the sm_100a kernels.

``AdvancedMixConsole`` mirrors ``mst.modules.AdvancedMixConsole`` (reference
mst/modules.py:100-487): same constructor arguments, attributes (``sample_rate``,
``param_ranges``, ``num_*_control_params``), ``forward`` / ``forward_mix_console``
signatures (note that the two order ``use_master_bus`` / ``use_fx_bus`` differently, as
upstream does), return tuples, nested parameter-dict keys and the out-of-range
``ValueError`` (mst/modules.py:86-89).  The module has no parameters or buffers, so the
strict ``load_state_dict({})`` of mst/utils.py:245-249 keeps working.

``BasicMixConsole`` does not exist at the reference commit; it is reconstructed from
README.md:14 and mst/mixing.py:122-164, 935-945 (gain + pan per track, bus sum).

All arithmetic on audio runs in libdiffmst_b200.so; torch is used for memory, streams and
the autograd graph.  The denormalised parameter dictionaries returned to the caller are
produced with the same torch expressions as upstream (they are tiny and callers may
differentiate through them).
"""
import contextlib
import ctypes
from typing import Optional

import torch

from . import _lib

EQ_BANDS = ["low_shelf", "band0", "band1", "band2", "band3", "high_shelf"]
EQ_KEYS = [f"{b}_{p}" for b in EQ_BANDS for p in ("gain_db", "cutoff_freq", "q_factor")]
COMP_KEYS = ["threshold_db", "ratio", "attack_ms", "release_ms", "knee_db", "makeup_gain_db"]
TRACK_LOOKAHEAD = 2048   # mst/modules.py:250
MASTER_LOOKAHEAD = 1024  # mst/modules.py:304


def denormalize(norm_val, max_val, min_val):
    """mst/modules.py:71-72."""
    return (norm_val * (max_val - min_val)) + min_val


def _first_out_of_range(flat_params: torch.Tensor):
    """Index of the first column holding a value outside [0, 1], or None; one device sync
    instead of the reference's two per parameter (mst/modules.py:86)."""
    if flat_params.numel() == 0:
        return None
    bad = ((flat_params < 0) | (flat_params > 1)).any(dim=0)
    idx = torch.nonzero(bad)
    return int(idx[0]) if idx.numel() else None


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"diffmst_b200: {what} must be a CUDA tensor; there is no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"diffmst_b200: {what} must be float32 (got {t.dtype})")


class _ConsoleFunction(torch.autograd.Function):
    """tracks (B,N,T), track_params (B,N,P), master_params (B,26)|None -> mix, mixed."""

    @staticmethod
    def forward(ctx, tracks, track_params, master_params, ranges, sample_rate, flags, la_t, la_m,
                want_mixed, grad_enabled=True):
        lib = _lib.lib()
        _require_cuda(tracks, "tracks")
        _require_cuda(track_params, "track_params")
        B, N, T = tracks.shape
        if tracks.stride(2) != 1 or tracks.stride(0) < 0 or tracks.stride(1) < 0:
            tracks = tracks.contiguous()
        track_params = track_params.contiguous()
        if master_params is not None:
            _require_cuda(master_params, "master_bus_params")
            master_params = master_params.contiguous()
        dev = tracks.device
        flags = int(flags) | (_lib.WANT_MIXED_TRACKS if want_mixed else 0)
        # what backward will be asked for shapes what forward keeps: per-section checkpoints only when a gradient
        # w.r.t. the audio is wanted (the classic adjoint), nothing at all under torch.no_grad() (mst/mixing.py:72)
        # (needs_input_grad ignores torch.no_grad(); the caller passes the grad mode it was called under)
        need = [bool(grad_enabled) and n for n in ctx.needs_input_grad[:3]]
        if need[0]:
            flags |= _lib.WANT_GRAD_TRACKS
        if not any(need):
            flags |= _lib.FORWARD_ONLY
        with torch.cuda.device(dev):
            nbytes = lib.dmst_console_workspace_bytes(B, N, T, flags)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            mix = torch.empty(B, 2, T, dtype=torch.float32, device=dev)
            mixed = torch.empty(B, 2, N, T, dtype=torch.float32, device=dev) if want_mixed else None
            status = torch.empty(4, dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.dmst_console_forward(
                _ptr(tracks), tracks.stride(0), tracks.stride(1), _ptr(track_params), _ptr(master_params),
                ctypes.byref(ranges), float(sample_rate), B, N, T, flags, la_t, la_m, _ptr(mix),
                _ptr(mixed), _ptr(status), _ptr(ws), nbytes, ctypes.c_void_p(stream))
        _lib.check(rc, "dmst_console_forward")
        ctx.save_for_backward(tracks, track_params, master_params if master_params is not None else torch.empty(0, device=dev))
        ctx.has_master = master_params is not None
        ctx.ws, ctx.nbytes, ctx.ranges = ws, nbytes, ranges
        ctx.cfg = (float(sample_rate), flags, la_t, la_m)
        ctx.mark_non_differentiable(status)
        ctx.set_materialize_grads(False)   # an unused mixed_tracks output must not cost a (B,2,N,T) zero tensor
        if mixed is None:
            mixed = torch.empty(0, device=dev)
        return mix, mixed, status

    @staticmethod
    def backward(ctx, gmix, gmixed, _gstatus):
        lib = _lib.lib()
        tracks, track_params, master_params = ctx.saved_tensors
        master_params = master_params if ctx.has_master else None
        sample_rate, flags, la_t, la_m = ctx.cfg
        B, N, T = tracks.shape
        dev = tracks.device
        want_gtracks = bool(flags & _lib.WANT_GRAD_TRACKS)
        if gmix is None:
            gmix = torch.zeros(B, 2, T, dtype=torch.float32, device=dev)
        gmix = gmix.contiguous()
        use_gmixed = gmixed is not None and gmixed.numel() > 0 and (flags & _lib.WANT_MIXED_TRACKS)
        gmixed = gmixed.contiguous() if use_gmixed else None
        with torch.cuda.device(dev):
            gtp = torch.empty_like(track_params)
            gmp = torch.empty_like(master_params) if master_params is not None else None
            gtr = torch.empty(B, N, T, dtype=torch.float32, device=dev) if want_gtracks else None
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.dmst_console_backward(
                _ptr(tracks), tracks.stride(0), tracks.stride(1), _ptr(track_params), _ptr(master_params),
                ctypes.byref(ctx.ranges), sample_rate, B, N, T, flags, la_t, la_m, _ptr(gmix), _ptr(gmixed),
                _ptr(gtp), _ptr(gmp), _ptr(gtr), _ptr(ctx.ws), ctx.nbytes, ctypes.c_void_p(stream))
        _lib.check(rc, "dmst_console_backward")
        return gtr, gtp, gmp, None, None, None, None, None, None, None


def _flags(use_track_input_fader, use_track_eq, use_track_compressor, use_track_panner,
           use_fx_bus, use_master_bus, use_output_fader):
    f = 0
    f |= _lib.USE_TRACK_INPUT_FADER if use_track_input_fader else 0
    f |= _lib.USE_TRACK_EQ if use_track_eq else 0
    f |= _lib.USE_TRACK_COMPRESSOR if use_track_compressor else 0
    f |= _lib.USE_TRACK_PANNER if use_track_panner else 0
    f |= _lib.USE_MASTER_BUS if use_master_bus else 0
    f |= _lib.USE_OUTPUT_FADER if use_output_fader else 0
    if use_fx_bus:
        raise NotImplementedError(
            "use_fx_bus=True: the fx bus (stereo_bus + noise_shaped_reverberation, "
            "mst/modules.py:275-284) is outside the accelerated hot path; every shipped config "
            "keeps it off (configs/models/naive.yaml:7). Pass use_fx_bus=False.")
    if not use_track_panner:
        # upstream's else-branch calls .repeat(1, 2, 1) on a 4-D tensor (mst/modules.py:269)
        raise RuntimeError(
            "Number of dimensions of repeat dims can not be smaller than number of dimensions of "
            "tensor (use_track_panner=False is broken upstream, mst/modules.py:269)")
    return f


class AdvancedMixConsole(torch.nn.Module):
    def __init__(
        self,
        sample_rate: float,
        input_min_gain_db: float = -48.0,
        input_max_gain_db: float = 48.0,
        output_min_gain_db: float = -48.0,
        output_max_gain_db: float = 48.0,
        min_send_db: float = -80.0,
        max_send_db: float = +12.0,
        eq_min_gain_db: float = -12.0,
        eq_max_gain_db: float = 12.0,
        min_pan: float = 0.0,
        max_pan: float = 1.0,
        reverb_min_band_gain: float = 0.0,
        reverb_max_band_gain: float = 1.0,
        reverb_min_band_decay: float = 0.0,
        reverb_max_band_decay: float = 1.0,
    ):
        super().__init__()
        self.sample_rate = sample_rate
        hi_cut = (sample_rate // 2) - 1000
        cut = {"low_shelf": (20, 2000), "band0": (80, 2000), "band1": (2000, 8000),
               "band2": (8000, 12000), "band3": (12000, hi_cut), "high_shelf": (6000, hi_cut)}
        eq = {}
        for b in EQ_BANDS:
            eq[f"{b}_gain_db"] = (eq_min_gain_db, eq_max_gain_db)
            eq[f"{b}_cutoff_freq"] = cut[b]
            eq[f"{b}_q_factor"] = (0.1, 5.0)
        rev = {f"band{i}_gain": (reverb_min_band_gain, reverb_max_band_gain) for i in range(12)}
        rev.update({f"band{i}_decay": (reverb_min_band_decay, reverb_max_band_decay) for i in range(12)})
        rev["mix"] = (0.0, 1.0)
        self.param_ranges = {
            "input_fader": {"gain_db": (input_min_gain_db, input_max_gain_db)},
            "output_fader": {"gain_db": (output_min_gain_db, output_max_gain_db)},
            "parametric_eq": eq,
            "compressor": {"threshold_db": (-60.0, 0.0), "ratio": (1.0, 10.0), "attack_ms": (5.0, 250.0),
                           "release_ms": (10.0, 250.0), "knee_db": (3.0, 12.0), "makeup_gain_db": (0.0, 6.0)},
            "reverberation": rev,
            "fx_bus": {"send_db": (min_send_db, max_send_db)},
            "stereo_panner": {"pan": (min_pan, max_pan)},
        }
        self.num_track_control_params = 27
        self.num_fx_bus_control_params = 25
        self.num_master_bus_control_params = 26
        # Not part of the upstream interface: set False to skip materialising the
        # (bs, 2, num_tracks, seq_len) tensor that forward returns first (an empty tensor is
        # returned instead).
        self.materialize_tracks = True
        # Range test of mst/modules.py:86-89.  True: as upstream, the ValueError is raised inside the call (one
        # host sync per call instead of 156).  "async": the test runs on the device inside the console's own
        # prepare kernel (plus one tiny launch for the fx-bus block), the verdict is copied to pinned host memory
        # without synchronising, and the ValueError is raised by the first later forward() that finds it, or by
        # check_pending_ranges(); CUDA-graph capturable.  False: no test.
        self.check_ranges = True
        self._pending_ranges = []   # [(pinned int32 host tensor, event or None)]
        self._capture_slots = []    # pinned words reserved for calls made during a CUDA-graph capture
        self._side_streams = {}     # device -> stream the returned parameter dictionaries are computed on
        self.side_stream_dicts = True

    # ---- parameter plumbing (mst/modules.py:353-466) ----
    def _track_ranges(self):
        pr = self.param_ranges
        return ([pr["input_fader"]["gain_db"]] + [pr["parametric_eq"][k] for k in EQ_KEYS]
                + [pr["compressor"][k] for k in COMP_KEYS]
                + [pr["stereo_panner"]["pan"], pr["fx_bus"]["send_db"]])

    def _master_ranges(self):
        pr = self.param_ranges
        return ([pr["parametric_eq"][k] for k in EQ_KEYS] + [pr["compressor"][k] for k in COMP_KEYS]
                + [pr["output_fader"]["gain_db"], pr["input_fader"]["gain_db"]])

    def _c_ranges(self):
        r = _lib.Ranges()
        for i, (lo, hi) in enumerate(self._track_ranges()):
            r.track_lo[i], r.track_hi[i] = lo, hi
        for i, (lo, hi) in enumerate(self._master_ranges()):
            r.master_lo[i], r.master_hi[i] = lo, hi
        return r

    @staticmethod
    def _split_track(p):
        d = {"input_fader": {"gain_db": p[..., 0]},
             "parametric_eq": {k: p[..., 1 + i] for i, k in enumerate(EQ_KEYS)},
             "compressor": {k: p[..., 19 + i] for i, k in enumerate(COMP_KEYS)},
             "stereo_panner": {"pan": p[..., 25]},
             "fx_bus": {"send_db": p[..., 26]}}
        return d

    @staticmethod
    def _split_fx(p):
        rev = {f"band{i}_gain": p[..., i] for i in range(12)}
        rev.update({f"band{i}_decay": p[..., 12 + i] for i in range(12)})
        rev["mix"] = torch.ones_like(p[..., 24])
        return {"reverberation": rev}

    @staticmethod
    def _split_master(p):
        return {"parametric_eq": {k: p[..., i] for i, k in enumerate(EQ_KEYS)},
                "compressor": {k: p[..., 18 + i] for i, k in enumerate(COMP_KEYS)},
                "output_fader": {"gain_db": p[..., 24]},
                "input_fader": {"gain_db": p[..., 25]}}

    def _denormalize(self, param_dict):
        """Reference form (mst/modules.py:79-97), one affine per entry; kept for callers that
        hand in dictionaries.  forward() uses the batched equivalent below."""
        out = {}
        for effect, params in param_dict.items():
            out[effect] = {}
            for name, t in params.items():
                lo, hi = self.param_ranges[effect][name]
                out[effect][name] = denormalize(t, hi, lo)
        return out

    def _affine(self, ranges, like):
        """(scale, offset) rows for a whole parameter vector.  Built as float32 from the same
        Python floats upstream multiplies/adds with, so `p * scale + offset` is bit-identical to
        the per-entry `(p * (hi - lo)) + lo` of mst/modules.py:71-72."""
        key = (tuple(ranges), like.device)
        cache = self.__dict__.setdefault("_affine_cache", {})
        if key not in cache:
            scale = torch.tensor([hi - lo for lo, hi in ranges], dtype=torch.float32, device=like.device)
            offset = torch.tensor([lo for lo, hi in ranges], dtype=torch.float32, device=like.device)
            cache[key] = (scale, offset)
        return cache[key]

    def _denormalize_batched(self, params, ranges, split):
        scale, offset = self._affine(ranges, params)
        return split(params * scale + offset)

    def _raise_if_out_of_range(self, track_params, fx_bus_params, master_bus_params):
        # same traversal order as three denormalize_parameters calls (mst/modules.py:462-466)
        flat = torch.cat([track_params.reshape(-1, 27).amin(0), track_params.reshape(-1, 27).amax(0),
                          fx_bus_params.reshape(-1, 25).amin(0), fx_bus_params.reshape(-1, 25).amax(0),
                          master_bus_params.reshape(-1, 26).amin(0), master_bus_params.reshape(-1, 26).amax(0)])
        flat = flat.detach().cpu()
        tmin, tmax, fmin, fmax, mmin, mmax = torch.split(flat, [27, 27, 25, 25, 26, 26])
        track_names = [("input_fader", "gain_db")] + [("parametric_eq", k) for k in EQ_KEYS] + \
            [("compressor", k) for k in COMP_KEYS] + [("stereo_panner", "pan"), ("fx_bus", "send_db")]
        fx_names = [("reverberation", f"band{i}_gain") for i in range(12)] + \
            [("reverberation", f"band{i}_decay") for i in range(12)]  # "mix" is forced to ones
        master_names = [("parametric_eq", k) for k in EQ_KEYS] + [("compressor", k) for k in COMP_KEYS] + \
            [("output_fader", "gain_db"), ("input_fader", "gain_db")]
        for names, lo, hi in ((track_names, tmin, tmax), (fx_names, fmin, fmax), (master_names, mmin, mmax)):
            for i, (effect, name) in enumerate(names):
                if lo[i] < 0 or hi[i] > 1:
                    raise ValueError(f"Parameter {name} of effect {effect} is out of range.")

    # ---- asynchronous range test (check_ranges = "async") ----
    _TRACK_NAMES = [("input_fader", "gain_db")] + [("parametric_eq", k) for k in EQ_KEYS] + \
        [("compressor", k) for k in COMP_KEYS] + [("stereo_panner", "pan"), ("fx_bus", "send_db")]
    _FX_NAMES = [("reverberation", f"band{i}_gain") for i in range(12)] + \
        [("reverberation", f"band{i}_decay") for i in range(12)] + [("reverberation", "mix")]
    _MASTER_NAMES = [("parametric_eq", k) for k in EQ_KEYS] + [("compressor", k) for k in COMP_KEYS] + \
        [("output_fader", "gain_db"), ("input_fader", "gain_db")]

    @classmethod
    def _status_error(cls, code: int):
        """Device status word (include/diffmst_b200.h) -> the reference's ValueError, or None."""
        if code == _lib.STATUS_OK or code <= 0:
            return None
        if code > 1000:
            effect, name = cls._MASTER_NAMES[code - 1001]
        elif code > 500:
            effect, name = cls._FX_NAMES[code - 501]
        else:
            effect, name = cls._TRACK_NAMES[code - 1]
        return ValueError(f"Parameter {name} of effect {effect} is out of range.")

    def _queue_range_status(self, status, fx_bus_params):
        lib = _lib.lib()
        dev = status.device
        capturing = torch.cuda.is_current_stream_capturing()
        with torch.cuda.device(dev):
            if capturing:
                # pinned memory cannot be allocated while a stream is capturing: take a slot reserved beforehand
                if not self._capture_slots:
                    raise RuntimeError("AdvancedMixConsole(check_ranges='async') inside a CUDA-graph capture needs "
                                       "console.reserve_capture_slots(n) before the capture (GraphedStep does this)")
                host = self._capture_slots.pop()
            else:
                host = torch.empty(1, dtype=torch.int32, pin_memory=True)
            host.fill_(_lib.STATUS_OK)
            fx, rows, np_ = None, 0, 0
            if fx_bus_params is not None and fx_bus_params.numel():
                # 24 tested columns: upstream forces "mix" (column 24) to ones before the test (mst/modules.py:420)
                fx = fx_bus_params.detach().reshape(-1, fx_bus_params.shape[-1])
                if fx.stride(1) != 1 or fx.dtype != torch.float32:
                    fx = fx.float().contiguous()
                rows, np_ = fx.shape[0], min(24, fx.shape[1])
            # one kernel: the fx-bus test and the verdict stored straight into the pinned host word
            _lib.check(lib.dmst_console_report_ranges(_ptr(fx), fx.stride(0) if fx is not None else 0, rows, np_, 500,
                                                      _ptr(status), ctypes.c_void_p(host.data_ptr()),
                                                      ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "dmst_console_report_ranges")
            ev = None
            if not capturing:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
        self._pending_ranges.append((host, ev))
        if len(self._pending_ranges) > 64 and not capturing:   # bounded: the oldest verdicts have long landed
            self.check_pending_ranges(wait=False)
            if len(self._pending_ranges) > 64:
                self.check_pending_ranges(wait=True)

    def reserve_capture_slots(self, n: int = 8):
        """Pinned host words for the verdicts of up to n forward() calls inside a CUDA-graph capture."""
        while len(self._capture_slots) < n:
            self._capture_slots.append(torch.empty(1, dtype=torch.int32, pin_memory=True))

    def check_pending_ranges(self, wait: bool = True):
        """Raise the reference's ValueError (mst/modules.py:86-89) for the oldest earlier forward() whose parameters
        were out of range (check_ranges="async").  wait=True synchronises with the device first; wait=False looks only
        at verdicts that have already landed in host memory."""
        if wait and self._pending_ranges:
            torch.cuda.synchronize()
        keep, err = [], None
        for host, ev in self._pending_ranges:
            # ev is None: recorded during CUDA-graph capture; every replay rewrites that host word, so the entry
            # (and its pinned buffer) lives as long as the console and is only read after a synchronisation
            done = wait or (ev is not None and ev.query())
            if ev is None or not done:
                keep.append((host, ev))
            if done and err is None:
                err = self._status_error(int(host[0]))
                if ev is None:
                    host.fill_(_lib.STATUS_OK)
        self._pending_ranges = keep
        if err is not None:
            self._pending_ranges = [e for e in keep if e[1] is None]
            raise err

    # ---- the chain (mst/modules.py:186-314) ----
    def _run(self, tracks, track_params_norm, master_params_norm, flags, ranges=None, fx_bus_params=None):
        mix, mixed, status = _ConsoleFunction.apply(
            tracks, track_params_norm, master_params_norm, ranges if ranges is not None else self._c_ranges(),
            self.sample_rate, flags, TRACK_LOOKAHEAD, MASTER_LOOKAHEAD, self.materialize_tracks, torch.is_grad_enabled())
        if self.check_ranges == "async" and ranges is None:
            self._queue_range_status(status, fx_bus_params)
        return mixed, mix

    @staticmethod
    def _identity_ranges():
        r = _lib.Ranges()
        for i in range(_lib.NUM_TRACK_PARAMS):
            r.track_lo[i], r.track_hi[i] = 0.0, 1.0
        for i in range(_lib.NUM_MASTER_PARAMS):
            r.master_lo[i], r.master_hi[i] = 0.0, 1.0
        return r

    @staticmethod
    def _stack_dict(denorm_dict, keys, like):
        cols = []
        for effect, name in keys:
            v = denorm_dict[effect][name]
            cols.append(v if torch.is_tensor(v) else torch.as_tensor(v, dtype=like.dtype, device=like.device))
        cols = torch.broadcast_tensors(*cols)
        return torch.stack(cols, dim=-1).to(like.dtype)

    def forward_mix_console(
        self,
        tracks: torch.Tensor,
        track_param_dict: dict,
        fx_bus_param_dict: dict,
        master_bus_param_dict: dict,
        use_track_input_fader: bool = True,
        use_track_eq: bool = True,
        use_track_compressor: bool = True,
        use_track_panner: bool = True,
        use_fx_bus: bool = True,
        use_master_bus: bool = True,
        use_output_fader: bool = True,
    ):
        """Same contract as mst/modules.py:186-314: takes DENORMALISED parameter dicts and
        returns (tracks (bs, 2, num_tracks, seq_len), master_bus (bs, 2, seq_len)).  As upstream, the values are
        applied as given: no range test and no clamping (mst/mixing.py's knowledge-engineering mix passes values
        outside ``param_ranges``).  The kernels denormalise `p * (hi - lo) + lo` in float64 themselves, so the
        denormalised values go to them with the identity range (lo, hi) = (0, 1): exact, and gradients arrive with
        respect to the dictionary entries."""
        flags = _flags(use_track_input_fader, use_track_eq, use_track_compressor, use_track_panner,
                       use_fx_bus, use_master_bus, use_output_fader)
        tp = self._stack_dict(track_param_dict, self._TRACK_NAMES, tracks)
        mp = self._stack_dict(master_bus_param_dict, self._MASTER_NAMES, tracks)
        return self._run(tracks, tp, mp, flags, ranges=self._identity_ranges())

    def forward(
        self,
        tracks: torch.Tensor,
        track_params: torch.Tensor,
        fx_bus_params: torch.Tensor,
        master_bus_params: torch.Tensor,
        use_track_input_fader: bool = True,
        use_track_eq: bool = True,
        use_track_compressor: bool = True,
        use_track_panner: bool = True,
        use_master_bus: bool = True,
        use_fx_bus: bool = True,
        use_output_fader: bool = True,
    ):
        """Create a mix from tracks and mixing parameters in (0, 1); mst/modules.py:316-487."""
        flags = _flags(use_track_input_fader, use_track_eq, use_track_compressor, use_track_panner,
                       use_fx_bus, use_master_bus, use_output_fader)
        if self.check_ranges == "async":
            if not torch.cuda.is_current_stream_capturing():   # (event queries are not allowed during a capture)
                self.check_pending_ranges(wait=False)          # verdicts of earlier calls that have landed
        elif self.check_ranges:
            self._raise_if_out_of_range(track_params, fx_bus_params, master_bus_params)
        # The returned dictionaries of denormalised values (two elementwise kernels per tensor instead of two per
        # dictionary entry, 156 upstream) do not feed the audio path: they are computed on a side stream, beside the
        # console kernels, and joined before this call returns (in a captured CUDA graph: a parallel branch).
        fx_ranges = [self.param_ranges["reverberation"][f"band{i}_gain"] for i in range(12)] + \
            [self.param_ranges["reverberation"][f"band{i}_decay"] for i in range(12)] + \
            [self.param_ranges["reverberation"]["mix"]]
        side = cur = ready = None
        if tracks.is_cuda and self.side_stream_dicts:
            cur = torch.cuda.current_stream(tracks.device)
            side = self._side_streams.get(tracks.device)
            if side is None:
                side = self._side_streams[tracks.device] = torch.cuda.Stream(tracks.device)
            ready = cur.record_event()   # the parameters are ready here; the side stream need not wait for the console
        # (the console call comes first: the parameters' first differentiable use, hence their gradient accumulation,
        # stays on the caller's stream)
        mixed_tracks, mix = self._run(tracks, track_params, master_bus_params, flags, fx_bus_params=fx_bus_params)
        if side is not None:
            side.wait_event(ready)
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            track_param_dict = self._denormalize_batched(track_params, self._track_ranges(), self._split_track)
            fx_bus_param_dict = self._denormalize_batched(fx_bus_params, fx_ranges, self._split_fx)
            master_bus_param_dict = self._denormalize_batched(master_bus_params, self._master_ranges(),
                                                              self._split_master)
            if side is not None:   # allocated on the side stream, consumed on the caller's (the entries of a dictionary
                for d in (track_param_dict, fx_bus_param_dict, master_bus_param_dict):   # are views of one tensor)
                    seen = set()
                    for effect in d.values():
                        for v in effect.values():
                            if torch.is_tensor(v):
                                base = v._base if v._base is not None else v
                                if id(base) not in seen:
                                    seen.add(id(base))
                                    base.record_stream(cur)
        if side is not None:
            cur.wait_stream(side)
        return mixed_tracks, mix, track_param_dict, fx_bus_param_dict, master_bus_param_dict


class BasicMixConsole(torch.nn.Module):
    """Gain + pan per track, stereo bus sum (reconstruction, see module docstring).
    Track parameters: [..., 0] gain_db, [..., 1] pan, both normalised to (0, 1)."""

    def __init__(self, sample_rate: float, min_gain_db: float = -48.0, max_gain_db: float = 48.0,
                 min_pan: float = 0.0, max_pan: float = 1.0):
        super().__init__()
        self.sample_rate = sample_rate
        self.param_ranges = {"input_gain": {"gain_db": (min_gain_db, max_gain_db)},
                             "stereo_panner": {"pan": (min_pan, max_pan)}}
        self.num_track_control_params = 2
        self.num_fx_bus_control_params = 0
        self.num_master_bus_control_params = 0
        self.materialize_tracks = True
        self.check_ranges = True

    def _c_ranges(self):
        r = _lib.Ranges()
        r.track_lo[0], r.track_hi[0] = self.param_ranges["input_gain"]["gain_db"]
        r.track_lo[25], r.track_hi[25] = self.param_ranges["stereo_panner"]["pan"]
        return r

    def forward(self, tracks, track_params, fx_bus_params=None, master_bus_params=None, **flags):
        if self.check_ranges is True:   # (synchronous test only; "async" is the advanced console's device-side form)
            lo = track_params.reshape(-1, 2).amin(0).detach().cpu()
            hi = track_params.reshape(-1, 2).amax(0).detach().cpu()
            for i, (effect, name) in enumerate((("input_gain", "gain_db"), ("stereo_panner", "pan"))):
                if lo[i] < 0 or hi[i] > 1:
                    raise ValueError(f"Parameter {name} of effect {effect} is out of range.")
        pr = self.param_ranges
        track_param_dict = {
            "input_gain": {"gain_db": denormalize(track_params[..., 0], pr["input_gain"]["gain_db"][1],
                                                  pr["input_gain"]["gain_db"][0])},
            "stereo_panner": {"pan": denormalize(track_params[..., 1], pr["stereo_panner"]["pan"][1],
                                                 pr["stereo_panner"]["pan"][0])}}
        f = (_lib.BASIC_CONSOLE | _lib.USE_TRACK_INPUT_FADER | _lib.USE_TRACK_PANNER)
        mix, mixed, _ = _ConsoleFunction.apply(tracks, track_params, None, self._c_ranges(),
                                               self.sample_rate, f, 0, 0, self.materialize_tracks, torch.is_grad_enabled())
        return mixed, mix, track_param_dict, {}, {}
