"""ctypes binding of libdiffmst_b200.so (include/diffmst_b200.h).

The library is the only compute path: if it is missing, or is not a device build, every
entry point raises.  There is no CPU or PyTorch fallback."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdiffmst_b200.so")

NUM_TRACK_PARAMS, NUM_FX_PARAMS, NUM_MASTER_PARAMS = 27, 25, 26
USE_TRACK_INPUT_FADER, USE_TRACK_EQ, USE_TRACK_COMPRESSOR, USE_TRACK_PANNER = 1, 2, 4, 8
USE_MASTER_BUS, USE_FX_BUS, USE_OUTPUT_FADER = 16, 32, 64
WANT_MIXED_TRACKS, WANT_GRAD_TRACKS, BASIC_CONSOLE, FORWARD_ONLY = 128, 256, 512, 1024
EINVAL = -22
STATUS_OK = 0x7F7F7F7F
MRSTFT_MAX_RES = 8


class Ranges(ctypes.Structure):
    _fields_ = [("track_lo", ctypes.c_float * NUM_TRACK_PARAMS), ("track_hi", ctypes.c_float * NUM_TRACK_PARAMS),
                ("master_lo", ctypes.c_float * NUM_MASTER_PARAMS), ("master_hi", ctypes.c_float * NUM_MASTER_PARAMS)]


class MrstftCfg(ctypes.Structure):
    _fields_ = [("n_res", ctypes.c_int), ("fft_size", ctypes.c_int * MRSTFT_MAX_RES),
                ("hop_size", ctypes.c_int * MRSTFT_MAX_RES), ("win_length", ctypes.c_int * MRSTFT_MAX_RES),
                ("w_sc", ctypes.c_float), ("w_log_mag", ctypes.c_float), ("w_lin_mag", ctypes.c_float),
                ("eps", ctypes.c_float)]


_lib = None


def lib():
    """Load (once) and return the CUDA library; raise loudly if it is unusable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the diffmst_b200 CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU fallback.")
    l = ctypes.CDLL(LIB_PATH)
    l.dmst_is_device_build.restype = ctypes.c_int
    if l.dmst_is_device_build() != 1:
        raise ImportError(f"{LIB_PATH} is not a CUDA device build; refusing to use it")
    vp, i, u, ll, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t
    l.dmst_version.restype = i
    l.dmst_console_workspace_bytes.restype = sz
    l.dmst_console_workspace_bytes.argtypes = [i, i, i, u]
    l.dmst_console_forward.restype = i
    l.dmst_console_forward.argtypes = [vp, ll, ll, vp, vp, ctypes.POINTER(Ranges), f, i, i, i, u, i, i,
                                       vp, vp, vp, vp, sz, vp]
    l.dmst_console_backward.restype = i
    l.dmst_console_backward.argtypes = [vp, ll, ll, vp, vp, ctypes.POINTER(Ranges), f, i, i, i, u, i, i,
                                        vp, vp, vp, vp, vp, vp, sz, vp]
    l.dmst_console_check_ranges.restype = i
    l.dmst_console_check_ranges.argtypes = [vp, i, i, i, vp, vp]
    l.dmst_console_report_ranges.restype = i
    l.dmst_console_report_ranges.argtypes = [vp, ll, i, i, i, vp, vp, vp]
    l.dmst_ola_hann_add.restype = i
    l.dmst_ola_hann_add.argtypes = [vp, ll, vp, ll, i, i, i, i, vp]
    l.dmst_mrstft_workspace_bytes.restype = sz
    l.dmst_mrstft_workspace_bytes.argtypes = [ctypes.POINTER(MrstftCfg), i, i]
    l.dmst_mrstft_forward.restype = i
    l.dmst_mrstft_forward.argtypes = [vp, ll, vp, ll, vp, ctypes.POINTER(MrstftCfg), i, i, vp, vp, vp, sz, vp]
    l.dmst_mrstft_forward_keep.restype = i
    l.dmst_mrstft_forward_keep.argtypes = [vp, ll, vp, ll, vp, ctypes.POINTER(MrstftCfg), i, i, vp, vp, vp, sz, vp]
    l.dmst_mrstft_backward.restype = i
    l.dmst_mrstft_backward.argtypes = [vp, ctypes.POINTER(MrstftCfg), i, i, vp, vp, vp, sz, vp]
    l.dmst_afl_workspace_bytes.restype = sz
    l.dmst_afl_workspace_bytes.argtypes = [i, i, i, i]
    fp5 = ctypes.POINTER(ctypes.c_float)
    l.dmst_afl_forward.restype = i
    l.dmst_afl_forward.argtypes = [vp, vp, ll, ll, vp, vp, fp5, i, i, i, i, vp, vp, sz, vp]
    l.dmst_afl_backward.restype = i
    l.dmst_afl_backward.argtypes = [vp, ll, ll, vp, vp, fp5, vp, i, i, i, i, vp, vp, sz, vp]
    l.dmst_peak_normalize.restype = i
    l.dmst_peak_normalize.argtypes = [vp, ll, ll, vp, i, i, vp]
    l.dmst_conv_nchw_to_padded_nhwc.restype = i
    l.dmst_conv_nchw_to_padded_nhwc.argtypes = [vp, vp, i, i, i, i, vp]
    l.dmst_conv_round_tf32.restype = i
    l.dmst_conv_round_tf32.argtypes = [vp, vp, ll, i, i, i, i, vp]
    l.dmst_conv_repack_weights.restype = i
    l.dmst_conv_repack_weights.argtypes = [vp, vp, i, i, vp]
    l.dmst_conv_repack_weights_dgrad.restype = i
    l.dmst_conv_repack_weights_dgrad.argtypes = [vp, vp, i, i, vp]
    l.dmst_conv3x3_forward.restype = i
    l.dmst_conv3x3_forward.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    l.dmst_conv3x3_workspace_bytes.restype = sz
    l.dmst_conv3x3_workspace_bytes.argtypes = [i, i, i, i, i]
    l.dmst_conv3x3_forward_ws.restype = i
    l.dmst_conv3x3_forward_ws.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, sz, vp]
    l.dmst_conv_stats_workspace_bytes.restype = sz
    l.dmst_conv_stats_workspace_bytes.argtypes = [i, i, i, i]
    l.dmst_conv_channel_stats.restype = i
    l.dmst_conv_channel_stats.argtypes = [vp, i, i, i, i, vp, vp, vp, sz, vp]
    l.dmst_conv_affine_relu.restype = i
    l.dmst_conv_affine_relu.argtypes = [vp, vp, vp, i, i, i, i, i, vp]
    l.dmst_conv_avgpool.restype = i
    l.dmst_conv_avgpool.argtypes = [vp, vp, i, i, i, i, i, i, i, vp]
    l.dmst_conv3x3_wgrad_workspace_bytes.restype = sz
    l.dmst_conv3x3_wgrad_workspace_bytes.argtypes = [i, i, i, i, i]
    l.dmst_conv3x3_wgrad.restype = i
    l.dmst_conv3x3_wgrad.argtypes = [vp, vp, vp, i, i, i, i, i, vp, sz, vp]
    l.dmst_conv_affine_relu_to.restype = i
    l.dmst_conv_affine_relu_to.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    l.dmst_conv_bn_relu_backward.restype = i
    l.dmst_conv_bn_relu_backward.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, sz, vp]
    l.dmst_conv_avgpool_backward.restype = i
    l.dmst_conv_avgpool_backward.argtypes = [vp, vp, i, i, i, i, i, i, i, vp]
    l.dmst_conv_bn_relu_avgpool.restype = i
    l.dmst_conv_bn_relu_avgpool.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
    l.dmst_conv_bn_relu_avgpool_backward.restype = i
    l.dmst_conv_bn_relu_avgpool_backward.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, sz, vp]
    l.dmst_spectrogram_workspace_bytes.restype = sz
    l.dmst_spectrogram_workspace_bytes.argtypes = [i, i, i, i, i]
    l.dmst_spectrogram_frontend.restype = i
    l.dmst_spectrogram_frontend.argtypes = [vp, ll, vp, i, i, i, i, i, f, f, vp, vp, sz, vp]
    l.dmst_profile_enable.restype = i
    l.dmst_profile_enable.argtypes = [i]
    l.dmst_profile_read.restype = i
    l.dmst_profile_read.argtypes = [i, ctypes.POINTER(ctypes.c_float), i]
    _lib = l
    return l


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc == EINVAL:
        raise ValueError(f"{what}: invalid argument (DMST_EINVAL)")
    raise RuntimeError(f"{what}: CUDA error {rc}")
