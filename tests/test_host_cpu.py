"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares,
host-side parameter plumbing, and that the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from diffmst_b200 import _lib
    header = open(os.path.join(ROOT, "include", "diffmst_b200.h")).read()
    names = sorted(set(re.findall(r"\b(dmst_[a-z0-9_]+)\s*\(", header)))
    assert "dmst_console_forward" in names and "dmst_console_backward" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/diffmst_b200.h but not exported"
    assert lib.dmst_is_device_build() == 1
    lib.dmst_console_workspace_bytes.restype = ctypes.c_size_t
    assert lib.dmst_console_workspace_bytes(8, 16, 262144, 0) > 8 * 16 * 262144 * 4


def test_no_cpu_fallback():
    from diffmst_b200 import AdvancedMixConsole
    con = AdvancedMixConsole(44100)
    x = torch.randn(1, 2, 4096)
    with pytest.raises(RuntimeError, match="no CPU path"):
        con(x, torch.rand(1, 2, 27), torch.rand(1, 25), torch.rand(1, 26), use_fx_bus=False)


def test_param_ranges_and_dict_layout_match_reference_tables():
    from diffmst_b200 import AdvancedMixConsole, BasicMixConsole
    from oracle.console import param_ranges
    con = AdvancedMixConsole(44100, eq_min_gain_db=-6.0)
    assert con.param_ranges == param_ranges(44100, eq_min_gain_db=-6.0)
    assert (con.num_track_control_params, con.num_fx_bus_control_params, con.num_master_bus_control_params) == (27, 25, 26)
    assert con.param_ranges["parametric_eq"]["band3_cutoff_freq"] == (12000, 21050)
    p = torch.rand(2, 3, 27)
    d = con._denormalize(con._split_track(p))
    assert torch.equal(d["compressor"]["ratio"], p[..., 20] * 9.0 + 1.0)
    assert torch.equal(d["fx_bus"]["send_db"], p[..., 26] * (12.0 - -80.0) + -80.0)
    m = con._denormalize(con._split_master(torch.rand(2, 26)))
    assert list(m.keys()) == ["parametric_eq", "compressor", "output_fader", "input_fader"]
    with pytest.raises(ValueError, match="Parameter low_shelf_q_factor of effect parametric_eq is out of range."):
        bad = p.clone(); bad[1, 2, 3] = 1.01
        con._raise_if_out_of_range(bad, torch.rand(2, 25), torch.rand(2, 26))
    with pytest.raises(ValueError, match="Parameter band11_decay of effect reverberation is out of range."):
        fx = torch.rand(2, 25); fx[0, 23] = -1.0
        con._raise_if_out_of_range(p, fx, torch.rand(2, 26))
    con._raise_if_out_of_range(p, torch.rand(2, 25), torch.rand(2, 26))
    assert len(con.state_dict()) == 0 and len(BasicMixConsole(44100).state_dict()) == 0
    r = con._c_ranges()
    assert (r.track_lo[20], r.track_hi[20]) == (1.0, 10.0) and (r.master_lo[25], r.master_hi[25]) == (-48.0, 48.0)


def test_host_side_sizing_and_refusals_of_the_encoder_path():
    """Host arithmetic of the convolution entry points (no GPU work): the weight-gradient workspace follows the
    split plan (about two CTAs per SM, whole 32-row stages), unsupported channel counts report 0 so that the caller
    takes the library arm, and the graph / encoder front ends refuse CPU tensors."""
    from diffmst_b200 import GraphedStep, SpectrogramEncoder, _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    f = lib.dmst_conv3x3_wgrad_workspace_bytes
    f.restype = ctypes.c_size_t
    f.argtypes = [ctypes.c_int] * 5
    # block1.conv2 at batch 4: 64 -> 64, one 128 x 64 tile, two tap groups -> 148 pixel splits of partial sums
    P = 4 * 1027 * 259
    n = f(4, 1025, 257, 64, 64)
    assert n % (9 * 64 * 64 * 4) == 0
    splits = n // (9 * 64 * 64 * 4)
    assert 140 <= splits <= 148 and splits * ((P + splits - 1) // splits + 31) >= P
    # the deepest layer: many tiles, one split
    assert f(4, 2, 4, 2048, 2048) == 9 * 2048 * 2048 * 4
    # first layer (streaming kernel), and shapes neither kernel covers
    assert f(4, 1025, 257, 1, 64) > 0
    assert f(1, 8, 8, 3, 64) == 0 and f(1, 8, 8, 48, 40) == 0 and f(0, 8, 8, 64, 64) == 0
    with pytest.raises(RuntimeError, match="no CPU path"):
        GraphedStep(lambda: None, params=[torch.zeros(1, requires_grad=True)])
    with pytest.raises(ValueError, match="at least one parameter"):
        GraphedStep(lambda: None, params=[])
    enc = SpectrogramEncoder(embed_dim=8)
    assert set(k.split(".")[0] for k in enc.state_dict()) == {"window", "model"}
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.zeros(1, 1, 40000))
