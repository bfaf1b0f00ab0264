"""GPU parity of the tensor-core ConvBlock / Cnn14 forward against a plain PyTorch float32
reference of the same op (TF32 disabled in the reference).  Tolerance: TF32 operands carry a
10-bit mantissa (2^-11 relative rounding per operand), accumulation is FP32; we require 3e-3 of the
output's max magnitude, the same class as cuDNN's default TF32 convolutions."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.panns import OracleCnn14, OracleConvBlock

pytestmark = pytest.mark.gpu
TOL = 3e-3


def relmax(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _randomise_bn(block, g):
    for bn in (block.bn1, block.bn2):
        if isinstance(bn, torch.nn.BatchNorm2d):
            bn.weight.data = torch.rand(bn.num_features, generator=g) + 0.5
            bn.bias.data = torch.randn(bn.num_features, generator=g) * 0.1
            bn.running_mean.data = torch.randn(bn.num_features, generator=g) * 0.1
            bn.running_var.data = torch.rand(bn.num_features, generator=g) + 0.5


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("cin,cout,shape,pool", [(64, 64, (2, 37, 29), (2, 2)), (64, 128, (3, 33, 20), (4, 4)),
                                                 (128, 256, (1, 40, 16), (4, 2)), (1, 64, (2, 50, 31), (2, 2))])
def test_conv_block_eval(cin, cout, shape, pool):
    from diffmst_b200 import ConvBlock
    g = torch.Generator().manual_seed(cin * 7 + cout)
    ref = OracleConvBlock(cin, cout)
    _randomise_bn(ref, g)
    ref = ref.cuda().eval()
    ours = ConvBlock(cin, cout).cuda().eval()
    ours.load_state_dict(ref.state_dict(), strict=True)   # same parameter / buffer names as the reference
    x = torch.randn(shape[0], cin, shape[1], shape[2], generator=g).cuda()
    with torch.no_grad():
        want = ref(x, pool)
        got = ours(x, pool)
    assert got.shape == want.shape
    assert relmax(got, want) <= TOL, relmax(got, want)


def test_conv_block_train_mode_batchnorm_and_errors():
    from diffmst_b200 import ConvBlock
    g = torch.Generator().manual_seed(5)
    ref = OracleConvBlock(64, 128).cuda().train()
    ours = ConvBlock(64, 128).cuda().train()
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(4, 64, 24, 18, generator=g).cuda()
    with torch.no_grad():
        want = ref(x, (2, 2))
        got = ours(x, (2, 2))
    assert relmax(got, want) <= TOL
    assert torch.allclose(ours.bn1.running_mean, ref.bn1.running_mean, atol=2e-3)
    assert torch.allclose(ours.bn2.running_var, ref.bn2.running_var, rtol=5e-3, atol=1e-4)
    assert int(ours.bn1.num_batches_tracked) == 1
    with pytest.raises(RuntimeError, match="no CPU path"):
        with torch.no_grad():
            ours(x.cpu(), (2, 2))


@pytest.mark.parametrize("cin,cout,shape,pool,train", [(64, 64, (2, 21, 18), (2, 2), True), (64, 128, (3, 24, 20), (4, 2), True),
                                                       (1, 64, (2, 30, 17), (2, 2), True), (128, 128, (2, 16, 12), (2, 2), False)])
def test_conv_block_backward(cin, cout, shape, pool, train):
    """Gradients of the tensor-core ConvBlock (dgrad on tcgen05, wgrad as library GEMMs, BatchNorm in
    batch-statistics or running-statistics mode) against autograd of the float32 reference module."""
    from diffmst_b200 import ConvBlock
    g = torch.Generator().manual_seed(cin + 3 * cout)
    ref = OracleConvBlock(cin, cout)
    _randomise_bn(ref, g)
    ref = ref.cuda().train(train)
    ours = ConvBlock(cin, cout).cuda().train(train)
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(shape[0], cin, shape[1], shape[2], generator=g).cuda()
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    want = ref(xr, pool)
    got = ours(xo, pool)
    assert relmax(got, want) <= TOL
    probe = torch.randn(want.shape, generator=g).cuda()
    (want * probe).sum().backward()
    (got * probe).sum().backward()
    # TF32 rounding flips the ReLU mask of activations that are within rounding of zero, which changes single
    # gradient entries completely (any reduced-precision convolution shows this against an FP32 reference), so the
    # gradients are compared in relative L2 norm; the convolution gradients alone are exact to TF32 rounding
    # (scripts/debug_conv_bwd.py: 7e-4 of the maximum for dgrad, 3e-7 for the FP32 wgrad GEMMs)
    def rl2(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    GT = 6e-2   # ~0.3 % of the activations sit within TF32 rounding of zero: sqrt(0.003) ~ 5 % in L2
    assert rl2(xo.grad, xr.grad) <= GT, rl2(xo.grad, xr.grad)
    for (n, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert po.grad is not None, n
        assert rl2(po.grad, pr.grad) <= GT, (n, rl2(po.grad, pr.grad))


def test_conv3x3_function_gradients_single_layer():
    """dgrad (tcgen05 kernel on the flipped taps) and wgrad (GEMMs over the flattened zero-bordered layout) of one
    convolution against torch's conv2d autograd: no ReLU in between, so element-wise tolerances apply."""
    import torch.nn.functional as F
    from diffmst_b200.conv import _Conv3x3Function
    g = torch.Generator().manual_seed(9)
    for (B, cin, cout, H, W) in [(2, 64, 64, 9, 7), (2, 64, 128, 21, 18), (1, 8, 8, 5, 5), (2, 1, 64, 12, 10),
                                 (3, 128, 256, 30, 17), (2, 256, 96, 12, 9)]:
        x = torch.randn(B, cin, H, W, generator=g).cuda()
        w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.1).cuda()
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        yr = F.conv2d(xr, wr, padding=1)
        probe = torch.randn(yr.shape, generator=g).cuda()
        (yr * probe).sum().backward()
        xo, wo = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        z = _Conv3x3Function.apply(F.pad(xo.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous(), wo)
        yo = z[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
        (yo * probe).sum().backward()
        assert relmax(yo.detach(), yr.detach()) <= TOL
        assert relmax(xo.grad, xr.grad) <= TOL, relmax(xo.grad, xr.grad)
        # tensor-core wgrad (TF32) where the channel counts allow it, FP32 library GEMMs otherwise
        assert relmax(wo.grad, wr.grad) <= (TOL if cin % 64 == 0 else 1e-5), relmax(wo.grad, wr.grad)


def test_cnn14_backward_small():
    """Whole-network gradient against the FLOAT64 evaluation of the reference modules (the oracle principle; the
    reference's own TF32 cuDNN path is 0.9996 away from float64 in cosine on the same case,
    tests/tools/cnn14_grad_probe.py).  Ours is bitwise reproducible within and across processes
    (tests/tools/cnn14_repro.py, cnn14_checksum.py, cnn14_poison.py).  The weights come from the seeded global generator
    (conftest.py).  This torch build seeds itself randomly per process, and about one weight draw in twelve is
    pathological for this quantity: with seed 20261017 even the reference's float32 CPU evaluation has cosine 0.75 to
    its own float64 evaluation in the first layer (a ReLU / max decision sits on a tie), and so did ours (0.74); for
    seeds 0..3 float32 agrees with float64 to 1e-10 or better."""
    from diffmst_b200 import Cnn14
    g = torch.Generator().manual_seed(3)
    # running-statistics BatchNorm: the last blocks see 1x2 / 1x1 maps, where batch statistics over two items are
    # ill-conditioned (batch-statistics BatchNorm is covered per block in test_conv_block_backward)
    ref = OracleCnn14(num_classes=32).cuda().double().eval()
    ours = Cnn14(num_classes=32).cuda().eval()
    ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    x = (torch.rand(2, 1, 1024, 128, generator=g) ** 2).cuda()   # smallest input that survives the six poolings
    want, got = ref(x.double()), ours(x)
    assert relmax(got, want) <= 2e-2
    want.square().mean().backward()
    got.square().mean().backward()
    for (n, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert po.grad is not None and torch.isfinite(po.grad).all(), n
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))
    # The reference's own numerics (float32 modules on cuDNN with TF32 convolutions, torch's default) on the same
    # weights and input: for weight draws whose deep 1x1 / 2x2 maps leave little gradient signal, the first block's
    # gradient - after twelve TF32 layers - is rounding-noise dominated for ANY TF32 evaluation, so the bound is the
    # reference's own distance from float64 where that is larger than the nominal one (SURVEY.md section 8c protocol).
    ref32 = OracleCnn14(num_classes=32).cuda().eval()
    ref32.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    torch.backends.cudnn.allow_tf32 = True
    ref32(x).square().mean().backward()
    torch.backends.cudnn.allow_tf32 = False
    report = {}
    for name in ("conv_block1.conv1", "conv_block3.conv1", "conv_block6.conv2"):
        w64 = dict(ref.named_parameters())[name + ".weight"].grad
        c_ours = cos(dict(ours.named_parameters())[name + ".weight"].grad, w64)
        c_ref = cos(dict(ref32.named_parameters())[name + ".weight"].grad, w64)
        report[name] = (c_ours, c_ref)
    print("cosine to float64 (ours, reference TF32):", report)
    for name, (c_ours, c_ref) in report.items():
        assert 1.0 - c_ours <= max(2e-3, 2.0 * (1.0 - c_ref)), (name, report, relmax(got, want))
    # and the step is reproducible bit for bit
    g1 = ours.conv_block1.conv1.weight.grad.clone()
    for p in ours.parameters():
        p.grad = None
    ours(x).square().mean().backward()
    assert torch.equal(g1, ours.conv_block1.conv1.weight.grad)


def test_cnn14_forward_full_stack():
    """The whole conv trunk at the encoder's real input size (mst/modules.py:786-806:
    1025 bins x 257 frames for a 131072-sample excerpt)."""
    from diffmst_b200 import Cnn14
    g = torch.Generator().manual_seed(11)
    ref = OracleCnn14(num_classes=512)
    for i in range(6):
        _randomise_bn(getattr(ref, f"conv_block{i + 1}"), g)
    ref = ref.cuda().eval()
    ours = Cnn14(num_classes=512).cuda().eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = (torch.rand(2, 1, 1025, 257, generator=g) ** 3).cuda()   # spectrogram-like, non-negative
    with torch.no_grad():
        want = ref(x)
        got = ours(x)
    assert got.shape == (2, 512)
    assert relmax(got, want) <= 1e-2, relmax(got, want)   # 12 TF32 layers deep


def test_spectrogram_encoder_matches_reference_composition():
    """mst/modules.py:740-806: STFT 2048/512 -> (|X|+1e-8)^0.3 -> Cnn14; state dict names as upstream
    (window, model.conv_block*.{conv,bn}*, model.fc)."""
    from diffmst_b200 import SpectrogramEncoder
    g = torch.Generator().manual_seed(21)
    enc = SpectrogramEncoder(embed_dim=64, n_inputs=1).cuda().eval()
    ref = OracleCnn14(num_classes=64).cuda().eval()
    ref.load_state_dict(enc.model.state_dict(), strict=True)
    assert set(k.split(".")[0] for k in enc.state_dict()) == {"window", "model"}
    x = (torch.randn(2, 1, 131072, generator=g) * 0.1).cuda()
    with torch.no_grad():
        got = enc(x)
        X = torch.stft(x.view(-1, 131072), n_fft=2048, hop_length=512, window=torch.hann_window(2048).cuda(),
                       return_complex=True).view(2, 1, 1025, -1)
        want = ref(torch.pow(X.abs() + 1e-8, 0.3))
    assert got.shape == (2, 64)
    assert relmax(got, want) <= 1e-2, relmax(got, want)


def test_spectrogram_frontend_matches_torch_stft_composition():
    """dmst_spectrogram_frontend against the reference's lines (mst/modules.py:787-800): torch.stft(n_fft, hop,
    Hann) -> (|X| + 1e-8)^0.3, as the zero-bordered NHWC tensor conv_block1 reads.  float32 FFTs: 1e-4 relative
    to the largest value; the border is exactly zero; odd lengths, several channels and ragged tiles covered."""
    from diffmst_b200 import SpectrogramEncoder
    g = torch.Generator().manual_seed(5)
    for bs, chs, T, n_fft, hop in ((2, 1, 131072, 2048, 512), (1, 2, 50001, 2048, 512), (3, 1, 9000, 512, 128)):
        enc = SpectrogramEncoder(embed_dim=8, n_inputs=chs, n_fft=n_fft, hop_length=hop).cuda().eval()
        x = (torch.randn(bs, chs, T, generator=g) * 0.1).cuda()
        got = enc._frontend(x)
        X = torch.stft(x.view(-1, T), n_fft=n_fft, hop_length=hop, window=torch.hann_window(n_fft).cuda(),
                       return_complex=True).view(bs, chs, n_fft // 2 + 1, -1)
        want = torch.pow(X.abs() + 1e-8, 0.3)
        assert got.shape == (bs, want.shape[2] + 2, want.shape[3] + 2, chs)
        inner = got[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
        assert relmax(inner, want) <= 1e-4, relmax(inner, want)
        assert float(got[:, 0].abs().max()) == 0.0 and float(got[:, -1].abs().max()) == 0.0
        assert float(got[:, :, 0].abs().max()) == 0.0 and float(got[:, :, -1].abs().max()) == 0.0
    # a waveform that needs a gradient takes the differentiable composition; too-short input is torch.stft's error
    enc = SpectrogramEncoder(embed_dim=8).cuda().eval()
    xg = (torch.randn(1, 1, 65536, generator=g) * 0.1).cuda().requires_grad_(True)
    enc(xg).sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()
    with pytest.raises(RuntimeError, match="Padding size"):
        enc._frontend(torch.zeros(1, 1, 1000).cuda())


@pytest.mark.parametrize("shape", [(1, 6, 5, 32, 32), (2, 17, 9, 64, 64), (1, 33, 20, 32, 128), (2, 8, 8, 128, 256),
                                   (1, 6, 5, 1, 64), (2, 33, 17, 1, 64), (3, 40, 24, 96, 160)])
def test_wgrad_tensor_core_matches_float64(shape):
    """dmst_conv3x3_wgrad (tcgen05, MN-major TF32 operands; streaming kernel for the 1-channel first layer) against
    torch's float64 weight gradient of the same convolution.  TF32 products with FP32 accumulation: 3e-3 of the
    largest entry (the tolerance of the forward convolution); the first layer runs in FP32: 1e-5.  Also: the result
    is deterministic and the library-GEMM arm used for uncovered channel counts agrees."""
    from diffmst_b200 import conv
    B, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(B * 1000 + H + Cin)
    x = torch.randn(B, Cin, H, W, generator=g).cuda()
    gy = torch.randn(B, Cout, H, W, generator=g).cuda()
    x_pad = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
    gz = F.pad(gy.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()

    def wgrad(tc):
        conv._TC_WGRAD = tc
        try:
            w = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
            conv._Conv3x3Function.apply(x_pad.clone(), w, True).backward(gz)
            return w.grad
        finally:
            conv._TC_WGRAD = True

    want = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, 3, 3), gy.double(), padding=1)
    got = wgrad(True)
    assert relmax(got, want) <= (1e-5 if Cin == 1 else TOL), relmax(got, want)
    assert torch.equal(got, wgrad(True))
    assert relmax(wgrad(False), want) <= TOL


def test_encoder_matches_reference_golden(golden):
    """The product against the golden vectors written by the REFERENCE's own classes (tests/golden/make_golden_panns.py:
    mst.panns.Cnn14 in float64 with the generator-free parameter fill, and the spectrogram lines of
    mst.modules.SpectrogramEncoder.forward).  The case is well conditioned - the reference algorithm with its
    operands rounded to TF32 on the CPU is 8.6e-5 from the golden output, cosine 0.99960 / 0.99998 for the first-layer
    / head gradients - so the bounds below are the convolution tolerance and about 3x those gradient distances.  Every
    producer of a tensor-core operand rounds it to nearest TF32 (csrc/conv_tc.cuh: tf32_rn); the tensor core itself
    truncates, which on this structured case put the first-layer gradient at cosine 0.982."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from panns_fill import fill_state
    from diffmst_b200 import Cnn14, SpectrogramEncoder
    d = golden("panns")
    net = Cnn14(num_classes=6).cuda().eval()
    fill_state(net)
    g = torch.Generator().manual_seed(42)
    x = (torch.rand(1, 1, 1024, 128, generator=g) ** 2).cuda()
    with torch.no_grad():
        out_eval = net(x)                      # inference path (BatchNorm folded into the convolution epilogue)
    out = net(x)                               # differentiable path
    out.square().mean().backward()
    want = torch.from_numpy(d["cnn14_out"]).cuda()
    assert relmax(out_eval, want) <= TOL, relmax(out_eval, want)
    assert relmax(out, want) <= TOL, relmax(out, want)
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))
    c_first = cos(net.conv_block1.conv1.weight.grad, torch.from_numpy(d["cnn14_g_first"]).cuda())
    c_fc = cos(net.fc.weight.grad, torch.from_numpy(d["cnn14_g_fc"]).cuda())
    # measured with operands rounded to nearest TF32 (tests/tools/cnn14_tf32_diag.py, profiles/cnn14_tf32_diag_r2.txt):
    # 1 - cosine = 3.9e-4 / 2.4e-5, cuDNN's TF32 convolutions on the same case 3.0e-4 / 2.3e-5; with the operands
    # truncated (what the tensor core does to unrounded float32 bits) it was 1.8e-2 / 1.6e-4
    assert c_first >= 0.999 and c_fc >= 0.9999, (c_first, c_fc)
    enc = SpectrogramEncoder(embed_dim=8).cuda().eval()
    S = enc._frontend(torch.from_numpy(d["spec_wave"]).cuda())[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
    assert S.shape == d["spec_out"].shape
    assert relmax(S, torch.from_numpy(d["spec_out"]).cuda()) <= 1e-4
