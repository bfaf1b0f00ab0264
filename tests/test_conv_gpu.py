"""GPU parity of the tensor-core ConvBlock / Cnn14 forward against a plain PyTorch float32
reference of the same op (TF32 disabled in the reference).  Tolerance: TF32 operands carry a
10-bit mantissa (2^-11 relative rounding per operand), accumulation is FP32; we require 3e-3 of the
output's max magnitude, the same class as cuDNN's default TF32 convolutions."""
import numpy as np
import pytest
import torch

from oracle.panns import OracleCnn14, OracleConvBlock

pytestmark = pytest.mark.gpu
TOL = 3e-3


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


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
    with pytest.raises(NotImplementedError):
        ours(x, (2, 2))                      # grad-enabled call: backward is not built yet
    with pytest.raises(RuntimeError, match="no CPU path"):
        with torch.no_grad():
            ours(x.cpu(), (2, 2))


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
