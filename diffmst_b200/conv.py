"""Cnn14 convolution blocks on the B200 tensor cores (forward).

``ConvBlock`` and ``Cnn14`` mirror mst/panns.py:27-85 and :126-209: same constructor arguments,
same sub-module / parameter names (``conv1.weight``, ``bn1.running_mean`` ..., ``fc.weight``), so a
reference checkpoint loads with ``load_state_dict``.  The 3x3 convolutions run as TF32 implicit
GEMMs on tcgen05 with TMA-fed shared-memory tiles and TMEM accumulators (csrc/conv_tc.cuh);
BatchNorm is folded into the epilogue in eval mode and applied from batch statistics in training
mode; activations stay in zero-bordered NHWC between the layers of a block (and between blocks
inside ``Cnn14``).

Round-1 limitation (stated, not hidden): forward only.  Calling these modules with autograd
enabled on tensors that require grad raises; the backward kernels (dgrad / wgrad) are the next row.
"""
import ctypes
from typing import List

import torch
import torch.nn as nn

from . import _lib
from .console import _ptr, _require_cuda


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _to_padded_nhwc(x):
    lib = _lib.lib()
    B, C, H, W = x.shape
    y = torch.empty(B, H + 2, W + 2, C, dtype=torch.float32, device=x.device)
    _lib.check(lib.dmst_conv_nchw_to_padded_nhwc(_ptr(x.contiguous()), _ptr(y), B, C, H, W, _stream(x.device)),
               "dmst_conv_nchw_to_padded_nhwc")
    return y


def _repack(w):
    lib = _lib.lib()
    Cout, Cin = w.shape[0], w.shape[1]
    w9 = torch.empty(9, Cout, Cin, dtype=torch.float32, device=w.device)
    _lib.check(lib.dmst_conv_repack_weights(_ptr(w.detach().contiguous()), _ptr(w9), Cout, Cin, _stream(w.device)),
               "dmst_conv_repack_weights")
    return w9


def _conv_bn_relu(x_pad, conv: nn.Conv2d, bn, training: bool):
    """x_pad (B, H+2, W+2, Cin) -> relu(bn(conv(x))) as (B, H+2, W+2, Cout)."""
    lib = _lib.lib()
    B, Hp, Wp, Cin = x_pad.shape
    H, W = Hp - 2, Wp - 2
    Cout = conv.weight.shape[0]
    dev = x_pad.device
    w9 = _repack(conv.weight)
    y = torch.empty(B, Hp, Wp, Cout, dtype=torch.float32, device=dev)
    is_bn = isinstance(bn, nn.BatchNorm2d)
    use_batch_stats = is_bn and (training or bn.running_mean is None)
    if is_bn and not use_batch_stats:
        scale = (bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias.detach() - bn.running_mean * scale).float().contiguous()
        _lib.check(lib.dmst_conv3x3_forward(_ptr(x_pad), _ptr(w9), _ptr(scale), _ptr(shift), _ptr(y), B, H, W, Cin,
                                            Cout, 1, _stream(dev)), "dmst_conv3x3_forward")
        return y
    if not is_bn:
        _lib.check(lib.dmst_conv3x3_forward(_ptr(x_pad), _ptr(w9), None, None, _ptr(y), B, H, W, Cin, Cout, 1,
                                            _stream(dev)), "dmst_conv3x3_forward")
        return y
    # training-mode BatchNorm: raw conv -> batch statistics -> affine + ReLU in place
    _lib.check(lib.dmst_conv3x3_forward(_ptr(x_pad), _ptr(w9), None, None, _ptr(y), B, H, W, Cin, Cout, 0,
                                        _stream(dev)), "dmst_conv3x3_forward")
    nbytes = lib.dmst_conv_stats_workspace_bytes(B, H, W, Cout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    mean = torch.empty(Cout, dtype=torch.float32, device=dev)
    var = torch.empty(Cout, dtype=torch.float32, device=dev)
    _lib.check(lib.dmst_conv_channel_stats(_ptr(y), B, H, W, Cout, _ptr(mean), _ptr(var), _ptr(ws), nbytes,
                                           _stream(dev)), "dmst_conv_channel_stats")
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            n = B * H * W
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
            bn.running_var.mul_(1 - m).add_(var * (n / max(n - 1, 1)), alpha=m)
            bn.num_batches_tracked += 1
    scale = (bn.weight.detach() * torch.rsqrt(var + bn.eps)).contiguous()
    shift = (bn.bias.detach() - mean * scale).contiguous()
    _lib.check(lib.dmst_conv_affine_relu(_ptr(y), _ptr(scale), _ptr(shift), B, H, W, Cout, 1, _stream(dev)),
               "dmst_conv_affine_relu")
    return y


def _avgpool(x_pad, kh, kw, out_padded_nhwc):
    lib = _lib.lib()
    B, Hp, Wp, C = x_pad.shape
    H, W = Hp - 2, Wp - 2
    Ho, Wo = H // kh, W // kw
    if out_padded_nhwc:
        y = torch.zeros(B, Ho + 2, Wo + 2, C, dtype=torch.float32, device=x_pad.device)
    else:
        y = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=x_pad.device)
    _lib.check(lib.dmst_conv_avgpool(_ptr(x_pad), _ptr(y), B, C, H, W, kh, kw, 1 if out_padded_nhwc else 0,
                                     _stream(x_pad.device)), "dmst_conv_avgpool")
    return y


def _check_no_grad(x, module):
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        raise NotImplementedError(
            "diffmst_b200 ConvBlock/Cnn14: forward only in this round (tensor-core dgrad/wgrad are not built yet); "
            "call under torch.no_grad() or freeze the encoder")


def init_layer(layer):
    nn.init.xavier_uniform_(layer.weight)
    if hasattr(layer, "bias") and layer.bias is not None:
        layer.bias.data.fill_(0.0)


def init_bn(bn):
    bn.bias.data.fill_(0.0)
    bn.weight.data.fill_(1.0)


class ConvBlock(nn.Module):
    """mst/panns.py:27-85 (pool_type 'avg' only: the one Cnn14 uses)."""

    def __init__(self, in_channels, out_channels, use_batchnorm: bool = True, pool_type: str = "avg"):
        super().__init__()
        if pool_type != "avg":
            raise NotImplementedError("only pool_type='avg' (the Cnn14 default) is implemented")
        self.use_batchnorm = use_batchnorm
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        self.bn2 = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        self.init_weight()

    def init_weight(self):
        init_layer(self.conv1)
        init_layer(self.conv2)
        if self.use_batchnorm:
            init_bn(self.bn1)
            init_bn(self.bn2)

    def forward_nhwc(self, x_pad, pool_size, out_padded_nhwc: bool):
        x = _conv_bn_relu(x_pad, self.conv1, self.bn1, self.training)
        x = _conv_bn_relu(x, self.conv2, self.bn2, self.training)
        return _avgpool(x, int(pool_size[0]), int(pool_size[1]), out_padded_nhwc)

    def forward(self, input: torch.Tensor, pool_size: List[int]):
        _require_cuda(input, "input")
        _check_no_grad(input, self)
        return self.forward_nhwc(_to_padded_nhwc(input), pool_size, out_padded_nhwc=False)


class Cnn14(nn.Module):
    """mst/panns.py:126-209: six ConvBlocks (pools (2,2),(4,4),(4,2),(4,2),(4,2),(2,2)), mean over
    bins, max+mean over frames, linear head."""

    POOLS = [(2, 2), (4, 4), (4, 2), (4, 2), (4, 2), (2, 2)]

    def __init__(self, num_classes: int, n_inputs: int = 1, use_batchnorm: bool = True):
        super().__init__()
        chans = [n_inputs, 64, 128, 256, 512, 1024, 2048]
        for i in range(6):
            setattr(self, f"conv_block{i + 1}", ConvBlock(chans[i], chans[i + 1], use_batchnorm=use_batchnorm))
        self.fc = nn.Linear(2048, num_classes, bias=True)
        init_layer(self.fc)

    def forward(self, x: torch.Tensor):
        _require_cuda(x, "x")
        _check_no_grad(x, self)
        h = _to_padded_nhwc(x)
        for i, pool in enumerate(self.POOLS):
            h = getattr(self, f"conv_block{i + 1}").forward_nhwc(h, pool, out_padded_nhwc=(i < 5))
        h = torch.mean(h, dim=2)            # mean across stft bins
        x1, _ = torch.max(h, dim=2)
        x2 = torch.mean(h, dim=2)
        return self.fc(x1 + x2)
