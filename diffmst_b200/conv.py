"""Cnn14 convolution blocks on the B200 tensor cores.

``ConvBlock`` and ``Cnn14`` mirror mst/panns.py:27-85 and :126-209: same constructor arguments,
same sub-module / parameter names (``conv1.weight``, ``bn1.running_mean`` ..., ``fc.weight``), so a
reference checkpoint loads with ``load_state_dict``.  The 3x3 convolutions run as TF32 implicit
GEMMs on tcgen05 with TMA-fed shared-memory tiles and TMEM accumulators (csrc/conv_tc.cuh);
BatchNorm is folded into the epilogue in eval mode and applied from batch statistics in training
mode; activations stay in zero-bordered NHWC between the layers of a block (and between blocks
inside ``Cnn14``).

Training: with autograd enabled the 3x3 convolution is a ``torch.autograd.Function`` whose forward
and input gradient (dgrad = the same shifted-GEMM kernel run on the output gradient with the taps
flipped and the channel roles swapped) run on tcgen05, and so does the weight gradient
(``dmst_conv3x3_wgrad``: per tap ``dz^T @ x_shifted`` over the flattened zero-bordered NHWC tensors
with MN-major TF32 operands, split over pixel chunks; split-K batched library GEMMs remain only
for channel counts that kernel does not cover, none of them in Cnn14); BatchNorm + ReLU and the
average pooling of the differentiable path are CUDA autograd Functions too (batch statistics,
normalisation, their backward and the pooling backward in csrc/conv_tc.cuh).
"""
import ctypes
import weakref
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

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


# measurement aid (tests/tools/cnn14_train_bench.py): False routes BatchNorm / ReLU / pooling of the differentiable path
# through PyTorch ops on NHWC views, as before the CUDA Functions existed
_CUDA_BN_POOL = True

_TC_WGRAD = True       # False: weight gradient through the split-K batched library GEMMs (comparison / fallback arm)

_REPACK_CACHE = {}   # id(parameter) -> (weak reference to it, data_ptr, version, [9][Cout][Cin] tensor)
RELU, ROUND_TF32 = 1, 2   # mask of the `relu` argument of the convolution / affine entry points (include/diffmst_b200.h)


def clear_repack_cache():
    """Forget the repacked inference weights (call after modifying a weight through ``.data`` in eval mode)."""
    _REPACK_CACHE.clear()


def _repack(w):
    """[Cout][Cin][3][3] -> [9][Cout][Cin] (rounded to TF32 for the tensor-core kernel).

    Cached only for inference (autograd off, no CUDA-graph capture in progress): an optimizer step, or
    ``weight.data.copy_()``, does not bump ``weight._version``, so a cache consulted while training - or a graph
    captured after an eager warm-up had filled it - would silently keep the old weights in forward while dgrad
    repacks fresh ones.  With autograd on, or while a stream is capturing, the (small) repack kernel always runs and is
    part of the captured graph."""
    cacheable = not torch.is_grad_enabled() and not torch.cuda.is_current_stream_capturing()
    key = id(w)
    if cacheable:
        hit = _REPACK_CACHE.get(key)
        if hit is not None and hit[0]() is w and hit[1] == w.data_ptr() and hit[2] == w._version and hit[3].device == w.device:
            return hit[3]
    lib = _lib.lib()
    Cout, Cin = w.shape[0], w.shape[1]
    w9 = torch.empty(9, Cout, Cin, dtype=torch.float32, device=w.device)
    _lib.check(lib.dmst_conv_repack_weights(_ptr(w.detach().contiguous()), _ptr(w9), Cout, Cin, _stream(w.device)),
               "dmst_conv_repack_weights")
    if cacheable:
        _REPACK_CACHE[key] = (weakref.ref(w, lambda _r, k=key: _REPACK_CACHE.pop(k, None)), w.data_ptr(), w._version, w9)
    else:
        _REPACK_CACHE.pop(key, None)
    return w9


def _round_tf32(x, bordered_shape=None):
    """Copy of x rounded to nearest TF32; bordered_shape = (B, Hp, Wp, C) also clears the one-pixel border."""
    lib = _lib.lib()
    x = x.contiguous()
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    if bordered_shape is None:
        B = H = W = C = 0
    else:
        B, Hp, Wp, C = bordered_shape
        H, W = Hp - 2, Wp - 2
    _lib.check(lib.dmst_conv_round_tf32(_ptr(x), _ptr(y), x.numel(), B, H, W, C, _stream(x.device)), "dmst_conv_round_tf32")
    return y


def _tc_operand(C):
    return C % 32 == 0


def _conv3x3(x_pad, w9, scale, shift, y, B, H, W, Cin, Cout, relu, what="dmst_conv3x3_forward"):
    """One tensor-core convolution call (split-K workspace allocated on demand for the small deep layers)."""
    lib = _lib.lib()
    nbytes = lib.dmst_conv3x3_workspace_bytes(B, H, W, Cin, Cout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_pad.device) if nbytes else None
    _lib.check(lib.dmst_conv3x3_forward_ws(_ptr(x_pad), _ptr(w9), _ptr(scale), _ptr(shift), _ptr(y), B, H, W, Cin, Cout,
                                           relu, _ptr(ws), nbytes, _stream(x_pad.device)), what)


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
    use_batch_stats = is_bn and (bn.training or bn.running_mean is None)   # nn.BatchNorm2d.forward's rule
    if is_bn and not use_batch_stats:
        scale = (bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias.detach() - bn.running_mean * scale).float().contiguous()
        _conv3x3(x_pad, w9, scale, shift, y, B, H, W, Cin, Cout, RELU | ROUND_TF32)
        return y
    if not is_bn:
        _conv3x3(x_pad, w9, None, None, y, B, H, W, Cin, Cout, RELU | ROUND_TF32)
        return y
    # training-mode BatchNorm: raw conv -> batch statistics -> affine + ReLU in place
    _conv3x3(x_pad, w9, None, None, y, B, H, W, Cin, Cout, 0)
    mean, var = _channel_stats(y)
    if bn.training:
        _update_running_stats(bn, mean, var, B * H * W)
    scale = (bn.weight.detach() * torch.rsqrt(var + bn.eps)).contiguous()
    shift = (bn.bias.detach() - mean * scale).contiguous()
    _lib.check(lib.dmst_conv_affine_relu(_ptr(y), _ptr(scale), _ptr(shift), B, H, W, Cout, RELU | ROUND_TF32, _stream(dev)),
               "dmst_conv_affine_relu")
    return y


class _Conv3x3Function(torch.autograd.Function):
    """z = conv3x3(x) on zero-bordered NHWC tensors (no bias, stride 1, padding 1; mst/panns.py:33-47)."""

    @staticmethod
    def forward(ctx, x_pad, weight, clean_border=False, x_rounded=False):
        # clean_border: the gradient that will arrive for z is known to be zero on the border and rounded to TF32 (it
        # comes from _BnReluFunction), so backward need not clear / round it.
        # x_rounded: x_pad was produced by one of this library's kernels, which round what they hand to a convolution
        # to TF32; any other input is rounded here (the tensor core would truncate it, see csrc/conv_tc.cuh).
        ctx.clean_border = bool(clean_border)
        lib = _lib.lib()
        B, Hp, Wp, Cin = x_pad.shape
        Cout = weight.shape[0]
        x_pad = x_pad.contiguous()
        if not x_rounded and _tc_operand(Cin):
            x_pad = _round_tf32(x_pad)
        w9 = _repack(weight)
        z = torch.empty(B, Hp, Wp, Cout, dtype=torch.float32, device=x_pad.device)
        _conv3x3(x_pad, w9, None, None, z, B, Hp - 2, Wp - 2, Cin, Cout, 0)
        ctx.save_for_backward(x_pad, w9, weight)
        return z

    @staticmethod
    def backward(ctx, gz):
        lib = _lib.lib()
        x_pad, w9, weight = ctx.saved_tensors
        B, Hp, Wp, Cin = x_pad.shape
        Cout = w9.shape[1]
        gz = gz.contiguous()
        if not ctx.clean_border:
            # the border of the output is padding: whatever gradient arrives there does not exist upstream.  Out of
            # place (the incoming tensor belongs to autograd and may be shared), rounded to TF32 on the way.
            if _tc_operand(Cout) and _tc_operand(Cin):   # a tensor-core dgrad / wgrad will read it
                gz = _round_tf32(gz, (B, Hp, Wp, Cout))
            else:
                gz = gz.clone()
                gz[:, 0, :, :] = 0; gz[:, -1, :, :] = 0; gz[:, :, 0, :] = 0; gz[:, :, -1, :] = 0
        gx = gw = None
        if ctx.needs_input_grad[0]:
            # dgrad: correlation with the flipped taps, channel roles swapped -> the same kernel
            w9t = torch.empty(9, Cin, Cout, dtype=torch.float32, device=gz.device)   # [flipped tap][Cin][Cout]
            _lib.check(lib.dmst_conv_repack_weights_dgrad(_ptr(weight.detach().contiguous()), _ptr(w9t), Cout, Cin,
                                                          _stream(gz.device)), "dmst_conv_repack_weights_dgrad")
            gx = torch.empty_like(x_pad)
            _conv3x3(gz, w9t, None, None, gx, B, Hp - 2, Wp - 2, Cout, Cin, 0, "dmst_conv3x3_forward (dgrad)")
        if ctx.needs_input_grad[1]:
            # wgrad on the tensor cores: dmst_conv3x3_wgrad (tcgen05, MN-major operands straight from the NHWC tensors)
            nbytes = lib.dmst_conv3x3_wgrad_workspace_bytes(B, Hp - 2, Wp - 2, Cin, Cout) if _TC_WGRAD else 0
            if nbytes:
                ws = torch.empty(nbytes, dtype=torch.uint8, device=gz.device)
                gw = torch.empty(Cout, Cin, 3, 3, dtype=torch.float32, device=gz.device)
                _lib.check(lib.dmst_conv3x3_wgrad(_ptr(x_pad), _ptr(gz), _ptr(gw), B, Hp - 2, Wp - 2, Cin, Cout, _ptr(ws), nbytes,
                                                  _stream(gz.device)), "dmst_conv3x3_wgrad")
                return gx, gw, None, None
            # channel counts the kernels do not cover (Cin neither 1 nor a multiple of 32): library GEMMs.
            # dW[tap] = dz^T @ x shifted by the tap (a constant row offset in the flattened layout; dz is zero on the
            # border, so rows that would cross an image edge contribute nothing); K = all pixels is split into S chunks
            # run as one batched GEMM per tap.  TF32 follows torch.backends.cudnn.allow_tf32 (the reference's switch).
            P = B * Hp * Wp
            gzf, xf = gz.view(P, Cout), x_pad.view(P, Cin)
            first = Wp + 1                                   # rows before it / after P - first are border rows: dz = 0
            Pc = P - 2 * first
            tiles = -(-Cout // 128) * -(-Cin // 128)
            S = max(1, min(256, -(-296 // tiles), Pc // 512))
            Kc = Pc // S
            main = S * Kc
            tf32 = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = bool(torch.backends.cudnn.allow_tf32)
            try:
                a = gzf[first:first + main].view(S, Kc, Cout).transpose(1, 2)              # (S, Cout, Kc)
                part = torch.empty(9, S, Cout, Cin, dtype=torch.float32, device=gz.device)
                for t in range(9):
                    off = (t // 3 - 1) * Wp + (t % 3 - 1)
                    torch.bmm(a, xf[first + off:first + off + main].view(S, Kc, Cin), out=part[t])
                g9 = part.sum(dim=1)
                if main < Pc:                                                               # the last Pc - main < S rows
                    ar = gzf[first + main:first + Pc].t()
                    for t in range(9):
                        off = (t // 3 - 1) * Wp + (t % 3 - 1)
                        g9[t].addmm_(ar, xf[first + main + off:first + Pc + off])
            finally:
                torch.backends.cuda.matmul.allow_tf32 = tf32
            gw = g9.permute(1, 2, 0).reshape(Cout, Cin, 3, 3)
        return gx, gw, None, None


def _needs_grad(x, *modules):
    if not torch.is_grad_enabled():
        return False
    return x.requires_grad or any(p.requires_grad for m in modules if isinstance(m, nn.Module) for p in m.parameters())


def _channel_stats(z_pad):
    """Per-channel batch mean and biased variance of a zero-bordered NHWC convolution output."""
    lib = _lib.lib()
    B, Hp, Wp, C = z_pad.shape
    dev = z_pad.device
    if C % 4:   # odd channel counts (never in Cnn14): the vectorised reduction does not apply
        zi = z_pad[:, 1:-1, 1:-1, :]
        return zi.mean(dim=(0, 1, 2)), zi.var(dim=(0, 1, 2), unbiased=False)
    nbytes = lib.dmst_conv_stats_workspace_bytes(B, Hp - 2, Wp - 2, C)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    mean = torch.empty(C, dtype=torch.float32, device=dev)
    var = torch.empty(C, dtype=torch.float32, device=dev)
    _lib.check(lib.dmst_conv_channel_stats(_ptr(z_pad), B, Hp - 2, Wp - 2, C, _ptr(mean), _ptr(var), _ptr(ws), nbytes,
                                           _stream(dev)), "dmst_conv_channel_stats")
    return mean, var


def _update_running_stats(bn, mean, var, n):
    """nn.BatchNorm2d's bookkeeping in training mode (unbiased variance into the running estimate)."""
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
            bn.running_var.mul_(1 - m).add_(var * (n / max(n - 1, 1)), alpha=m)


class _BnReluFunction(torch.autograd.Function):
    """y = relu(BatchNorm(z)) on zero-bordered NHWC (mst/panns.py:79-80), forward and backward in CUDA.  `mean` / `var`
    are the statistics the normalisation uses: the batch's (batch_stats=True; the backward then carries the mean and
    variance terms) or the running estimates."""

    @staticmethod
    def forward(ctx, z_pad, gamma, beta, mean, var, eps, batch_stats):
        lib = _lib.lib()
        B, Hp, Wp, C = z_pad.shape
        z_pad = z_pad.contiguous()
        rstd = torch.rsqrt(var.detach().float() + eps)
        scale = (gamma.detach().float() * rstd).contiguous()
        shift = (beta.detach().float() - mean.detach().float() * scale).contiguous()
        y = torch.empty_like(z_pad)
        _lib.check(lib.dmst_conv_affine_relu_to(_ptr(z_pad), _ptr(y), _ptr(scale), _ptr(shift), B, Hp - 2, Wp - 2, C,
                                                _stream(z_pad.device)), "dmst_conv_affine_relu_to")
        ctx.save_for_backward(z_pad, scale, shift, mean.detach().float().contiguous(), rstd.contiguous())
        ctx.batch_stats = bool(batch_stats)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.lib()
        z_pad, scale, shift, mean, rstd = ctx.saved_tensors
        B, Hp, Wp, C = z_pad.shape
        dev = z_pad.device
        gy = gy.contiguous()
        dz = torch.empty_like(z_pad)
        dgamma = torch.empty(C, dtype=torch.float32, device=dev)
        dbeta = torch.empty(C, dtype=torch.float32, device=dev)
        nbytes = lib.dmst_conv_stats_workspace_bytes(B, Hp - 2, Wp - 2, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.dmst_conv_bn_relu_backward(_ptr(z_pad), _ptr(gy), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd),
                                                  1 if ctx.batch_stats else 0, B, Hp - 2, Wp - 2, C, _ptr(dz), _ptr(dgamma),
                                                  _ptr(dbeta), _ptr(ws), nbytes, _stream(dev)), "dmst_conv_bn_relu_backward")
        return (dz, dgamma if ctx.needs_input_grad[1] else None, dbeta if ctx.needs_input_grad[2] else None,
                None, None, None, None)


class _BnReluPoolFunction(torch.autograd.Function):
    """avg_pool2d(relu(BatchNorm(z))) as one unit (mst/panns.py:80-85): the BatchNorm+ReLU output and its gradient are
    never materialised; backward gathers the pooled gradient on the fly.  Pool sizes are powers of two."""

    @staticmethod
    def forward(ctx, z_pad, gamma, beta, mean, var, eps, batch_stats, kh, kw, out_padded_nhwc):
        lib = _lib.lib()
        B, Hp, Wp, C = z_pad.shape
        H, W = Hp - 2, Wp - 2
        z_pad = z_pad.contiguous()
        rstd = torch.rsqrt(var.detach().float() + eps)
        scale = (gamma.detach().float() * rstd).contiguous()
        shift = (beta.detach().float() - mean.detach().float() * scale).contiguous()
        Ho, Wo = H // kh, W // kw
        if out_padded_nhwc:
            y = torch.zeros(B, Ho + 2, Wo + 2, C, dtype=torch.float32, device=z_pad.device)
        else:
            y = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=z_pad.device)
        _lib.check(lib.dmst_conv_bn_relu_avgpool(_ptr(z_pad), _ptr(scale), _ptr(shift), _ptr(y), B, C, H, W, kh, kw,
                                                 1 if out_padded_nhwc else 0, _stream(z_pad.device)), "dmst_conv_bn_relu_avgpool")
        ctx.save_for_backward(z_pad, scale, shift, mean.detach().float().contiguous(), rstd.contiguous())
        ctx.cfg = (bool(batch_stats), kh, kw, bool(out_padded_nhwc))
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.lib()
        z_pad, scale, shift, mean, rstd = ctx.saved_tensors
        batch_stats, kh, kw, padded = ctx.cfg
        B, Hp, Wp, C = z_pad.shape
        dev = z_pad.device
        gy = gy.contiguous()
        dz = torch.empty_like(z_pad)
        dgamma = torch.empty(C, dtype=torch.float32, device=dev)
        dbeta = torch.empty(C, dtype=torch.float32, device=dev)
        nbytes = lib.dmst_conv_stats_workspace_bytes(B, Hp - 2, Wp - 2, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.dmst_conv_bn_relu_avgpool_backward(
            _ptr(z_pad), _ptr(gy), kh, kw, 1 if padded else 0, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd),
            1 if batch_stats else 0, B, Hp - 2, Wp - 2, C, _ptr(dz), _ptr(dgamma), _ptr(dbeta), _ptr(ws), nbytes, _stream(dev)),
            "dmst_conv_bn_relu_avgpool_backward")
        return (dz, dgamma if ctx.needs_input_grad[1] else None, dbeta if ctx.needs_input_grad[2] else None,
                None, None, None, None, None, None, None)


class _AvgPoolFunction(torch.autograd.Function):
    """F.avg_pool2d(x, (kh, kw)) of a zero-bordered NHWC tensor -> NCHW or zero-bordered NHWC (mst/panns.py:81-85)."""

    @staticmethod
    def forward(ctx, x_pad, kh, kw, out_padded_nhwc):
        ctx.cfg = (tuple(x_pad.shape), kh, kw, bool(out_padded_nhwc))
        return _avgpool_forward(x_pad.contiguous(), kh, kw, out_padded_nhwc)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.lib()
        (B, Hp, Wp, C), kh, kw, padded = ctx.cfg
        gy = gy.contiguous()
        gx = torch.empty(B, Hp, Wp, C, dtype=torch.float32, device=gy.device)
        _lib.check(lib.dmst_conv_avgpool_backward(_ptr(gy), _ptr(gx), B, C, Hp - 2, Wp - 2, kh, kw, 1 if padded else 0,
                                                  _stream(gy.device)), "dmst_conv_avgpool_backward")
        return gx, None, None, None


def _conv_bn_relu_autograd(x_pad, conv: nn.Conv2d, bn, training: bool, pool=None, x_rounded=False):
    """Differentiable twin of _conv_bn_relu: tensor-core conv Function, then BatchNorm + ReLU as one CUDA Function
    (statistics, normalisation and their backward in csrc/conv_tc.cuh).  pool = (kh, kw, out_padded_nhwc) fuses the
    block's average pooling into that Function when the pool sizes are powers of two."""
    C = conv.weight.shape[0]
    cuda_bn = isinstance(bn, nn.BatchNorm2d) and C % 4 == 0 and _CUDA_BN_POOL
    z = _Conv3x3Function.apply(x_pad, conv.weight, cuda_bn, x_rounded)
    if cuda_bn:
        batch_stats = bn.training or bn.running_mean is None   # nn.BatchNorm2d.forward's rule
        if batch_stats:
            mean, var = _channel_stats(z.detach())
            if bn.training:
                _update_running_stats(bn, mean, var, z.shape[0] * (z.shape[1] - 2) * (z.shape[2] - 2))
        else:
            mean, var = bn.running_mean, bn.running_var
        gamma = bn.weight if bn.weight is not None else torch.ones(C, device=z.device)
        beta = bn.bias if bn.bias is not None else torch.zeros(C, device=z.device)
        if pool is not None:
            kh, kw, out_padded = pool
            if kh & (kh - 1) == 0 and kw & (kw - 1) == 0:
                return _BnReluPoolFunction.apply(z, gamma, beta, mean, var, bn.eps, batch_stats, kh, kw, out_padded)
            return _avgpool(_BnReluFunction.apply(z, gamma, beta, mean, var, bn.eps, batch_stats), kh, kw, out_padded)
        return _BnReluFunction.apply(z, gamma, beta, mean, var, bn.eps, batch_stats)
    # no BatchNorm (use_batchnorm=False) or an odd channel count: PyTorch ops on NHWC views
    zi = z[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)                     # NCHW view of the interior (channels-last memory)
    if isinstance(bn, nn.BatchNorm2d):
        zi = bn(zi)
    y = F.pad(F.relu(zi).permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))    # back to zero-bordered NHWC
    return y if pool is None else _avgpool(y, *pool)


def _avgpool(x_pad, kh, kw, out_padded_nhwc):
    if torch.is_grad_enabled() and x_pad.requires_grad:
        if not _CUDA_BN_POOL or x_pad.shape[-1] % 4:
            y = F.avg_pool2d(x_pad[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2), kernel_size=(kh, kw))
            return F.pad(y.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)) if out_padded_nhwc else y.contiguous()
        return _AvgPoolFunction.apply(x_pad, kh, kw, out_padded_nhwc)
    return _avgpool_forward(x_pad, kh, kw, out_padded_nhwc)


def _avgpool_forward(x_pad, kh, kw, out_padded_nhwc):
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


def _to_padded_nhwc_any(x, grad: bool):
    if grad:
        return F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
    return _to_padded_nhwc(x)


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

    def forward_nhwc(self, x_pad, pool_size, out_padded_nhwc: bool, x_rounded: bool = False):
        """x_rounded: x_pad comes from one of this library's kernels (already rounded to TF32 where it matters)."""
        kh, kw = int(pool_size[0]), int(pool_size[1])
        if _needs_grad(x_pad, self):
            cuda_bn1 = isinstance(self.bn1, nn.BatchNorm2d) and self.conv1.weight.shape[0] % 4 == 0 and _CUDA_BN_POOL
            x = _conv_bn_relu_autograd(x_pad, self.conv1, self.bn1, self.training, x_rounded=x_rounded)
            return _conv_bn_relu_autograd(x, self.conv2, self.bn2, self.training, pool=(kh, kw, out_padded_nhwc),
                                          x_rounded=cuda_bn1)
        x = _conv_bn_relu(x_pad, self.conv1, self.bn1, self.training)
        x = _conv_bn_relu(x, self.conv2, self.bn2, self.training)
        return _avgpool(x, kh, kw, out_padded_nhwc)

    def forward(self, input: torch.Tensor, pool_size: List[int]):
        _require_cuda(input, "input")
        return self.forward_nhwc(_to_padded_nhwc_any(input, _needs_grad(input, self)), pool_size, out_padded_nhwc=False)


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
        return self.forward_padded_nhwc(_to_padded_nhwc_any(x, _needs_grad(x, self)))

    def forward_padded_nhwc(self, h: torch.Tensor):
        """Same as forward for an input already in the zero-bordered NHWC layout (B, H+2, W+2, n_inputs)."""
        for i, pool in enumerate(self.POOLS):
            # blocks 2..6 read the previous block's pooled output (rounded to TF32 by the pooling kernels whenever
            # _CUDA_BN_POOL units produced it; the fallback compositions are rounded inside the conv Function)
            h = getattr(self, f"conv_block{i + 1}").forward_nhwc(h, pool, out_padded_nhwc=(i < 5),
                                                                  x_rounded=(i > 0 and _CUDA_BN_POOL))
        h = torch.mean(h, dim=2)            # mean across stft bins
        x1, _ = torch.max(h, dim=2)
        x2 = torch.mean(h, dim=2)
        return self.fc(x1 + x2)


class SpectrogramEncoder(nn.Module):
    """mst/modules.py:740-806: waveform -> STFT (2048/512, Hann) -> (|X| + 1e-8)^0.3 -> Cnn14 -> embedding.

    Same constructor arguments, buffer (``window``) and sub-module names (``model``, ``bn``) as the
    reference, so its checkpoints load with ``load_state_dict``.  The convolution trunk is the
    tensor-core ``Cnn14`` above.  The spectrogram front-end is
    ``dmst_spectrogram_frontend``: 128-bit framing, one batched cuFFT R2C and one kernel that takes the
    magnitude, compresses it and writes the zero-bordered NHWC tensor the first convolution reads
    (three launches instead of pad / stft / abs / add / pow / layout conversion).  Only a waveform that
    itself needs a gradient, or ``input_batchnorm=True``, goes through the PyTorch composition."""

    def __init__(self, embed_dim: int = 128, n_inputs: int = 1, n_fft: int = 2048, hop_length: int = 512,
                 input_batchnorm: bool = False, encoder_batchnorm: bool = True) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.n_inputs = n_inputs
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.input_batchnorm = input_batchnorm
        self.register_buffer("window", torch.hann_window(window_length=int(n_fft)))
        self.model = Cnn14(n_inputs=n_inputs, num_classes=embed_dim, use_batchnorm=encoder_batchnorm)
        self.bn = nn.BatchNorm2d(3) if input_batchnorm else nn.Identity()

    def _frontend(self, x: torch.Tensor) -> torch.Tensor:
        """(bs, chs, T) waveform -> (bs, n_fft/2+1 + 2, frames + 2, chs) compressed magnitudes, zero border."""
        lib = _lib.lib()
        _require_cuda(x, "x")
        bs, chs, T = x.shape
        if T <= self.n_fft // 2:
            raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, "
                               f"but got: padding ({self.n_fft // 2}, {self.n_fft // 2}) at dimension 2 of input "
                               f"{[1, bs * chs, T]}")  # torch.stft's reflect-padding error
        x = x.contiguous()
        dev = x.device
        bins, frames = self.n_fft // 2 + 1, 1 + T // self.hop_length
        with torch.cuda.device(dev):
            nbytes = lib.dmst_spectrogram_workspace_bytes(bs, chs, T, self.n_fft, self.hop_length)
            if nbytes == 0:
                raise ValueError("dmst_spectrogram_frontend: unsupported STFT geometry (n_fft must be a power of two)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            out = torch.empty(bs, bins + 2, frames + 2, chs, dtype=torch.float32, device=dev)
            _lib.check(lib.dmst_spectrogram_frontend(_ptr(x), T, _ptr(self.window.float().contiguous()), bs, chs, T,
                                                     self.n_fft, self.hop_length, 1e-8, 0.3, _ptr(out), _ptr(ws), nbytes,
                                                     _stream(dev)), "dmst_spectrogram_frontend")
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        bs, chs, seq_len = x.size()
        if not self.input_batchnorm and not (torch.is_grad_enabled() and x.requires_grad):
            return self.model.forward_padded_nhwc(self._frontend(x))
        X = torch.stft(x.reshape(-1, seq_len), n_fft=self.n_fft, hop_length=self.hop_length, window=self.window,
                       return_complex=True)
        X = X.view(bs, chs, X.shape[-2], X.shape[-1])
        X = torch.pow(X.abs() + 1e-8, 0.3)
        if self.input_batchnorm:
            X = self.bn(X)
        return self.model(X)
