"""Generate panns.npz: golden vectors for the encoder rows from the REFERENCE's own
unmodified classes - mst.panns.ConvBlock, mst.panns.Cnn14 and the spectrogram lines of
mst.modules.SpectrogramEncoder.forward (mst/modules.py:787-800) - evaluated in float64 on CPU with the
generator-free parameter fill of panns_fill.py.  Runs ONLY in the build container (needs /root/reference); it also
asserts that the oracle restatement (oracle/panns.py) reproduces every number, so the fixture pins both.

    python tests/golden/make_golden_panns.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import mst.panns as ref_panns  # noqa: E402  (unmodified reference)
import mst.modules as ref_modules  # noqa: E402

from oracle.panns import OracleCnn14, OracleConvBlock  # noqa: E402
from panns_fill import fill_state  # noqa: E402


def block_case(mod):
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 3, 10, 7, generator=g, dtype=torch.float64).requires_grad_(True)
    probe = torch.randn(2, 8, 5, 3, generator=g, dtype=torch.float64)
    y = mod(x, (2, 2))
    (y * probe).sum().backward()
    return x, probe, y, {n: p.grad.clone() for n, p in mod.named_parameters()}


def cnn14_case(mod):
    g = torch.Generator().manual_seed(42)
    x = (torch.rand(1, 1, 1024, 128, generator=g) ** 2).double()
    out = mod(x)
    out.square().mean().backward()
    return x, out, {n: p.grad.clone() for n, p in mod.named_parameters()}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    arrays = {}
    # ---- ConvBlock, training mode (batch statistics), tiny: everything stored ----
    ref = ref_panns.ConvBlock(3, 8).double().train()
    fill_state(ref)
    orc = OracleConvBlock(3, 8).double().train()
    fill_state(orc)
    x, probe, y, grads = block_case(ref)
    xo, _, yo, gradso = block_case(orc)
    assert float((y - yo).abs().max()) <= 1e-12 and float((x.grad - xo.grad).abs().max()) <= 1e-12
    for n in grads:
        assert float((grads[n] - gradso[n]).abs().max()) <= 1e-11 * max(1.0, float(grads[n].abs().max())), n
    arrays.update(block_x=x.detach().numpy(), block_probe=probe.numpy(), block_y=y.detach().numpy(),
                  block_gx=x.grad.numpy(), **{"block_g_" + n.replace(".", "_"): v.numpy() for n, v in grads.items()},
                  block_running_mean1=ref.bn1.running_mean.numpy(), block_running_var1=ref.bn1.running_var.numpy())
    # ---- Cnn14, eval mode, smallest input that survives the six poolings: output + gradient summaries ----
    ref = ref_panns.Cnn14(num_classes=6).double().eval()
    fill_state(ref)
    orc = OracleCnn14(num_classes=6).double().eval()
    fill_state(orc)
    x, out, grads = cnn14_case(ref)
    _, outo, gradso = cnn14_case(orc)
    assert float((out - outo).abs().max()) <= 1e-12 * float(out.abs().max())
    names = ["conv_block1.conv1.weight", "conv_block3.conv2.weight", "conv_block6.conv2.weight", "conv_block2.bn1.weight", "fc.weight"]
    for n in names:
        assert float((grads[n] - gradso[n]).abs().max()) <= 1e-10 * float(grads[n].abs().max()), n
    arrays.update(cnn14_out=out.detach().numpy(), cnn14_grad_names=np.array(names),
                  cnn14_grad_sum=np.array([float(grads[n].sum()) for n in names]),
                  cnn14_grad_abs=np.array([float(grads[n].abs().sum()) for n in names]),
                  cnn14_g_first=grads["conv_block1.conv1.weight"].numpy(), cnn14_g_fc=grads["fc.weight"].numpy())
    print("Cnn14 out", out.detach().numpy().ravel(), "|g_first|", float(grads[names[0]].norm()))
    # ---- spectrogram front-end of SpectrogramEncoder (mst/modules.py:787-800): the reference class with its
    #      convolution trunk replaced by an identity ----
    enc = ref_modules.SpectrogramEncoder(embed_dim=8)
    enc.model = torch.nn.Identity()
    g = torch.Generator().manual_seed(43)
    w = torch.randn(2, 1, 16384, generator=g) * 0.1
    with torch.no_grad():
        S = enc(w)
    S2 = torch.pow(torch.stft(w.view(-1, 16384), n_fft=2048, hop_length=512, window=torch.hann_window(2048),
                              return_complex=True).abs().view(2, 1, 1025, -1) + 1e-8, 0.3)
    assert torch.equal(S, S2), float((S - S2).abs().max())
    arrays.update(spec_wave=w.numpy(), spec_out=S.numpy())
    path = os.path.join(HERE, "panns.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote panns.npz  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
