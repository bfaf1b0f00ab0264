"""Data-parallel training step around the hot path.

The reference trains with Lightning DDP (configs/config.yaml:40-41): one process per GPU, the batch sharded
over ranks, and ONE collective on the step: the NCCL all-reduce of the model gradients (two Cnn14 encoders +
the transformer controller: 190.9 M parameters = 763.7 MB float32; the mix console has no parameters and the
console / loss kernels need no exchange).  This module is that step, with the pieces of this package in it:

    random reference mix (no grad, device RNG)  ->  peak normalise          mst/system.py:232-253
    A/B split                                                                mst/system.py:255-258
    SpectrogramEncoder x2 (tensor-core Cnn14)  ->  TransformerController     mst/modules.py:31-68
    AdvancedMixConsole(tracks_b, predicted parameters)                       mst/system.py:280-292
    MRSTFT / AudioFeatureLoss                                                mst/system.py:332-338
    backward, bucketed gradient all-reduce overlapped with it, clip (10.0) + Adam   config.yaml:33, system.py:419-424

``TransformerController`` and ``MixStyleTransferModel`` keep the reference's constructor arguments and
parameter names (mst/modules.py:17-68, 809-914), so its checkpoints load; they are the stock
``torch.nn.TransformerEncoder`` / ``Linear`` layers (library GEMMs: 37.9 M parameters over at most 36 tokens,
SURVEY.md section 2 row 11: not a hot path, no kernel of ours).

``BucketedGradAllReduce`` is the collective: gradients live in flat 25 MiB buckets (``p.grad`` are views, so
nothing is copied), filled in the order backward produces them; the moment a bucket's last gradient has been
accumulated its all-reduce (average) is issued on NCCL's stream and overlaps the rest of backward.  In Cnn14
90 % of the parameters sit in the two deepest blocks, whose backward runs FIRST, and 90 % of the FLOPs in the
two shallowest, whose backward runs LAST, so most of the 764 MB moves under the convolution backward.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from .mixing import random_reference_mix


class TransformerController(torch.nn.Module):
    """mst/modules.py:809-914 (stock layers; same names and shapes)."""

    def __init__(self, embed_dim: int, num_track_control_params: int, num_fx_bus_control_params: int,
                 num_master_bus_control_params: int, num_layers: int = 6, nhead: int = 8,
                 use_fx_bus: bool = False, use_master_bus: bool = False) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.num_track_control_params = num_track_control_params
        self.num_fx_bus_control_params = num_fx_bus_control_params
        self.num_master_bus_control_params = num_master_bus_control_params
        self.num_layers, self.nhead = num_layers, nhead
        self.use_fx_bus, self.use_master_bus = use_fx_bus, use_master_bus
        self.track_embedding = torch.nn.Parameter(torch.randn(1, 1, embed_dim))
        self.mix_embedding = torch.nn.Parameter(torch.randn(1, 2, embed_dim))
        self.fx_bus_embedding = torch.nn.Parameter(torch.randn(1, 1, embed_dim))
        self.master_bus_embedding = torch.nn.Parameter(torch.randn(1, 1, embed_dim))
        layer = torch.nn.TransformerEncoderLayer(d_model=embed_dim, nhead=nhead, batch_first=True, dropout=0.0)
        self.transformer_encoder = torch.nn.TransformerEncoder(layer, num_layers=num_layers)
        self.track_projection = torch.nn.Linear(embed_dim, num_track_control_params)
        self.fx_bus_projection = torch.nn.Linear(embed_dim, num_fx_bus_control_params)
        self.master_bus_projection = torch.nn.Linear(embed_dim, num_master_bus_control_params)

    def forward(self, track_embeds, mix_embeds, track_padding_mask: Optional[torch.Tensor] = None):
        bs, num_tracks, _ = track_embeds.shape
        tokens = torch.cat((track_embeds + self.track_embedding, mix_embeds + self.mix_embedding,
                            self.fx_bus_embedding.expand(bs, -1, -1), self.master_bus_embedding.expand(bs, -1, -1)), dim=1)
        if track_padding_mask is not None:   # the four extra tokens are always attended to
            extra = torch.zeros(bs, 4, dtype=torch.bool, device=track_padding_mask.device)
            track_padding_mask = torch.cat((track_padding_mask, extra), dim=1)
        out = self.transformer_encoder(tokens, src_key_padding_mask=track_padding_mask)
        return (torch.sigmoid(self.track_projection(out[:, :num_tracks])),
                torch.sigmoid(self.fx_bus_projection(out[:, -2])),
                torch.sigmoid(self.master_bus_projection(out[:, -1])))


class MixStyleTransferModel(torch.nn.Module):
    """mst/modules.py:17-68: track encoder over (bs * tracks) waveforms, mix encoder over the two channels of the
    reference mix (or its mid / side with sum_and_diff), controller over the embeddings."""

    def __init__(self, track_encoder: torch.nn.Module, mix_encoder: torch.nn.Module, controller: torch.nn.Module,
                 sum_and_diff: bool = False) -> None:
        super().__init__()
        self.track_encoder, self.mix_encoder, self.controller = track_encoder, mix_encoder, controller
        self.sum_and_diff = sum_and_diff

    def forward(self, tracks, ref_mix, track_padding_mask: Optional[torch.Tensor] = None):
        bs, num_tracks, seq_len = tracks.shape
        track_embeds = self.track_encoder(tracks.reshape(bs * num_tracks, 1, seq_len)).view(bs, num_tracks, -1)
        if self.sum_and_diff:
            mid = self.mix_encoder(ref_mix.sum(dim=1, keepdim=True))
            side = self.mix_encoder(ref_mix[:, 0:1] - ref_mix[:, 1:2])
            mix_embeds = torch.stack((mid, side), dim=1)
        else:
            mix_embeds = self.mix_encoder(ref_mix.reshape(bs * 2, 1, -1)).view(bs, 2, -1)
        return self.controller(track_embeds, mix_embeds, track_padding_mask)


class BucketedGradAllReduce:
    """Gradient all-reduce (average) in flat buckets, overlapped with backward (the one collective of the
    reference's DDP step).  Unused parameters (the fx-bus head while the fx bus is off:
    the reference runs ddp_find_unused_parameters_true) keep a zero gradient and their bucket is flushed at
    ``finish``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20, group=None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.native_avg = on and dist.get_backend(group) == "nccl"   # (gloo, used by the CPU tests, has no AVG)
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[dict] = []
        self.bucket_of = {}
        order = list(reversed(self.params))    # roughly the order in which backward finishes them
        cur, cur_bytes = [], 0
        for p in order:
            nb = p.numel() * p.element_size()
            if cur and cur_bytes + nb > bucket_bytes:
                self._make_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            self._make_bucket(cur)
        self.handles = []
        self.exposed_ms = None
        self._ev = None
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _make_bucket(self, ps):
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)   # autograd accumulates into the bucket in place
            off += p.numel()
        b = {"flat": flat, "params": ps, "pending": len(ps), "launched": False}
        for p in ps:
            self.bucket_of[p] = b
        self.buckets.append(b)

    @property
    def total_bytes(self):
        return sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)

    def zero_grad(self):
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["launched"] = len(b["params"]), False
            off = 0
            for p in b["params"]:   # (an optimizer or a caller may have replaced .grad)
                if p.grad is None or p.grad.data_ptr() != b["flat"].data_ptr() + off * b["flat"].element_size():
                    p.grad = b["flat"][off:off + p.numel()].view_as(p)
                off += p.numel()
        self.handles = []

    def _launch(self, b):
        b["launched"] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self.native_avg else dist.ReduceOp.SUM
            self.handles.append((dist.all_reduce(b["flat"], op=op, group=self.group, async_op=True), b))

    def _hook(self, p):
        b = self.bucket_of[p]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"]:
            self._launch(b)

    def finish(self, time_exposed: bool = False):
        """Flush the buckets backward never completed, then make the current stream wait for every all-reduce.
        time_exposed: also measure (device time) how long the stream had to wait, i.e. the part of the
        collective that backward did not hide."""
        for b in self.buckets:
            if not b["launched"]:
                self._launch(b)
        if time_exposed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for h, b in self.handles:
            h.wait()
            if not self.native_avg:
                b["flat"].div_(self.world)
        if time_exposed:
            e1.record()
            self._ev = (e0, e1)
        self.handles = []

    def read_exposed_ms(self):
        if self._ev is None:
            return None
        self._ev[1].synchronize()
        return self._ev[0].elapsed_time(self._ev[1])


def reduce_loss_dict(loss):
    """mst/system.py:334-338: a dictionary loss (AudioFeatureLoss) is the sum of its terms' means."""
    if isinstance(loss, dict):
        return sum(v.mean() for v in loss.values())
    return loss


def training_step(model, console, loss_fn, tracks, reducer: BucketedGradAllReduce, optimizer,
                  generator: Optional[torch.Generator] = None, clip: float = 10.0, time_exposed: bool = False):
    """One step of mst/system.py:102-338 (generate_mix=True) + the optimizer step of Lightning's loop.
    Flags as the shipped configs run them from epoch 0 (configs/models/naive.yaml:5-8): EQ, compressor, master bus on,
    fx bus off; the reference mix without input / output fader (system.py:232-246), the predicted mix with both."""
    T = tracks.shape[-1]
    mid = T // 2
    ref_mix, has_nan, _ = random_reference_mix(tracks, console, generator=generator, use_track_input_fader=False,
                                               use_fx_bus=False, use_ouput_fader=False)
    ref_a, ref_b, tracks_b = ref_mix[..., :mid], ref_mix[..., mid:], tracks[..., mid:]
    reducer.zero_grad()
    track_params, fx_params, master_params = model(tracks_b, ref_a)
    mix_b = console(tracks_b, track_params, fx_params, master_params, use_fx_bus=False)[1]
    loss = reduce_loss_dict(loss_fn(mix_b, ref_b))
    loss.backward()
    reducer.finish(time_exposed=time_exposed)
    torch.nn.utils.clip_grad_norm_(reducer.params, clip, foreach=True)
    optimizer.step()
    return loss.detach(), has_nan
