#!/usr/bin/env python
"""DDP training-step benchmark (BASELINE.json configs[2] and configs[3]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/train_step_bench.py --config 3 --steps 10 --warmup 3

One process per GPU, the batch sharded over ranks (weak scaling: the per-GPU batch is fixed), the step of
diffmst_b200/training.py: reference-mix generation, two tensor-core Cnn14 encoders + the transformer controller,
console forward / backward, loss, the bucketed NCCL all-reduce of the 764 MB of model gradients overlapped with
backward, gradient clipping and Adam, all inside the timed region.  Rank 0 prints one JSON line.

configs[2]: 32 tracks, AudioFeatureLoss, global batch 16 on 8 GPUs = 2 items per GPU.
configs[3]: configs/models/naive.yaml - 8 tracks, MRSTFT, batch_size 4 per GPU (configs/data/medley+cambridge-8.yaml:13).
Both: 262144-sample excerpts, second half through the model (mst/system.py:255-258), float32 (TF32 tensor-core convolutions).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SR, T = 44100, 262144
RES = dict(fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096], win_lengths=[512, 2048, 8192])
CONFIGS = {2: dict(tracks=32, batch=2, loss="afl", name="configs[2]: 32 tracks + AudioFeatureLoss, 2 items per GPU (global batch 16 at 8 GPUs)"),
           3: dict(tracks=8, batch=4, loss="mrstft", name="configs[3]: naive.yaml step, 8 tracks + MRSTFT, 4 items per GPU")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3, choices=[2, 3])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="items per GPU (default: the config's)")
    ap.add_argument("--no-allreduce", action="store_true", help="skip the collective (measures what it costs)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from diffmst_b200 import AdvancedMixConsole, AudioFeatureLoss, MRSTFTLoss, SpectrogramEncoder
    from diffmst_b200.dist_util import max_over_ranks
    from diffmst_b200.training import BucketedGradAllReduce, MixStyleTransferModel, TransformerController, training_step

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    B, N = args.batch or cfg["batch"], cfg["tracks"]

    torch.manual_seed(0)   # identical replicas
    model = MixStyleTransferModel(SpectrogramEncoder(embed_dim=512), SpectrogramEncoder(embed_dim=512),
                                  TransformerController(512, 27, 25, 26, num_layers=12, nhead=8)).to(dev).train()
    console = AdvancedMixConsole(SR).to(dev)
    console.materialize_tracks = False
    console.check_ranges = "async"
    loss_fn = AudioFeatureLoss([0.1, 0.001, 1.0, 1.0, 0.1], SR) if cfg["loss"] == "afl" else MRSTFTLoss(**RES)
    reducer = BucketedGradAllReduce(model.parameters())
    if args.no_allreduce:
        reducer.world = 1
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, betas=(0.9, 0.999), fused=True)   # (torch's fused CUDA Adam: one pass)
    g = torch.Generator().manual_seed(1000 + rank)   # every rank its own shard of the data
    tracks = (torch.randn(B, N, T, generator=g) * 0.1).to(dev)
    gen = torch.Generator(device=dev).manual_seed(2000 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        loss, _ = training_step(model, console, loss_fn, tracks, reducer, opt, generator=gen)
    barrier()
    exposed = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss, has_nan = training_step(model, console, loss_fn, tracks, reducer, opt, generator=gen, time_exposed=True)
        exposed.append(reducer._ev)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    exp_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in exposed) / len(exposed), dev)
    console.check_pending_ranges()
    # replicas must still be identical: the all-reduce is the only thing that keeps them so
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2**30
    if rank == 0:
        units = world * B * N * T / SR
        print(json.dumps({
            "metric": "track-seconds/sec, whole DDP training step (encoders + controller + console + loss + all-reduce + Adam)",
            "value": units / (ms / 1e3), "unit": "track-seconds/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32 (TF32 tensor-core convolutions)",
            "data": "synthetic", "config": {"workload": cfg["name"], "items_per_gpu": B, "tracks": N, "samples": T,
                                            "parallelism": f"dp{world}", "batchnorm": "per-rank batch statistics (SyncBatchNorm of configs/config.yaml:41 not implemented)"},
            "allreduce": {"enabled": world > 1 and not args.no_allreduce, "bytes_per_step": reducer.total_bytes, "buckets": len(reducer.buckets),
                          "exposed_ms_per_step": exp_ms, "note": "exposed = device time the compute stream waited for NCCL after backward (max over ranks)"},
            "parameters": sum(p.numel() for p in model.parameters()), "replicas_in_sync": bool(float(hi - lo) == 0.0) if not args.no_allreduce else None,
            "loss": float(loss), "found_nan_in_ref_mix": bool(has_nan), "peak_memory_gib": peak_gb}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
