"""world_size-2 gloo test of the data-parallel plumbing (no GPU): batch sharding covers the
batch exactly, timing is the max over ranks, throughput aggregates over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffmst_b200.dist_util import aggregate_throughput, max_over_ranks, mean_over_ranks, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_batch(5, rank, world)
    mx = max_over_ranks(10.0 + rank)
    thr = aggregate_throughput(100.0, 10.0 + 10.0 * rank)
    mean = float(mean_over_ranks(torch.tensor([float(rank)])))
    out[rank] = (lo, hi, mx, thr, mean)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][:2] == (0, 3) and res[1][:2] == (3, 5)
    assert res[0][2] == res[1][2] == 11.0
    assert abs(res[0][3] - 2 * 100.0 / 0.020) < 1e-9 and res[0][3] == res[1][3]
    assert res[0][4] == res[1][4] == 0.5


def test_shard_batch_partitions():
    for gb in (1, 7, 8, 16):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def _reducer_worker(rank, world, port, out):
    """Two replicas of a small model with an unused head (the fx-bus head of the controller while the fx bus is off),
    different data per rank: after backward + finish every rank holds the average of the per-rank gradients."""
    from diffmst_b200.training import BucketedGradAllReduce
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.ModuleDict({"a": torch.nn.Linear(16, 32), "b": torch.nn.Linear(32, 8), "unused": torch.nn.Linear(32, 4)})
    reducer = BucketedGradAllReduce(model.parameters(), bucket_bytes=1024)   # several buckets
    assert len(reducer.buckets) > 2 and reducer.total_bytes == 4 * sum(p.numel() for p in model.parameters())
    x = torch.randn(5, 16, generator=torch.Generator().manual_seed(10 + rank))
    ok = True
    for it in range(2):   # second iteration: zero_grad() re-arms the buckets
        reducer.zero_grad()
        model["b"](torch.relu(model["a"](x))).square().sum().backward()
        reducer.finish()
        # reference: every rank's own gradient, gathered and averaged
        ref = torch.nn.ModuleDict({"a": torch.nn.Linear(16, 32), "b": torch.nn.Linear(32, 8)})
        ref.load_state_dict({k: v for k, v in model.state_dict().items() if not k.startswith("unused")})
        ref["b"](torch.relu(ref["a"](x))).square().sum().backward()
        for name in ("a", "b"):
            own = ref[name].weight.grad
            gathered = [torch.zeros_like(own) for _ in range(world)]
            dist.all_gather(gathered, own)
            ok = ok and torch.allclose(model[name].weight.grad, sum(gathered) / world, rtol=1e-6, atol=1e-7)
        ok = ok and float(model["unused"].weight.grad.abs().max()) == 0.0
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_two_rank_gloo():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_reducer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert res == {0: True, 1: True}
