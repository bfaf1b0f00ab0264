"""Aggregate pinned host->device bandwidth with N ranks copying at once (context for bench.py's e2e figure at N GPUs:
is the end-to-end number bound by the host, not by the pipeline?).  Run under torch.distributed.run.

Arms: (a) all ranks copy 134 MB (one step's tracks) back to back, simultaneously; (b) the same with the pinned
buffer allocated after binding the process to the NUMA node of its GPU (numactl-free: os.sched_setaffinity to the
node's CPUs, then first-touch); (c) staggered: rank r starts r/N of a copy time later."""
import os, sys, time, glob
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 8 * 16 * 262144

def gpu_numa_node():
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dv = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dv:02x}.0/numa_node"
        return int(open(path).read())
    except Exception as e:
        return -1

def run(x, tag, stagger=0.0):
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3):
        d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if stagger:
        time.sleep(stagger * rank / world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    reps = 20
    for _ in range(reps):
        d.copy_(x, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / reps
    gbs = torch.tensor([n * 4 / ms / 1e6], device=dev, dtype=torch.float64)
    tot, mn = gbs.clone(), gbs.clone()
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{tag}: {world} ranks x 134 MB copies: per-rank {float(gbs):.1f} GB/s (min {float(mn):.1f}), aggregate {float(tot):.1f} GB/s", flush=True)

node = gpu_numa_node()
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
if rank == 0:
    print(f"host: {os.cpu_count()} CPUs, {len(nodes)} NUMA node(s); GPU {local} on node {node}", flush=True)
x = torch.empty(n, dtype=torch.float32).pin_memory(); x.fill_(1.0)
run(x, "simultaneous")
run(x, "staggered", stagger=0.0025)
if len(nodes) > 1 and node >= 0:
    cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    ids = []
    for part in cpus.split(","):
        a, _, b = part.partition("-"); ids += list(range(int(a), int(b or a) + 1))
    os.sched_setaffinity(0, ids)
    y = torch.empty(n, dtype=torch.float32).pin_memory(); y.fill_(1.0)   # first touch on the GPU's node
    run(y, f"NUMA-local pinned buffer (node {node})")
# (d) write-combined pinned memory (cudaHostAllocWriteCombined): no cache snooping on the host side of the DMA
try:
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n * 4), ctypes.c_uint(0x04))   # cudaHostAllocWriteCombined
    assert rc == 0, rc
    buf = (ctypes.c_float * n).from_address(ptr.value)
    z = torch.frombuffer(buf, dtype=torch.float32)
    z.fill_(1.0)
    run(z, f"write-combined pinned buffer (is_pinned={z.is_pinned()})")
except Exception as e:
    if rank == 0:
        print("write-combined arm failed:", repr(e)[:200], flush=True)
if world > 1:
    dist.destroy_process_group()
