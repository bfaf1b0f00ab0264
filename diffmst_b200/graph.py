"""CUDA-graph replay of a fixed-shape console/loss step ("use CUDA graphs
for the fixed-shape step").

A training step on this path is ~25 kernel launches of 5-500 us each plus the autograd and
allocator work around them; at 1.4 ms per step the launch gaps are ~10 % of the time.  Every C-ABI
entry point of libdiffmst_b200.so is capture-safe (stream-ordered work only: kernels, memsets,
cuFFT executions on plans created beforehand, event fork/join onto the library's side streams; no
allocation, host synchronisation or host read inside a call), so a whole forward + loss + backward
can be captured once and replayed.

    step = GraphedStep(lambda: loss_fn(console(tracks, tp, fp, mp, ...)[1], target), params=[tp, mp])
    tracks.copy_(next_batch); tp.data.copy_(...)          # refill the static inputs in place
    loss = step()                                          # replay; loss, tp.grad, mp.grad are static tensors

The console's synchronous range check (`check_ranges=True`, one host read per call) cannot run inside a
capture: the callable is warmed up eagerly first, with the check as the caller left it, and for the capture
itself a console whose check is on runs it in its asynchronous, device-side form (`check_ranges="async"`: the
verdict of every replay lands in pinned host memory; `console.check_pending_ranges()` raises the reference's
ValueError for it).
"""
from typing import Callable, Iterable, Optional

import torch


class GraphedStep:
    """Capture `loss_fn()` (forward) and `loss.backward()` into one CUDA graph.

    loss_fn:  no-argument callable reading *static* CUDA tensors (refill them in place between
              replays) and returning a scalar loss tensor.
    params:   leaf tensors whose ``.grad`` the step produces; after every replay ``p.grad`` holds the
              gradient of that replay (a static tensor: copy it out before the next replay).
    consoles: modules whose synchronous range check is replaced by the asynchronous one during capture.
    """

    def __init__(self, loss_fn: Callable[[], torch.Tensor], params: Iterable[torch.Tensor],
                 consoles: Iterable[torch.nn.Module] = (), warmup: int = 3,
                 device: Optional[torch.device] = None):
        self.params = list(params)
        if not self.params:
            raise ValueError("GraphedStep needs at least one parameter to differentiate")
        dev = device if device is not None else self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("diffmst_b200: GraphedStep needs CUDA tensors; there is no CPU path")
        self.device = dev
        consoles = list(consoles)
        # eager warm-up on a side stream (cuFFT plans, side streams and the allocator's pools come into
        # being here, outside the capture)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                for p in self.params:
                    p.grad = None
                loss_fn().backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        saved = [getattr(c, "check_ranges", None) for c in consoles]
        for c in consoles:
            if getattr(c, "check_ranges", False):
                c.check_ranges = "async"
                if hasattr(c, "reserve_capture_slots"):
                    c.reserve_capture_slots()
        try:
            for p in self.params:
                p.grad = None  # the captured backward allocates .grad from the graph's private pool
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = loss_fn()
                self.loss.backward()
        finally:
            for c, s in zip(consoles, saved):
                if s is not None:
                    c.check_ranges = s
        self.grads = [p.grad for p in self.params]

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g  # (a caller may have set .grad to None in between)
        return self.loss
