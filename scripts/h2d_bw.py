"""Raw pinned host->device copy bandwidth (context for bench.py's e2e figure)."""
import torch
x = torch.empty(8 * 16 * 262144, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for name, stream in (("default", torch.cuda.current_stream()), ("side", s)):
    with torch.cuda.stream(stream):
        for _ in range(3):
            d.copy_(x, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            d.copy_(x, non_blocking=True)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {name}: {x.numel()*4/1e6:.1f} MB in {ms:.3f} ms = {x.numel()*4/ms/1e6:.1f} GB/s")
