#!/usr/bin/env python
"""Benchmark of the Diff-MST hot path on B200 (driver contract: see DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): AdvancedMixConsole forward + backward with the
multi-resolution STFT loss, batch 8 x 16 tracks x 262144 samples per GPU, float32, synthetic
white-noise tracks (0.1 * randn), uniform random parameters, fx bus off (as every shipped
config).  One "step" = console forward -> MRSTFT(mix, target mix) -> backward to the console
parameters.  Metric: track-seconds per second = B*N*T / 44100 / step time, summed over GPUs
(weak scaling, batch-sharded, no collective on the path).

`--impl reference` times the reference algorithm's CPU path instead: the float32 oracle port
(oracle/, the reference classes cannot travel to the GPU box because their third-party DSP
dependencies are not installable) on all host cores, on the same configuration (batch 8), a bounded
number of steps.

Timing hygiene: the range test of mst/modules.py:86-89 stays ON inside every timed region (device-side,
asynchronous: `check_ranges="async"`, verdict read after the region); the timed region lasts at least
`--min-seconds` (default 1 s) whatever `--steps` says (`steps` in the JSON line is the number actually
timed, `steps_requested` what was asked for); DRAM traffic and instruction counts of the dominant kernel
come from the committed ncu summary under profiles/ (never hard-coded).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
B, N, T = 8, 16, 262144
RES = dict(fft_sizes=[512, 2048, 8192], hop_sizes=[256, 1024, 4096], win_lengths=[512, 2048, 8192])
METRIC = "track-seconds/sec AdvancedMixConsole fwd+bwd (16trk,262144) @1/2/4/8 B200"
UNIT = "track-seconds/s"
FLAGS = dict(use_track_input_fader=True, use_track_eq=True, use_track_compressor=True,
             use_track_panner=True, use_master_bus=True, use_fx_bus=False, use_output_fader=True)
# algorithmic bytes per track-sample (SURVEY.md section 8d / BASELINE.md section 3, bus-only mode)
BYTES_FWD = 4.0 + 8.0 / N
BYTES_BWD = 4.0 + 8.0 / N


def ncu_summary(kernel_substring):
    """(dram bytes per launch, warp-instructions per launch, file) of the dominant kernel from the newest committed
    profiles/ncu_r*_summary.json (written by scripts/make_profile_summary.py from an `ncu --set full` capture), or
    (None, None, None) when no capture of this kernel exists."""
    import glob
    import re
    def rnd(p):
        m = re.search(r"ncu_r(\d+)", os.path.basename(p))
        return int(m.group(1)) if m else -1
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*_summary.json")), key=rnd, reverse=True):
        try:
            with open(path) as f:
                kernels = json.load(f).get("kernels", [])
        except (OSError, ValueError):
            continue
        for k in kernels:
            if kernel_substring in k.get("kernel", ""):
                num = lambda v: float(str(v).replace(",", ""))
                mb = num(k["dram_read_MB"]) + num(k["dram_write_MB"])
                return mb * 1e6, num(k["warp_instructions"]), os.path.relpath(path, ROOT)
    return None, None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []      # (host time, fields)
        self.proc = None
        self.windows = []   # (t0, t1) host times of the timed regions

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if self.windows:
            inside = [r for r in rows if any(w[0] <= r[0] <= w[1] + 0.05 for w in self.windows)]
            rows = inside if inside else rows[-3:]   # regions shorter than the sampling period
        for _, r in rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(torch, seed, batch, device):
    g = torch.Generator().manual_seed(seed)
    tracks = torch.randn(batch, N, T, generator=g) * 0.1
    tp = torch.rand(batch, N, 27, generator=g)
    fp = torch.rand(batch, 25, generator=g)
    mp = torch.rand(batch, 26, generator=g)
    # target mix: a second random mix of the same tracks (mst/system.py:232-249)
    tp2, mp2 = torch.rand(batch, N, 27, generator=g), torch.rand(batch, 26, generator=g)
    return tracks, tp, fp, mp, tp2, mp2


def run_reference(args):
    """CPU path of the reference algorithm (oracle port, float32) on the arm's own configuration (batch 8: about
    10 GB of temporaries, the box has 196 GB), bounded number of steps."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.auraloss.freq import MultiResolutionSTFTLoss
    from oracle.console import OracleAdvancedMixConsole
    from oracle.loss import batch_stereo_peak_normalize
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = B
    tracks, tp, fp, mp, tp2, mp2 = make_inputs(torch, 0, batch, "cpu")
    con = OracleAdvancedMixConsole(SR)
    loss_fn = MultiResolutionSTFTLoss(**RES)
    with torch.no_grad():
        target = batch_stereo_peak_normalize(con(tracks, tp2, fp, mp2, **FLAGS)[1])
    tp.requires_grad_(True); mp.requires_grad_(True)

    def step():
        tp.grad = None; mp.grad = None
        mix = con(tracks, tp, fp, mp, **FLAGS)[1]
        loss = loss_fn(mix, target)
        loss.backward()
        return float(loss.detach())

    for _ in range(max(1, min(args.warmup, 1))):
        step()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = batch * N * T / SR / dt
    sample = f"the full batch of {batch} x {N} tracks x {T} samples, fwd+bwd+MRSTFT, float32, {steps} steps after 1 warm-up"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "AdvancedMixConsole fwd+bwd + MRSTFT loss, batch 8 x 16 tracks x 262144 "
                                   "samples (BASELINE configs[1])", "global_batch": batch, "tracks": N, "samples": T},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(torch):
    from oracle.auraloss.freq import MultiResolutionSTFTLoss
    from oracle.console import OracleAdvancedMixConsole
    from oracle.loss import batch_stereo_peak_normalize
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tracks, tp, fp, mp, tp2, mp2 = make_inputs(torch, 0, B, "cpu")
    con = OracleAdvancedMixConsole(SR)
    loss_fn = MultiResolutionSTFTLoss(**RES)
    with torch.no_grad():
        target = batch_stereo_peak_normalize(con(tracks, tp2, fp, mp2, **FLAGS)[1])
    tp.requires_grad_(True); mp.requires_grad_(True)
    times = []
    for i in range(3):
        tp.grad = None; mp.grad = None
        t0 = time.perf_counter()
        loss_fn(con(tracks, tp, fp, mp, **FLAGS)[1], target).backward()
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times[1:])
    return {"value": B * N * T / SR / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"the full batch ({B} x {N} tracks x {T} samples) fwd+bwd+MRSTFT, float32 oracle port on {cores} host "
                      f"threads, median of 2 after 1 warm-up ({dt:.2f} s/step)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffmst_b200 import AdvancedMixConsole, GraphedStep, MRSTFTLoss, batch_stereo_peak_normalize, _lib
    from diffmst_b200.dist_util import max_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    con = AdvancedMixConsole(SR).to(dev)
    con.materialize_tracks = False   # bus-only mode
    con.check_ranges = "async"       # the range test of mst/modules.py:86-89 runs on the device in every step
    con_full = AdvancedMixConsole(SR).to(dev)   # the reference's full return contract: mixed_tracks (B,2,N,T) materialised
    con_full.check_ranges = "async"
    loss_fn = MRSTFTLoss(**RES)
    tracks_h, tp_h, fp_h, mp_h, tp2, mp2 = make_inputs(torch, rank, B, "cpu")
    tracks = tracks_h.to(dev); fp = fp_h.to(dev)
    tp = tp_h.to(dev).requires_grad_(True); mp = mp_h.to(dev).requires_grad_(True)
    with torch.no_grad():
        target = batch_stereo_peak_normalize(con(tracks, tp2.to(dev), fp, mp2.to(dev), **FLAGS)[1])

    def step(x, console=con):
        tp.grad = None; mp.grad = None
        mix = console(x, tp, fp, mp, **FLAGS)[1]
        loss = loss_fn(mix, target)
        loss.backward()
        return loss

    def timed(fn, steps):
        """`steps` calls of fn between barriers; device time (CUDA events), max over ranks, and the host window."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        sampler.windows.append((h0, time.perf_counter()))
        return max_over_ranks(e0.elapsed_time(e1), dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    lib = _lib.lib()
    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        step(tracks)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs a moment to come up: start it before the last warm-up
    for _ in range(3):
        step(tracks)
    barrier()
    # number of timed steps: what was asked for, but at least --min-seconds of device time (a 20-step region of this
    # workload is 25 ms: three clock samples and a cold-boost burst); every rank agrees on the count
    probe_ms = timed(lambda: step(tracks), 3) / 3
    sampler.windows.clear()
    steps = max(args.steps, int(-(-args.min_seconds * 1e3 // max(probe_ms, 1e-3))))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_eager_max = timed(lambda: step(tracks), steps)
    # per-kernel device times: a second pass with the library's event brackets around the three chain kernels (the
    # brackets serialise them: in the timed regions the track backward kernel runs beside the master backward kernel)
    lib.dmst_profile_enable(steps)
    for _ in range(steps):
        step(tracks)
    barrier()
    kern_ms = {}
    buf = (ctypes.c_float * steps)()
    # kind 0: fused forward kernel (tracks + master bus), 2: master-bus backward, 3: track backward
    for kind, name in ((0, "console_fwd"), (2, "master_bwd"), (3, "track_bwd")):
        n = lib.dmst_profile_read(kind, buf, steps)
        kern_ms[name] = sum(buf[i] for i in range(n)) / n if n > 0 else None
    lib.dmst_profile_enable(0)
    # full-contract mode (mst/modules.py:314 returns the panned tracks (B,2,N,T)): same step, mixed_tracks materialised
    for _ in range(3):
        step(tracks, con_full)
    ms_full_max = timed(lambda: step(tracks, con_full), max(steps // 4, 10))
    steps_full = max(steps // 4, 10)

    # ---------------- device-resident timing, the step replayed as one CUDA graph ----------------
    # (the public GraphedStep of the package: same kernels, same tensors, no launch gaps)
    graph_err = None
    try:
        graphed = GraphedStep(lambda: loss_fn(con(tracks, tp, fp, mp, **FLAGS)[1], target), params=[tp, mp],
                              consoles=[con], warmup=2)
        for _ in range(max(args.warmup, 3)):
            graphed()
        torch.cuda.synchronize(dev)
    except Exception as e:   # a box where capture is not possible must still produce a (eager) number
        graph_err = repr(e)[:300]
    ok = torch.tensor([0 if graph_err else 1], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # every rank takes the same branch
    use_graph = int(ok.item()) == 1
    if use_graph:
        ms_max = timed(graphed, steps)
        step_loss = float(graphed.loss.detach())
        launch_note = ("the step replayed as one CUDA graph (diffmst_b200.GraphedStep); eager launches of the same "
                       "step: see `eager`")
    else:
        ms_max = ms_eager_max
        step_loss = float(step(tracks).detach())
        launch_note = f"eager launches (CUDA-graph capture failed on some rank: {graph_err})"

    # the asynchronous range verdicts of every step above (eager, full contract, graph replays)
    con.check_pending_ranges()
    con_full.check_pending_ranges()
    # clocks / throttle reasons: the samples that fall inside the timed regions (eager, full contract, graph replay)
    if rank == 0:
        time.sleep(0.06)         # let the sampler deliver the sample taken during the region
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end to end: host buffers, H2D of the step's inputs, D2H of its results ----
    pinned = [tracks_h.pin_memory(), tracks_h.clone().pin_memory()]
    tp_pin, mp_pin = tp_h.pin_memory(), mp_h.pin_memory()
    dev_bufs = [torch.empty_like(tracks), torch.empty_like(tracks)]
    # the parameters are leaves of the autograd graph: one pair per pipeline slot, uploaded on the copy
    # stream AHEAD of the slot's tracks (a small copy issued on the compute stream would queue behind the
    # next step's 134 MB upload in the H2D copy engine and serialise copy and compute)
    tp_bufs = [torch.empty_like(tp).requires_grad_(True) for _ in range(2)]
    mp_bufs = [torch.empty_like(mp).requires_grad_(True) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    host_out = {"loss": torch.empty(1).pin_memory(), "gtp": torch.empty(B, N, 27).pin_memory(),
                "gmp": torch.empty(B, 26).pin_memory()}
    h2d = tracks_h.numel() * 4 + tp_h.numel() * 4 + mp_h.numel() * 4
    d2h = 4 + (B * N * 27 + B * 26) * 4

    def e2e_step(x, tpb, mpb):
        tpb.grad = None; mpb.grad = None
        mix = con(x, tpb, fp, mpb, **FLAGS)[1]
        loss = loss_fn(mix, target)
        loss.backward()
        return loss

    def e2e_loop(steps):
        # two-deep pipeline: the upload of step i+1's inputs overlaps the compute of step i
        main = torch.cuda.current_stream(dev)
        for b in range(2):
            freed[b].record(main)
        def upload(i):
            b = i & 1
            with torch.cuda.stream(copy_stream), torch.no_grad():
                copy_stream.wait_event(freed[b])
                tp_bufs[b].copy_(tp_pin, non_blocking=True)
                mp_bufs[b].copy_(mp_pin, non_blocking=True)
                dev_bufs[b].copy_(pinned[b], non_blocking=True)
                ready[b].record(copy_stream)
        upload(0)
        for i in range(steps):
            b = i & 1
            if i + 1 < steps:
                upload(i + 1)
            main.wait_event(ready[b])
            loss = e2e_step(dev_bufs[b], tp_bufs[b], mp_bufs[b])
            host_out["loss"].copy_(loss.detach().reshape(1), non_blocking=True)
            host_out["gtp"].copy_(tp_bufs[b].grad, non_blocking=True)
            host_out["gmp"].copy_(mp_bufs[b].grad, non_blocking=True)
            freed[b].record(main)
        torch.cuda.synchronize(dev)

    e2e_loop(max(args.warmup, 3))
    e2e_steps = max(args.steps, int(-(-args.min_seconds * 1e3 // max(2 * probe_ms, 1e-3))))
    barrier()
    ev0.record()
    e2e_loop(e2e_steps)
    ev1.record()
    barrier()
    e2e_ms_max = max_over_ranks(ev0.elapsed_time(ev1), dev)
    con.check_pending_ranges()

    if rank == 0:
        units = world * B * N * T / SR  # track-seconds per step over all ranks
        value = units / (ms_max / 1e3 / steps)
        e2e_value = units / (e2e_ms_max / 1e3 / e2e_steps)
        peak, peak_src = measured_peaks()
        samples = B * N * T
        roof = None
        if kern_ms.get("track_bwd"):
            achieved = BYTES_BWD * samples / (kern_ms["track_bwd"] * 1e-3) / 1e9
            # dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of one launch of this kernel,
            # from the committed `ncu --set full` capture (profiles/); null when there is no capture of this kernel
            traffic, winst, src = ncu_summary("track_bwd2_kernel")
            sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
            roof = {"bound": "hbm", "kernel": "track_bwd2_kernel (per-track chain backward, parameter gradients)",
                    "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": BYTES_BWD * samples,
                    "avg_launch_ms": kern_ms["track_bwd"],
                    "traffic": traffic, "traffic_source": src,
                    "kernel_ms": kern_ms,
                    # honest co-limit: the kernel is bound by FP32 issue / pipe, not by HBM:
                    # executed warp-instructions of one launch (ncu) / (4 schedulers x SMs x measured SM clock)
                    "issue_rate": None if winst is None else {
                        "warp_instructions_per_launch": winst,
                        "peak_warp_instructions_per_s": 4 * 148 * sm_hz * 1e6,
                        "frac": winst / (kern_ms["track_bwd"] * 1e-3) / (4 * 148 * sm_hz * 1e6)},
                    "step_frac_of_hbm_roofline": (BYTES_FWD + BYTES_BWD) * samples / (ms_max / steps * 1e-3) / 1e9 / peak}
        bytes_full = (BYTES_FWD + 8.0) + BYTES_BWD   # + the (B,2,N,T) write: 17 B per track-sample at N = 16
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                "steps_requested": args.steps, "min_seconds": args.min_seconds,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_max / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "AdvancedMixConsole fwd+bwd + MRSTFT loss, batch 8 x 16 tracks x 262144 "
                                       "samples per GPU (BASELINE configs[1])",
                           "global_batch": world * B, "tracks": N, "samples": T, "parallelism": f"dp{world}",
                           "mode": "bus-only (mixed_tracks not materialised); full-contract mode: see `full_contract`",
                           "range_check": "on in every timed step (device-side, asynchronous verdict; mst/modules.py:86-89)",
                           "launch": launch_note,
                           "l2": "inputs larger than L2 (134 MB of tracks per step, re-read every step)"},
                "eager": {"ms_per_step": ms_eager_max / steps,
                          "value": units / (ms_eager_max / 1e3 / steps), "unit": UNIT,
                          "note": "same step launched eagerly from Python; the per-kernel times of `roofline` were "
                                  "taken over this region (`clocks`: samples inside this region and the graph-replay region)"},
                "full_contract": {"ms_per_step": ms_full_max / steps_full, "steps": steps_full,
                                  "value": units / (ms_full_max / 1e3 / steps_full), "unit": UNIT,
                                  "algorithmic_bytes_per_track_sample": bytes_full,
                                  "step_frac_of_hbm_roofline": bytes_full * samples / (ms_full_max / steps_full * 1e-3) / 1e9 / peak,
                                  "note": "same step launched eagerly with mixed_tracks (B,2,N,T) materialised, as "
                                          "mst/modules.py:314 returns it (268 MB more written per step)"},
                "loss": step_loss,
                "clocks": clocks, "gpu_launches": GPU_LAUNCHES_PER_STEP * steps,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_max / e2e_steps, "steps": e2e_steps,
                        "note": "pinned host tracks + parameters copied in every step (copy of step i+1 overlaps "
                                "compute of step i), loss and parameter gradients copied out every step"},
                "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(torch)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# our kernels per step: console forward 3 (prepare, fused chain kernel, range verdict), MRSTFT 8 (per
# resolution the fused STFT + loss kernel and the fused gradient + inverse-FFT kernel; the loss totals; one overlap-add for all at
# backward), console backward 4 (recursion-table prepare, master chain kernel, track chain kernel beside it, one
# epilogue); torch glue kernels are not counted
GPU_LAUNCHES_PER_STEP = 15


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="lower bound on the device time of each timed region (more steps are run if needed)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
