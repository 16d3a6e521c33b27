#!/usr/bin/env python
"""bench.py — denoiser-steps/sec of the NaturalSpeech2 hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one `Model.forward` (the per-timestep denoiser call, ns2.py:929-1000) on the per-GPU batch of the
workload BASELINE.json quotes the metric on: configs[1] = Model(dim=512, depth=12, heads=8) unconditional,
seq=1024, batch=32, bf16 tensor-core operands, random-init weights, synthetic latents.  With N GPUs every rank
runs its own batch of 32 (weak scaling, independent samples), computes its local scalar loss (MSE against a
fixed synthetic target) and the ranks all-reduce that 4-byte scalar over NCCL — the only collective the path has.

Printed JSON (one line, rank 0): the base contract keys plus
  roofline      dominant kernel = the FFN causal-conv GEMM (43% of the step's FLOPs): algorithmic FLOPs per launch
                / mean launch time measured with CUDA events inside real steps, against MEASURED_PEAKS.json
  cpu_baseline  the torch-CPU port of the reference's path (oracle/denoiser_torch_port.py, all host threads) on a
                bounded sample (batch 2 of the 32-sample step), N=1 only
  e2e           the same metric with HOST buffers: pinned-host -> device copy of the step's inputs and device ->
                pinned-host copy of the full prediction inside the timed region (double-buffered on side streams)
`--impl reference` times the reference-arm: the oracle port of the reference's CPU implementation on the host
cores of the box (rank 0 only), same metric/unit/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = dict(dim=512, depth=12, heads=8)
BATCH, SEQ = 32, 1024
FLOPS_PER_SAMPLE = 316.37e9          # SURVEY Appendix C, analytic forward FLOPs per sample at N=1024
FF_INNER = 1365                      # int(512 * 4 * 2 / 3)
CONV_FLOPS_PER_LAUNCH = 2.0 * BATCH * SEQ * FF_INNER * (3 * FF_INNER)   # algorithmic (unpadded) FLOPs
WORKLOAD = "configs[1]: Model(dim=512, depth=12, heads=8) unconditional, bf16 operands, seq=1024, batch=32 per GPU"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r01_dominant_kernel_ncu.json); None if the capture is absent."""
    p = ROOT / "profiles" / "r01_dominant_kernel_ncu.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text())
    vals = [(l["dram_read_MB"] + l["dram_write_MB"]) * 1e6 for l in d["launches"]]
    return round(sum(vals) / len(vals))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # median over the busier half of the samples = the clock under load
        sm_sorted = sorted(sm)
        return {"sm_mhz": statistics.median(sm_sorted[: max(1, len(sm_sorted) // 2 + 1)]) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from naturalspeech2_pytorch_b200 import Model, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))

    torch.manual_seed(0)
    model = Model(**CFG).to(dev).eval()
    model.packed()
    model.freeze_packed = True
    model.use_cuda_graphs = not args.no_cuda_graphs   # the step replays one captured graph (111 kernel nodes)
    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    x_host = torch.randn(BATCH, SEQ, CFG["dim"], generator=g).pin_memory()
    t_host = torch.rand(BATCH, generator=g).pin_memory()
    x = x_host.to(dev)
    times = t_host.to(dev)
    target = torch.randn(BATCH, SEQ, CFG["dim"], device=dev)
    loss_rows = torch.empty(BATCH, device=dev)
    mse_scratch = torch.empty(BATCH * 64, device=dev)

    def step():
        out = model(x, times)
        ops.mse_rows(out, target, loss_rows, mse_scratch)
        loss = loss_rows.mean()
        if world > 1:
            dist.all_reduce(loss)  # the path's only collective: 4-byte scalar loss (SUM; mean = / world)
        return out, loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    graphs = model.use_cuda_graphs
    model.use_cuda_graphs = False     # count this library's launches of one step with eager launches
    l0 = ops.launch_count()
    step()
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - l0
    model.use_cuda_graphs = graphs

    # ------------------------------- timed region: K steps, device-resident inputs -------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total / 1e3)

    # ------------------------------- e2e: host buffers in, host buffers out --------------------------
    out_host = [torch.empty(BATCH, SEQ, CFG["dim"]).pin_memory() for _ in range(2)]
    x_dev = [torch.empty_like(x) for _ in range(2)]
    t_dev = [torch.empty_like(times) for _ in range(2)]
    o_dev = [torch.empty_like(x) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def e2e_loop(n):
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(s_in):           # H2D of this step's inputs (pinned host memory)
                if i >= 2:
                    s_in.wait_event(ev_free[b])
                x_dev[b].copy_(x_host, non_blocking=True)
                t_dev[b].copy_(t_host, non_blocking=True)
                ev_in[b].record(s_in)
            main.wait_event(ev_in[b])
            out = model(x_dev[b], t_dev[b])         # the public API call
            ev_free[b].record(main)
            if i >= 2:
                main.wait_event(ev_out[b])          # o_dev[b] has been drained to the host
            o_dev[b].copy_(out)
            ev_done[b].record(main)
            with torch.cuda.stream(s_out):          # D2H of the step's full prediction
                s_out.wait_event(ev_done[b])
                out_host[b].copy_(o_dev[b], non_blocking=True)
                ev_out[b].record(s_out)
        s_out.synchronize()

    e2e_loop(3)
    barrier()
    w0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / float(t_e.item())
    h2d = x_host.numel() * 4 + t_host.numel() * 4
    d2h = out_host[0].numel() * 4

    # ------------------------------- roofline: the FFN conv GEMM inside real steps -------------------
    roof = None
    cpu_base = None
    if rank == 0:
        model._prof = []
        for _ in range(3):
            model(x, times)  # rank-local: no collective here (the other ranks have left the step loop)
        torch.cuda.synchronize()
        conv_ms = [a.elapsed_time(b) for (name, a, b) in model._prof if name == "ff_conv"]
        by_name = {}
        for (name, a, b) in model._prof:
            by_name.setdefault(name, []).append(a.elapsed_time(b))
        model._prof = None
        peaks, peak_src = _peaks()
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        conv_mean = statistics.mean(conv_ms)
        achieved = CONV_FLOPS_PER_LAUNCH / (conv_mean * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_kernel<256,1,BF16> (FFN causal conv k=3 as 3 shifted GEMM segments)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                "flops_per_launch": CONV_FLOPS_PER_LAUNCH, "ms_per_launch": round(conv_mean, 4),
                "traffic": _ncu_traffic(),
                "step_tflops": round(FLOPS_PER_SAMPLE * BATCH / (ms_per_step * 1e-3) / 1e12, 1),
                "step_frac_of_peak": round(FLOPS_PER_SAMPLE * BATCH / (ms_per_step * 1e-3) / 1e12 / peak, 4),
                "per_op_ms_per_step": {k: round(sum(v) / 3, 4) for k, v in sorted(by_name.items())}}
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_baseline(model)

    if rank == 0:
        line = {
            "metric": "denoiser-steps/sec", "value": round(value, 3), "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "seq_len": SEQ,
                       "parallelism": f"dp{world}",
                       "l2": "no flush needed: each step streams ~1.3 GB of activations + 0.5 GB of weights, >> 126 MB L2",
                       "sample_steps_per_s": round(value * BATCH, 1)},
            "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
            "e2e": {"value": round(e2e_value, 3), "unit": "steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "roofline": roof, "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


CPU_SAMPLE_BATCH = 2   # bounded sample of the 32-sample workload step


def _host_threads() -> int:
    """CPU threads this process may really use: affinity mask, capped by the cgroup CPU quota (oversubscribing a
    quota-limited container with one thread per visible core makes the CPU baseline many times slower)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


class _TorchPort:
    """oracle/denoiser_torch_port.py wrapped with the numpy oracle's calling convention."""

    def __init__(self, mod):
        self.mod = mod

    def model_forward(self, P, cfg, x, t, dtype=None):
        return self.mod.model_forward(P, cfg, x, t)


def _oracle_setup(seed_model=None):
    """fp32 CPU parameters of the cfg2 model for the torch port of the reference CPU path (all host cores)."""
    import torch
    from naturalspeech2_pytorch_b200 import Model
    from oracle import denoiser_oracle, denoiser_torch_port
    torch.set_num_threads(_host_threads())
    if seed_model is None:
        torch.manual_seed(0)
        seed_model = Model(**CFG)
    P = {k: v.detach().cpu().float() for k, v in seed_model.state_dict().items()}
    cfg = denoiser_oracle.ModelConfig(**CFG)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(CPU_SAMPLE_BATCH, SEQ, CFG["dim"], generator=g)
    t = torch.rand(CPU_SAMPLE_BATCH, generator=g)
    return _TorchPort(denoiser_torch_port), P, cfg, x, t


def cpu_baseline(model=None, repeats=2):
    """The torch-CPU port of the reference's path on a bounded sample: batch 2 of the 32-sample workload step."""
    oracle, P, cfg, x, t = _oracle_setup(model)
    t0 = time.perf_counter()
    oracle.model_forward(P, cfg, x, t)  # warm-up (thread pool, page faults)
    best = time.perf_counter() - t0
    if best < 30.0:  # keep the leg bounded: re-time only when a call is cheap
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.model_forward(P, cfg, x, t)
            best = min(best, time.perf_counter() - t0)
    return {"value": round(CPU_SAMPLE_BATCH / (best * BATCH), 5), "unit": "steps/s", "cores": _host_threads(),
            "kind": "port",
            "sample": f"torch fp32 CPU port of the reference path (oracle/denoiser_torch_port.py), batch "
                      f"{CPU_SAMPLE_BATCH} x seq 1024 ({best:.2f} s), scaled x{BATCH // CPU_SAMPLE_BATCH} to the 32-sample step"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port), host cores only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle, P, cfg, x, t = _oracle_setup()
    t0 = time.perf_counter()
    oracle.model_forward(P, cfg, x, t)  # warm-up, also sizes the timed loop
    first = time.perf_counter() - t0
    # each timed call is a bounded sample (batch 2 = 1/16 of a workload step); keep the whole arm under ~2 minutes
    steps = max(1, min(args.steps, 10, int(90.0 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.model_forward(P, cfg, x, t)
    dt = (time.perf_counter() - t0) / steps
    value = CPU_SAMPLE_BATCH / (dt * BATCH)
    sample = (f"torch fp32 CPU port of the reference path (oracle/denoiser_torch_port.py, all host threads), batch "
              f"{CPU_SAMPLE_BATCH} x seq 1024 per timed call ({dt:.2f} s), scaled x{BATCH // CPU_SAMPLE_BATCH} to the workload step")
    line = {"impl": "reference", "metric": "denoiser-steps/sec", "value": round(value, 5), "unit": "steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": round(dt * BATCH / CPU_SAMPLE_BATCH * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH, "seq_len": SEQ, "parallelism": "cpu"},
            "cpu_baseline": {"value": round(value, 5), "unit": "steps/s", "cores": _host_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(value, 5), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graphs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
