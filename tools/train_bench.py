#!/usr/bin/env python
"""Forward + backward step of the unconditional cfg2 denoiser (Model(512, depth 12, heads 8), B=32, N=1024) through
`NaturalSpeech2.forward(...).backward()`; prints ms per phase and, with --prof, CUDA-event time per kernel family."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import Model, NaturalSpeech2, ops  # noqa: E402

prof = "--prof" in sys.argv
B = int(next((a.split("=")[1] for a in sys.argv if a.startswith("--batch=")), 32))
torch.manual_seed(0)
model = Model(dim=512, depth=12, heads=8).cuda().train()
ns = NaturalSpeech2(model, target_sample_hz=24000)
g = torch.Generator().manual_seed(1)
lat = torch.randn(B, 1024, 512, generator=g).cuda()
times, noise = torch.rand(B, generator=g), torch.randn(B, 1024, 512, generator=g)

acc = {}
if prof:
    def wrap(name):
        fn = getattr(ops, name)

        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            acc.setdefault(name, []).append((e0, e1))
            return out
        setattr(ops, name, inner)
    for n in ("gemm", "wgrad", "attention", "attention_bwd", "rmsnorm_film", "rmsnorm_film_bwd", "geglu_bwd",
              "wavenet_gate_bwd", "colsum", "group_sum", "film_wgrad", "cast_bf16", "mse_bwd", "mse_rows"):
        wrap(n)


def step():
    for p in model.parameters():
        p.grad = None
    loss = ns(lat, times=times, noise=noise)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize()
    return loss, t1


for _ in range(2):
    step()
acc.clear()
t0 = time.perf_counter()
loss, t1 = step()
t2 = time.perf_counter()
print(f"B={B}: forward(train) {1e3 * (t1 - t0):.1f} ms, backward {1e3 * (t2 - t1):.1f} ms, total {1e3 * (t2 - t0):.1f} ms, "
      f"loss {float(loss.detach()):.4f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
fl = 3 * 316.37e9 * B
print(f"  = {fl / (t2 - t0) / 1e12:.0f} TFLOP/s on 3x the forward FLOPs ({1.0 / (t2 - t0):.2f} train-steps/s)")
if prof:
    tot = 0.0
    for name, evs in sorted(acc.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
        ms = sum(a.elapsed_time(b) for a, b in evs)
        tot += ms
        print(f"  {name:18s} {ms:8.2f} ms  ({len(evs)} launches)")
    print(f"  sum of kernels     {tot:8.2f} ms")
