#!/usr/bin/env python
"""Where the configs[4]-style training step spends its time: torch.profiler over one step (ns2 kernels vs ATen kernels
vs host gaps) and coarse synchronized timers around repacking / forward / backward / optimizer."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import Model, NaturalSpeech2  # noqa: E402

cond = "--cond" in sys.argv
B = 32
torch.manual_seed(0)
kw = dict(dim=512, depth=12, heads=8)
if cond:
    kw.update(dim_prompt=512, condition_on_prompt=True)
model = Model(**kw).cuda().train()
ns = NaturalSpeech2(model, target_sample_hz=24000)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
g = torch.Generator().manual_seed(1)
lat = torch.randn(B, 1024, 512, generator=g).cuda()
extra = {}
if cond:
    extra = dict(prompt_enc=torch.randn(B, 103, 512, generator=g).cuda(), cond=torch.randn(B, 512, 1024, generator=g).cuda())


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def step(timers=None):
    t0 = sync()
    opt.zero_grad(set_to_none=True)
    model.packed()
    t1 = sync()
    model.packed_transposed()
    t2 = sync()
    loss = ns(lat, **extra)
    t3 = sync()
    loss.backward()
    t4 = sync()
    opt.step()
    t5 = sync()
    if timers is not None:
        timers.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
    return loss


for _ in range(2):
    step()
T = []
for _ in range(3):
    step(T)
names = ("repack", "repack_T", "forward", "backward", "optimizer")
for i, n in enumerate(names):
    print(f"{n:10s} {1e3 * sum(t[i] for t in T) / len(T):8.2f} ms")
print(f"{'total':10s} {1e3 * sum(sum(t) for t in T) / len(T):8.2f} ms (synchronised phases)")

from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
