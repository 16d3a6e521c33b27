#!/usr/bin/env python
"""Per-stage clock64 timeline of CTA 0 of the RVQ encode kernel (stats[4 + q*8 + slot]) + throughput at 1 M frames."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
Q, K, D = 8, 1024, 128   # configs[3] of BASELINE.json, same synthetic data as bench.py secondary_rvq
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cb = torch.randn(Q, K, D, generator=torch.Generator().manual_seed(1234)).to(dev)
frames = torch.randn(F, D, generator=torch.Generator().manual_seed(1235)).to(dev)
prep = ops.rvq_prepare(cb)
stats = torch.zeros(_lib.NS2_RVQ_STATS_LEN, device=dev, dtype=torch.int64)
codes = ops.rvq_encode(frames, cb, prep, stats=stats)
torch.cuda.synchronize()
s = stats.cpu().tolist()
print(f"lookups {s[0]}  re-scored {s[1]} ({100.0 * s[1] / max(s[0], 1):.2f} %)  full scans {s[2]}  block scans {s[3]}")
names = ["B1", "A ready", "scan end", "B2", "own rows", "stage end", "B3", "-"]
print("q   " + " ".join(f"{n:>10s}" for n in names))
t0 = s[4]
for q in range(Q):
    row = s[4 + q * 8: 4 + q * 8 + 8]
    print(f"{q:<3d} " + " ".join(f"{v - t0:>10d}" for v in row))
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.rvq_encode(frames, cb, prep, codes=codes)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"encode {F} frames x {Q} quantisers: {ms:.3f} ms  {F * Q / ms / 1e3:.0f} Mcodes/s")
