#!/usr/bin/env python
"""Bring-up aid: clock64 timeline of CTA 0 of the two-tile attention kernel at the cfg2 shape."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from naturalspeech2_pytorch_b200 import ops  # noqa: E402

kern = int(sys.argv[1]) if len(sys.argv) > 1 else ops.ATTN_TWO_TILE_LOCKSTEP
B, H, N = 32, 8, 1024
inner = H * 64
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * inner, device="cuda").bfloat16()
out = torch.empty(B, N, inner, device="cuda", dtype=torch.bfloat16)
tl = torch.zeros(64 * 16, device="cuda", dtype=torch.int64)
args = (qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], out)
for _ in range(2):
    ops.attention(*args, heads=H, kernel=kern)
ops.attention(*args, heads=H, kernel=kern, debug_timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().view(64, 16)
t0 = int(t[0, 0])
names = ["w0:wait", "w0:S", "w0:ld", "w0:max", "w0:ofull", "w0:Pdone", "w1:wait", "w1:S", "w1:ld", "w1:max", "w1:ofull",
         "w1:Pdone", "mma:S0", "mma:S1", "mma:PV0", "mma:PV1"]
print("kernel", kern, "cycles relative to the first stamp; rows = global key-tile index g of CTA 0")
print("g   " + " ".join(f"{n:>9s}" for n in names))
for g in range(24):
    print(f"{g:<3d} " + " ".join(f"{int(t[g, s]) - t0:9d}" for s in range(16)))
