#!/usr/bin/env python
"""Bring-up aid: per-tile clock64 timeline of CTA pair 0 of the pair GEMM (NS2_GEMM_DEBUG=8)."""
import ctypes, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["NS2_GEMM_DEBUG"] = os.environ.get("NS2_GEMM_DEBUG", "8")
from naturalspeech2_pytorch_b200 import ops, _lib  # noqa: E402
B, N = 32, 1024
n, k = int(sys.argv[1]), int(sys.argv[2])
a = (torch.randn(B, N, k, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(n, k, device="cuda") * 0.02).to(torch.bfloat16)
out = torch.empty(B, N, n, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(a, w, out, n=n, epilogue=ops.EPI_BF16)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (16 * 64))()
lib.ns2_debug_gemm_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.ns2_debug_gemm_timeline(buf, 16 * 64) == 0
t0 = buf[0]
names = ["mma:wait_tempty", "mma:got_tempty", "mma:first_full", "mma:committed", "tma:first_issue", "tma:last_issue",
         "epi:wait_tfull", "epi:got_tfull", "epi:arrived"]
print(f"N={n} K={k}; cycles relative to tile 0 start")
for ti in range(12):
    print(ti, " ".join(f"{names[s].split(':')[1][:10]}={buf[ti*16+s]-t0:7d}" for s in range(9)))
