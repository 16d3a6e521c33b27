#!/usr/bin/env python
"""Bring-up aid: per-tile clock64 timeline of CTA pair 0 of the pair GEMM (ns2_gemm_args.debug_timeline).
usage: python tools/gemm_timeline.py qkv|ffin|ffout|conv|wavenet"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import ops  # noqa: E402

B, N, D, Dp = 32, 1024, 512, 1408
dev, bf = "cuda", torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
torch.manual_seed(0)
tl = torch.zeros(64 * 8, device=dev, dtype=torch.int64)
if which == "qkv":
    a = torch.randn(B, N, D, device=dev).to(bf); w = (torch.randn(1536, D, device=dev) * 0.04).to(bf)
    out = torch.empty(B, N, 1536, device=dev, dtype=bf)
    run = lambda **k: ops.gemm(a, w, out, n=1536, epilogue=ops.EPI_BF16, **k)
elif which == "ffin":
    a = torch.randn(B, N, D, device=dev).to(bf); w = (torch.randn(2 * Dp, D, device=dev) * 0.04).to(bf)
    b1 = torch.randn(2 * Dp, device=dev); out = torch.empty(B, N, Dp, device=dev, dtype=bf)
    run = lambda **k: ops.gemm(a, w, out, n=2 * Dp, epilogue=ops.EPI_GEGLU, bias=b1, **k)
elif which == "ffout":
    a = torch.randn(B, N, Dp, device=dev).to(bf); w = (torch.randn(D, Dp, device=dev) * 0.03).to(bf)
    b2 = torch.randn(D, device=dev); xr = torch.randn(B, N, D, device=dev)
    run = lambda **k: ops.gemm(a, w, xr, n=D, epilogue=ops.EPI_F32, bias=b2, resid=xr, **k)
elif which == "conv":
    a = (torch.randn(B, N, Dp, device=dev) * 0.5).to(bf); w = (torch.randn(Dp, 3 * Dp, device=dev) * 0.02).to(bf)
    bc = torch.randn(Dp, device=dev); out = torch.empty(B, N, Dp, device=dev, dtype=bf)
    run = lambda **k: ops.gemm(a, w, out, n=Dp, epilogue=ops.EPI_BF16, bias=bc, segs=ops.conv3_segs(Dp), **k)
else:
    G = 8
    a = (torch.randn(B, N, G * D, device=dev) * 0.5).to(bf); w = (torch.randn(G * D, 4 * D, device=dev) * 0.02).to(bf)
    bias = torch.randn(2 * G * D, device=dev); film = torch.randn(B, G * 2 * D, device=dev)
    out = torch.empty(B, N, G * D, device=dev, dtype=bf)
    segs = ops.conv3_segs(D) + [(0, 3 * D, D, 0, 1)]
    run = lambda **k: ops.gemm(a, w, out, n=D, epilogue=ops.EPI_WAVENET, bias=bias, bias1_off=G * D, segs=segs, film=film,
                               film_group_stride=2 * D, groups=G, a_group_col_stride=D, b_group_row_stride=D,
                               out_group_col_stride=D, dil=[2 ** i for i in range(G)], **k)
for _ in range(2):
    run()
run(debug_timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().view(64, 8)
t0 = int(t[0, 0])
names = ["mma:wait", "mma:go", "mma:commit", "e0:wait", "e0:go", "e0:done", "e7:wait", "e7:done"]
print(which, "- cycles relative to the first stamp; rows = tile index of CTA pair 0")
print("ti  " + " ".join(f"{n:>10s}" for n in names))
for ti in range(14):
    print(f"{ti:<3d} " + " ".join(f"{int(t[ti, s]) - t0:10d}" for s in range(8)))
