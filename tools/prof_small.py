#!/usr/bin/env python
"""Small launches of the attention and RVQ kernels for `ncu --set full --import-source on` captures."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import ops  # noqa: E402
dev = "cuda"
torch.manual_seed(0)
B, N, H = 8, 1024, 8
qkv = torch.randn(B, N, 3 * H * 64, device=dev).to(torch.bfloat16)
o = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv[:, :, :512], qkv[:, :, 512:1024], qkv[:, :, 1024:], o, heads=H)
cb = torch.randn(8, 1024, 128, device=dev)
prep = ops.rvq_prepare(cb)
F = 148 * 128 * 2
x = torch.randn(F, 128, device=dev)
for _ in range(2):
    ops.rvq_encode(x, cb, prep)
torch.cuda.synchronize()
