"""Timing of ns2_maximum_path (csrc/align.cu) at a few shapes; run on the GPU box."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from naturalspeech2_pytorch_b200 import ops

for b, t_x, t_y in [(32, 100, 1024), (32, 35, 200), (256, 100, 1024), (32, 256, 2048), (8, 1024, 4096)]:
    v = torch.rand(b, t_x, t_y, device="cuda")
    m = torch.ones(b, t_x, t_y, device="cuda")
    for want_path in (False, True):
        for _ in range(3):
            ops.maximum_path(v, m, want_path=want_path)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.maximum_path(v, m, want_path=want_path)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"b={b} t_x={t_x} t_y={t_y} path={want_path}: {ms*1e3:.1f} us  "
              f"({ms*1e6/t_y:.0f} ns/frame, {3*b*t_x*t_y*4/ms/1e6:.0f} GB/s algorithmic)")
