#!/usr/bin/env python
"""Bring-up probe for Model.use_cuda_graphs: prints progress, dumps the Python stack if it stalls."""
import faulthandler, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import Model
faulthandler.dump_traceback_later(40, exit=True)
T0 = time.time()
def log(*a): print(f"[+{time.time()-T0:5.1f}s]", *a, flush=True)
kw = dict(dim=128, depth=2, heads=2, wavenet_layers=3, wavenet_stacks=2) if "small" in sys.argv else dict(dim=128, depth=6)
torch.manual_seed(0)
m = Model(**kw).cuda().eval()
x = torch.randn(4, 1024, 128, device="cuda"); t = torch.rand(4, device="cuda")
e = m(x, t).clone(); torch.cuda.synchronize(); log("eager ok")
m.use_cuda_graphs = True
g = m(x, t); log("graph call returned (capture + first replay enqueued)")
torch.cuda.synchronize(); log("first replay synced; equal:", torch.equal(e, g))
for i in range(5):
    g = m(x, t)
torch.cuda.synchronize(); log("5 replays synced; equal:", torch.equal(e, g))
if "ddim" in sys.argv:
    from naturalspeech2_pytorch_b200 import NaturalSpeech2
    ns = NaturalSpeech2(m, target_sample_hz=24000, timesteps=4)
    out = ns.sample(length=256, batch_size=2); torch.cuda.synchronize(); log("ddim sample ok", tuple(out.shape))
    out = ns.sample(length=256, batch_size=2); torch.cuda.synchronize(); log("ddim sample 2 ok")
