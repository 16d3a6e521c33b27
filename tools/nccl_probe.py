#!/usr/bin/env python
"""Step-by-step NCCL bring-up probe (multi-GPU box): prints progress with flush so a hang can be located."""
import datetime, os, sys, time
import torch, torch.distributed as dist
r, w, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print(f"[rank {r} +{time.time() - T0:6.1f}s]", *a, flush=True)
T0 = time.time()
log("start; visible devices:", torch.cuda.device_count())
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
mode = sys.argv[1] if len(sys.argv) > 1 else "device_id"
kw = {"device_id": dev} if mode == "device_id" else {}
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=60), **kw)
log("init_process_group done", mode)
t = torch.ones(1, device=dev) * (r + 1)
dist.all_reduce(t)
torch.cuda.synchronize()
log("all_reduce ->", float(t))
dist.barrier(device_ids=[lr])
torch.cuda.synchronize()
log("barrier done")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from naturalspeech2_pytorch_b200 import Model
m = Model(dim=128, depth=1, heads=1, wavenet_layers=2, wavenet_stacks=1).to(dev).eval()
x = torch.randn(2, 256, 128, device=dev)
out = m(x, torch.rand(2, device=dev))
loss = out.float().pow(2).mean()
dist.all_reduce(loss)
torch.cuda.synchronize()
log("model forward + loss all_reduce ->", float(loss))
dist.destroy_process_group()
log("done")
