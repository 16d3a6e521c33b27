#!/usr/bin/env python
"""configs[4] data-parallel training step under torchrun with an SM budget for the ns2 kernels (NS2_SM_LIMIT) and a CTA
budget for NCCL (NCCL_MAX_CTAS, read by NCCL at communicator creation): how much of the gradient all-reduce hides
under the backward without stalling the persistent GEMM grids."""
import datetime
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from naturalspeech2_pytorch_b200 import ops  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
lim = int(os.environ.get("NS2_SM_LIMIT", "0"))
ops.set_sm_limit(lim)
res = bench.train_step_dp(dev, world, bench._peaks()[0], steps=5, warmup=2)
if rank == 0:
    print(json.dumps({"NS2_SM_LIMIT": lim, "NCCL_MAX_CTAS": os.environ.get("NCCL_MAX_CTAS"), "ms_per_step": res["ms_per_step"],
                      "value": res["value"]}))
dist.destroy_process_group()
