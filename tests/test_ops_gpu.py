"""GPU: every kernel behind the C ABI against a plain fp32 PyTorch formulation of the same op with the same
bf16-rounded operands (tools/gpu_check.py holds the cases; tolerances are written there per op)."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", ["elementwise", "gemm_plain", "gemm_conv", "gemm_fused", "attn", "rvq"])
def test_kernel_group(group):
    import torch
    import gpu_check
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    assert getattr(gpu_check, "run_" + group)(), f"kernel group {group} failed (see captured stdout)"
    torch.cuda.synchronize()
