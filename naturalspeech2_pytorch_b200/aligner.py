"""Monotonic alignment search of the reference's aligner (naturalspeech2_pytorch/aligner.py:88-122) on the GPU.

`maximum_path(value, mask, const=None)` keeps the reference signature and return value (the dense 0/1 path, same dtype
as `value`) but runs as two kernels of libns2b200.so (csrc/align.cu) instead of ~10 PyTorch launches per mel frame plus
a Python backtrack loop.  `patch_reference_aligner()` rebinds the reference module's function, so `Aligner.forward`
(aligner.py:199-217, called from NaturalSpeech2.forward ns2.py:1578) uses it unchanged.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def maximum_path(value: torch.Tensor, mask: torch.Tensor, const=None) -> torch.Tensor:
    """value: (b, t_x, t_y) alignment scores, mask: (b, t_x, t_y) 0/1.  Returns the hard monotonic path (b, t_x, t_y)
    in value's dtype, bit-identical to the reference.  There is no CPU path: both tensors must live on the GPU."""
    if not value.is_cuda:
        raise ValueError("maximum_path: value must be a CUDA tensor (the ns2_b200 ops have no CPU path)")
    if value.dtype != torch.float32:
        # the reference multiplies value*mask in value's dtype and then accumulates in the promoted type
        # (aligner.py:93-108); only the fp32 case (what Aligner.forward produces) is restated here
        raise NotImplementedError(f"maximum_path: value must be float32, got {value.dtype}")
    neg = float("-inf") if const is None else float(const)
    maskf = mask.to(device=value.device, dtype=torch.float32).expand_as(value).contiguous()
    _, path = ops.maximum_path(value.contiguous(), maskf, neg)
    return path


def alignment_indices(value: torch.Tensor, mask: torch.Tensor, const=None) -> torch.Tensor:
    """(b, t_y) int32: the text position aligned to every mel frame (the argmax over t_x of the path), without
    materialising the dense path."""
    neg = float("-inf") if const is None else float(const)
    maskf = mask.to(device=value.device, dtype=torch.float32).expand_as(value).contiguous()
    idx, _ = ops.maximum_path(value.contiguous(), maskf, neg, want_path=False)
    return idx


def patch_reference_aligner(ref_aligner_module=None) -> None:
    """Rebind `naturalspeech2_pytorch.aligner.maximum_path` (looked up as a module global by Aligner.forward,
    aligner.py:214) to the CUDA implementation."""
    if ref_aligner_module is None:
        import naturalspeech2_pytorch.aligner as ref_aligner_module  # noqa: PLC0415
    ref_aligner_module.maximum_path = maximum_path
