"""CPU oracle of the aligner's monotonic alignment search — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of `maximum_path(value, mask, const=None)`, naturalspeech2_pytorch/aligner.py:88-122 (called from
`Aligner.forward`, aligner.py:214).  Every floating-point operation is the reference's, in fp32 and in the same order
(one multiply by the mask, one add per cell, `>=` compares with ties resolved to "stay"), so the result is
bit-identical to the reference; pinned by tests/golden/aligner_mas.npz, which tests/golden/make_golden.py generates
by running the reference function itself.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module.
"""
from __future__ import annotations

import numpy as np


def maximum_path(value: np.ndarray, mask: np.ndarray, const: float = -np.inf, return_index: bool = False):
    """value, mask: (b, t_x, t_y).  Returns the 0/1 path (b, t_x, t_y) fp32 [and idx (b, t_y) int32]."""
    value = np.asarray(value, dtype=np.float32)
    maskf = np.asarray(mask, dtype=np.float32)
    const = np.float32(const)
    value = value * maskf                                            # aligner.py:93
    b, t_x, t_y = value.shape
    direction = np.zeros((b, t_x, t_y), dtype=np.int64)              # aligner.py:96
    v = np.zeros((b, t_x), dtype=np.float32)                         # aligner.py:97
    x_range = np.arange(t_x, dtype=np.float32)[None, :]              # aligner.py:98
    for j in range(t_y):                                             # aligner.py:100-108
        v0 = np.concatenate([np.full((b, 1), const, dtype=np.float32), v[:, :-1]], axis=1)
        v1 = v
        max_mask = v1 >= v0
        v_max = np.where(max_mask, v1, v0)
        direction[:, :, j] = max_mask
        index_mask = x_range <= j
        with np.errstate(invalid="ignore"):
            v = np.where(index_mask, (v_max + value[:, :, j]).astype(np.float32), const).astype(np.float32)
    direction = np.where(maskf != 0, direction, 1)                   # aligner.py:110
    path = np.zeros((b, t_x, t_y), dtype=np.float32)                 # aligner.py:112
    index = maskf[:, :, 0].sum(1).astype(np.int64) - 1               # aligner.py:113
    rng = np.arange(b)
    idx_out = np.zeros((b, t_y), dtype=np.int32)
    for j in reversed(range(t_y)):                                   # aligner.py:116-118
        path[rng, index, j] = 1
        idx_out[:, j] = np.where(index < 0, index + t_x, index)
        index = index + direction[rng, index, j] - 1
    path = path * maskf                                              # aligner.py:120
    return (path, idx_out) if return_index else path
