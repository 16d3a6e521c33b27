"""CPU oracle of the Encodec residual vector quantiser — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The RVQ arithmetic is NOT in /root/reference: `naturalspeech2_pytorch` reaches it through
`audiolm_pytorch.EncodecWrapper` (setup.py:24 `audiolm-pytorch>=0.30.2`, no lock file / pinned version), which
wraps Meta's `encodec` (`EncodecModel.encodec_model_24khz()`, `model.quantizer` = ResidualVectorQuantizer of
EuclideanCodebook).  Call sites in the reference: ns2.py:1445, 1611 (encode), 1496 (decode).
Published algorithm (encodec/quantization/core_vq.py, mirrored in the HF `transformers` Encodec port,
modeling_encodec.py:364-369 and 424-447):

    EuclideanCodebook.quantize:   dist = -(||x||^2 - 2 x E^T + ||E||^2);  idx = dist.max(-1).indices
    ResidualVectorQuantization:   for each layer: idx = quantize(residual); residual -= E[idx]; out += E[idx]

`torch.max(...).indices` returns the first maximum on ties.  The fp32 formula above is rounding-order
dependent on exact near-ties, so the oracle defines the index as the EXACT argmin of ||r - E_k||^2 (fp64 on the
fp32 operands, lowest index on ties) with the residual chain kept in fp32 exactly like the reference; that is
what the CUDA kernel guarantees.  Parity pin: tests/golden/rvq_*.npz holds codes produced by the `transformers`
Encodec quantiser (fp32 formula) on seeded synthetic codebooks; the oracle must agree with it on every code whose
fp64 top-2 gap is not below fp32 rounding noise (test_oracle_cpu.py reports the count; it is 0 on the fixtures).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def encode(frames: np.ndarray, codebooks: np.ndarray, return_gaps: bool = False, block: int = 4096):
    """frames (F, d) fp32, codebooks (Q, K, d) fp32 -> codes (F, Q) int64 [, relative top-2 gaps (F, Q)]."""
    frames = np.ascontiguousarray(frames, dtype=np.float32)
    codebooks = np.ascontiguousarray(codebooks, dtype=np.float32)
    F, d = frames.shape
    Q, K, _ = codebooks.shape
    codes = np.empty((F, Q), dtype=np.int64)
    gaps = np.empty((F, Q), dtype=np.float64) if return_gaps else None
    residual = frames.copy()
    for q in range(Q):
        E = codebooks[q].astype(np.float64)
        e2 = (E * E).sum(-1)
        for s in range(0, F, block):
            r = residual[s:s + block].astype(np.float64)
            # exact squared distance ||r||^2 - 2 r.E + ||E||^2 in fp64 (products of fp32 values are exact)
            dist = (r * r).sum(-1, keepdims=True) - 2.0 * (r @ E.T) + e2[None]
            idx = dist.argmin(axis=1)  # first minimum on ties
            codes[s:s + block, q] = idx
            if return_gaps:
                part = np.partition(dist, 1, axis=1)
                gaps[s:s + block, q] = (part[:, 1] - part[:, 0]) / np.maximum(part[:, 1], 1e-30)
            residual[s:s + block] -= codebooks[q][idx]  # fp32, the same op as the reference
    return (codes, gaps) if return_gaps else codes


def decode(codes: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    """sum of codeword lookups accumulated in layer order in fp32 (modeling_encodec.py:440-447)."""
    codebooks = np.asarray(codebooks, dtype=np.float32)
    out = np.zeros((codes.shape[0], codebooks.shape[-1]), dtype=np.float32)
    for q in range(codebooks.shape[0]):
        out = out + codebooks[q][codes[:, q]]
    return out


def encode_fp32_formula(frames: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    """The reference's literal fp32 formula (for the cpu_baseline timing and for cross-checks)."""
    frames = np.ascontiguousarray(frames, dtype=np.float32)
    codes = np.empty((frames.shape[0], codebooks.shape[0]), dtype=np.int64)
    residual = frames.copy()
    for q in range(codebooks.shape[0]):
        E = codebooks[q].astype(np.float32)
        dist = -((residual * residual).sum(1, keepdims=True) - 2 * (residual @ E.T) + (E * E).sum(1)[None])
        idx = dist.argmax(axis=1)
        codes[:, q] = idx
        residual -= E[idx]
    return codes
