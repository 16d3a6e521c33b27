"""Deterministic, platform-independent parameter fill shared by the golden generator and the tests.

Both the reference `Model` and `naturalspeech2_pytorch_b200.Model` expose the same state_dict keys, so filling
by key gives both the same weights without shipping them: every tensor is drawn from a CPU torch.Generator
seeded by (seed, crc32(key)).  Scales are chosen so activations stay O(1) through the network and so that biases,
gammas and the `null_*` parameters are non-trivial (the default init has zeros / ones there).
"""
from __future__ import annotations

import zlib

import torch


def fill_tensor(key: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if key.endswith("gamma"):
        return 1.0 + 0.1 * t
    if key.endswith("to_gamma_beta.weight") or key.endswith("to_time_cond.weight") and "stacks" in key:
        # FiLM projections: keep gamma near 1 via the bias, small data-dependent part
        fan_in = shape[-1]
        return t * (0.5 / fan_in ** 0.5)
    if key.endswith("to_gamma_beta.bias") or (key.endswith("to_time_cond.bias") and "stacks" in key):
        half = shape[0] // 2
        out = 0.1 * t
        out[:half] += 1.0  # gamma half
        return out
    if key.endswith("bias"):
        return 0.1 * t
    if key.endswith("weights"):  # sinusoidal frequencies
        return t
    if key in ("null_prompt_cond", "null_prompt_tokens", "null_cond", "perceiver_resampler.latents"):
        return 0.5 * t
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return t * (1.0 / fan_in ** 0.5)
    return t


@torch.no_grad()
def fill_module(module: torch.nn.Module, seed: int) -> None:
    for key, p in module.state_dict().items():
        p.copy_(fill_tensor(key, p.shape, seed).to(p.dtype))


def seeded(shape, seed: int, scale: float = 1.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32) * scale


def seeded_uniform(shape, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand(tuple(shape), generator=g, dtype=torch.float32)


def rvq_fixture_inputs(F: int = 2048, Q: int = 8, K: int = 1024, d: int = 128):
    """Seeded codebooks (with one duplicated entry) and two frame sets for the RVQ goldens:
    'random' ~ N(0,1) (plus four exact codeword sums) and 'realistic' = sum of codewords with halving scale
    + N(0, 0.05^2), so that residual norms shrink stage by stage as they do for real Encodec latents."""
    cb = seeded((Q, K, d), 1234)
    cb[0, 7] = cb[0, 3]  # duplicate entry: ties -> lowest index
    frames = seeded((F, d), 1235)
    frames[:4] = cb[0, 7] + cb[1, 11]
    idx = torch.randint(0, K, (F, Q), generator=torch.Generator().manual_seed(5))
    real = sum(cb[q][idx[:, q]] * (0.5 ** q) for q in range(Q)) + 0.05 * seeded((F, d), 1236)
    return cb, {"random": frames, "realistic": real}
