"""CPU oracle of the NaturalSpeech2 denoiser hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of `Model.forward` of lucidrains/naturalspeech2-pytorch @ 659bec7
(`naturalspeech2_pytorch/naturalspeech2_pytorch.py`, "ns2.py" below; `attend.py`).  Every function cites the
reference lines it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product (`naturalspeech2_pytorch_b200`) never does.

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned against
outputs of the reference itself, generated in the authoring container by `tests/golden/make_golden.py` (which
imports /root/reference) and committed under `tests/golden/*.npz`; `tests/test_oracle_cpu.py` checks the oracle
against every one of them.

Parameters are passed as a dict {reference state_dict key: numpy array}.  `dtype` selects the arithmetic
(np.float64 for the ground truth, np.float32 to mirror the reference's default precision).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
from scipy.special import erf  # exact-erf GELU (F.gelu default), ns2.py:1007

Params = Dict[str, np.ndarray]


class ModelConfig:
    """Constructor arguments of Model (ns2.py:814-831)."""

    def __init__(self, dim, depth, dim_head=64, heads=8, ff_mult=4, wavenet_layers=8, wavenet_stacks=4,
                 dim_cond_mult=4, dim_prompt=None, num_latents_m=32, resampler_depth=2,
                 condition_on_prompt=False):
        self.dim, self.depth, self.dim_head, self.heads = dim, depth, dim_head, heads
        self.ff_mult, self.wavenet_layers, self.wavenet_stacks = ff_mult, wavenet_layers, wavenet_stacks
        self.dim_cond_mult, self.dim_prompt, self.num_latents_m = dim_cond_mult, dim_prompt, num_latents_m
        self.resampler_depth, self.condition_on_prompt = resampler_depth, condition_on_prompt

    def to_dict(self):
        return dict(self.__dict__)


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def linear(x, w, b=None):
    """nn.Linear: x @ W^T + b."""
    y = x @ w.T
    return y if b is None else y + b


def silu(x):
    return x / (1.0 + np.exp(-x))


def gelu(x):
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def sinusoidal_pos_emb(times, weights):
    """LearnedSinusoidalPosEmb.forward, ns2.py:115-120: [t, sin(2 pi t w), cos(2 pi t w)]."""
    x = times[:, None]
    freqs = x * weights[None, :] * 2 * math.pi
    return np.concatenate((x, np.sin(freqs), np.cos(freqs)), axis=-1)


def causal_conv1d(x, w, b, dilation=1):
    """CausalConv1d.forward, ns2.py:593-595, on channel-first x (B, C, N): left-pad dilation*(k-1) zeros."""
    O, I, K = w.shape
    pad = dilation * (K - 1)
    B, C, N = x.shape
    xp = np.concatenate((np.zeros((B, C, pad), dtype=x.dtype), x), axis=-1)
    y = np.zeros((B, O, N), dtype=x.dtype)
    for t in range(K):
        # tap t reads x_padded[n + t*dilation] = x[n - (K-1-t)*dilation]
        y += np.matmul(w[:, :, t], xp[:, :, t * dilation:t * dilation + N])  # (O,I) @ (B,I,N) -> (B,O,N)
    return y + b[None, :, None]


def rmsnorm(x, gamma=None, film: Optional[np.ndarray] = None):
    """RMSNorm.forward, ns2.py:736-746.  F.normalize(x, dim=-1) = x / max(||x||_2, 1e-12).
    film: (B, 2*dim) = to_gamma_beta(cond); gamma_t = first half, beta_t = second half."""
    dim = x.shape[-1]
    nrm = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), 1e-12)
    out = x / nrm * (dim ** 0.5)
    if gamma is not None:
        out = out * gamma
    if film is None:
        return out
    g, b = film[:, None, :dim], film[:, None, dim:]
    return out * g + b


def attend(q, k, v):
    """Attend.forward, attend.py:112-155 with mask=None, causal=False, dropout=0: softmax(q k^T / sqrt(d)) v."""
    scale = q.shape[-1] ** -0.5
    sim = np.matmul(q, k.transpose(0, 1, 3, 2)) * scale
    sim = sim - sim.max(axis=-1, keepdims=True)
    p = np.exp(sim)
    p = p / p.sum(axis=-1, keepdims=True)
    return np.matmul(p, v)


def attention(P: Params, prefix: str, x, heads, context=None, include_queries=False):
    """Attention.forward, ns2.py:1055-1069."""
    has_context = context is not None
    ctx = context if has_context else x
    if has_context and include_queries:
        ctx = np.concatenate((x, ctx), axis=-2)
    q = linear(x, P[prefix + "to_q.weight"])
    kv = linear(ctx, P[prefix + "to_kv.weight"])
    k, v = np.split(kv, 2, axis=-1)

    def split(t):
        b, n, hd = t.shape
        return t.reshape(b, n, heads, hd // heads).transpose(0, 2, 1, 3)

    o = attend(split(q), split(k), split(v))
    b, h, n, d = o.shape
    o = o.transpose(0, 2, 1, 3).reshape(b, n, h * d)
    return linear(o, P[prefix + "to_out.weight"])


def feedforward(P: Params, prefix: str, x, causal_conv: bool):
    """FeedForward, ns2.py:1009-1025: Linear -> GEGLU (first half value, second half gate) -> [causal conv k=3]
    -> Linear."""
    h = linear(x, P[prefix + "0.weight"], P[prefix + "0.bias"])
    val, gate = np.split(h, 2, axis=-1)
    h = gelu(gate) * val
    if causal_conv:
        hc = causal_conv1d(h.transpose(0, 2, 1), P[prefix + "2.1.weight"], P[prefix + "2.1.bias"])
        h = hc.transpose(0, 2, 1)
        last = "3."
    else:
        last = "2."
    return linear(h, P[prefix + last + "weight"], P[prefix + last + "bias"])


def wavenet_block(P: Params, prefix: str, x, t, dilation, has_skip):
    """WavenetResBlock.forward, ns2.py:619-642 (x channel-first)."""
    tt = linear(t, P[prefix + "to_time_cond.weight"], P[prefix + "to_time_cond.bias"])
    dim = x.shape[1]
    t_gamma, t_beta = tt[:, :dim, None], tt[:, dim:, None]
    res = causal_conv1d(x, P[prefix + "res_conv.weight"], P[prefix + "res_conv.bias"])
    y = causal_conv1d(x, P[prefix + "conv.weight"], P[prefix + "conv.bias"], dilation)
    y = y * t_gamma + t_beta
    y = np.tanh(y) * (1.0 / (1.0 + np.exp(-y)))
    y = y + res
    skip = None
    if has_skip:
        skip = causal_conv1d(y, P[prefix + "skip_conv.weight"], P[prefix + "skip_conv.bias"])
    return y, skip


def wavenet(P: Params, cfg: ModelConfig, x, t):
    """Wavenet.forward + WavenetStack.forward, ns2.py:672-688, 718-725."""
    x = causal_conv1d(x, P["wavenet.init_conv.weight"], P["wavenet.init_conv.bias"])
    inputs = [x] * cfg.wavenet_layers
    skips = None
    for s in range(cfg.wavenet_stacks):
        has_skip = s == cfg.wavenet_stacks - 1
        residuals, skips = [], []
        for i in range(cfg.wavenet_layers):
            r, sk = wavenet_block(P, f"wavenet.stacks.{s}.blocks.{i}.", inputs[i], t, 2 ** i, has_skip)
            residuals.append(r)
            skips.append(sk)
        inputs = residuals
    total = skips[0]
    for sk in skips[1:]:
        total = total + sk
    return causal_conv1d(total, P["wavenet.final_conv.weight"], P["wavenet.final_conv.bias"])


def perceiver_resampler(P: Params, cfg: ModelConfig, prompt):
    """PerceiverResampler.forward, ns2.py:568-579 (no pre-norm; keys include the latents)."""
    pre = "perceiver_resampler."
    x = prompt
    if pre + "proj_context.weight" in P:
        x = linear(x, P[pre + "proj_context.weight"], P[pre + "proj_context.bias"])
    B = x.shape[0]
    lat = np.broadcast_to(P[pre + "latents"][None], (B,) + P[pre + "latents"].shape).copy()
    for i in range(cfg.resampler_depth):
        lat = attention(P, f"{pre}layers.{i}.0.", lat, cfg.heads, context=x, include_queries=True) + lat
        lat = feedforward(P, f"{pre}layers.{i}.1.", lat, causal_conv=False) + lat
    return rmsnorm(lat, gamma=P[pre + "norm.gamma"])


def transformer(P: Params, cfg: ModelConfig, x, t, context=None):
    """ConditionableTransformer.forward, ns2.py:786-809."""
    for l in range(cfg.depth):
        pre = f"transformer.layers.{l}."
        film = linear(t, P[pre + "0.to_gamma_beta.weight"], P[pre + "0.to_gamma_beta.bias"])
        x = attention(P, pre + "1.", rmsnorm(x, film=film), cfg.heads) + x
        if cfg.condition_on_prompt:
            film = linear(t, P[pre + "2.to_gamma_beta.weight"], P[pre + "2.to_gamma_beta.bias"])
            x = attention(P, pre + "3.", rmsnorm(x, film=film), cfg.heads, context=context) + x
        film = linear(t, P[pre + "4.to_gamma_beta.weight"], P[pre + "4.to_gamma_beta.bias"])
        x = feedforward(P, pre + "5.", rmsnorm(x, film=film), causal_conv=True) + x
    x = rmsnorm(x, gamma=P["transformer.to_pred.0.gamma"])
    return linear(x, P["transformer.to_pred.1.weight"])


def model_forward(P: Params, cfg: ModelConfig, x, times, prompt=None, cond=None, drop_prompt=None,
                  drop_cond=None, dtype=np.float64):
    """Model.forward, ns2.py:929-1000.  The two CFG drop masks (ns2.py:950, 980) are explicit boolean arrays
    (B,) here — None means "keep everything" (cond_drop_prob = 0)."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in P.items()}
    x = np.asarray(x, dtype=dtype)
    times = np.asarray(times, dtype=dtype)
    B, N, D = x.shape
    t = sinusoidal_pos_emb(times, P["to_time_cond.0.weights"])
    t = silu(linear(t, P["to_time_cond.1.weight"], P["to_time_cond.1.bias"]))
    c = None
    if cfg.condition_on_prompt:
        prompt = np.asarray(prompt, dtype=dtype)
        cond = np.asarray(cond, dtype=dtype)
        dp = np.zeros(B, dtype=bool) if drop_prompt is None else np.asarray(drop_prompt, dtype=bool)
        dc = np.zeros(B, dtype=bool) if drop_cond is None else np.asarray(drop_cond, dtype=bool)
        pc = silu(linear(prompt.mean(axis=1), P["to_prompt_cond.1.weight"], P["to_prompt_cond.1.bias"]))
        pc = np.where(dp[:, None], P["null_prompt_cond"][None], pc)
        t = np.concatenate((t, pc), axis=-1)
        tokens = perceiver_resampler(P, cfg, prompt)
        c = np.where(dp[:, None, None], P["null_prompt_tokens"][None], tokens)
    xc = x.transpose(0, 2, 1)  # 'b n d -> b d n', ns2.py:972
    if cfg.condition_on_prompt:
        cp = causal_conv1d(cond, P["cond_to_model_dim.weight"], P["cond_to_model_dim.bias"])  # k=1, no padding
        cp = np.where(dc[:, None, None], P["null_cond"][None], cp)
        L = cp.shape[-1]
        if L > N:  # pad_or_curtail_to_length, ns2.py:70-77
            cp = cp[..., :N]
        elif L < N:
            cp = np.concatenate((cp, np.zeros((B, D, N - L), dtype=dtype)), axis=-1)
        xc = xc + cp
    xc = wavenet(P, cfg, xc, t)
    x = xc.transpose(0, 2, 1)
    return transformer(P, cfg, x, t, context=c)
