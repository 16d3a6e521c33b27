"""CPU oracle of the diffusion wrapper arithmetic around the denoiser — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Numpy restatement of the element-wise maths of `NaturalSpeech2.forward` (ns2.py:1613-1671) and
`NaturalSpeech2.ddim_sample` (ns2.py:1379-1431) of lucidrains/naturalspeech2-pytorch @ 659bec7, with the
denoiser passed in as a callable.  Pinned by tests/golden/diffusion_*.npz (generated from the reference by
tests/golden/make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def sigmoid_schedule(t, start=-3, end=3, tau=1, clamp_min=1e-9):
    """ns2.py:1144-1148."""
    v_start = _sigmoid(np.asarray(start / tau, dtype=t.dtype))
    v_end = _sigmoid(np.asarray(end / tau, dtype=t.dtype))
    gamma = (-_sigmoid((t * (end - start) + start) / tau) + v_end) / (v_end - v_start)
    return np.clip(gamma, clamp_min, 1.0)


def gamma_to_alpha_sigma(gamma, scale=1.0):
    """ns2.py:1152-1153."""
    return np.sqrt(gamma) * scale, np.sqrt(1 - gamma)


def sampling_time_pairs(timesteps, dtype=np.float32):
    """get_sampling_timesteps, ns2.py:1303-1308: consecutive pairs of linspace(1, 0, timesteps + 1)."""
    times = np.linspace(1.0, 0.0, timesteps + 1, dtype=dtype)
    return list(zip(times[:-1], times[1:]))


def training_loss(model_fn, x_start, times, noise, objective="v", min_snr_gamma=5.0, scale=1.0):
    """ns2.py:1621-1666 with `times` and `noise` given (the reference draws them at 1621 and 1625).
    model_fn(noised, times) -> prediction.  Returns (loss scalar, dict of intermediates)."""
    dtype = x_start.dtype
    gamma = sigmoid_schedule(times.astype(dtype))
    alpha, sigma = gamma_to_alpha_sigma(gamma[:, None, None], scale)
    noised = alpha * x_start + sigma * noise
    pred = model_fn(noised, times)
    if objective == "v":
        target = alpha * noise - sigma * x_start
    elif objective == "eps":
        target = noise
    else:
        target = x_start
    per_sample = ((pred - target) ** 2).reshape(pred.shape[0], -1).mean(axis=1)
    snr = (alpha * alpha) / (sigma * sigma)
    clipped = np.minimum(snr, min_snr_gamma)
    if objective == "v":
        w = clipped / (snr + 1)
    elif objective == "eps":
        w = clipped / snr
    else:
        w = clipped
    # loss is (B,), loss_weight is (B,1,1): the reference broadcasts them to (B,1,B) before .mean() (ns2.py:1666)
    loss = (per_sample * w).mean()
    return loss, {"noised": noised, "target": target, "per_sample": per_sample, "weight": w}


def ddim_step(x, v, t, t_next, scale=1.0, objective="v"):
    """One iteration of ddim_sample, ns2.py:1396-1429 (time_difference = 0); `v` is the model output."""
    dtype = x.dtype
    g = sigmoid_schedule(np.asarray(t, dtype=dtype))
    gn = sigmoid_schedule(np.asarray(t_next, dtype=dtype))
    a, s = gamma_to_alpha_sigma(g[:, None, None], scale)
    an, sn = gamma_to_alpha_sigma(gn[:, None, None], scale)
    if objective == "v":          # ns2.py:1412-1421
        x0 = a * x - s * v
    elif objective == "eps":
        x0 = (x - s * v) / np.maximum(a, 1e-10)
    else:
        x0 = v
    eps = (x - a * x0) / np.maximum(s, 1e-10)
    return x0 * an + eps * sn


def ddim_sample(model_fn, x_init, timesteps, scale=1.0, objective="v"):
    """ddim_sample, ns2.py:1379-1431, from a given initial noise."""
    x = x_init
    B = x.shape[0]
    for t, tn in sampling_time_pairs(timesteps, dtype=x.dtype):
        tb = np.full((B,), t, dtype=x.dtype)
        tnb = np.full((B,), tn, dtype=x.dtype)
        v = model_fn(x, tb)
        x = ddim_step(x, v, tb, tnb, scale, objective)
    return x
