"""Shared test helpers: golden loading, model construction with deterministic weights."""
from __future__ import annotations

import ast
from pathlib import Path

import numpy as np
import torch

from param_fill import fill_module

GOLDEN = Path(__file__).resolve().parent / "golden"
MODEL_CASES = ["uncond_small", "cond_small", "cond_samedim", "readme_uncond"]


def load_model_golden(name: str):
    z = np.load(GOLDEN / f"model_{name}.npz")
    kwargs = dict(ast.literal_eval(str(z["config"])))
    return z, kwargs, int(z["fill_seed"])


def build_model(kwargs: dict, seed: int, device="cpu"):
    """naturalspeech2_pytorch_b200.Model with the deterministic weights the goldens were generated with."""
    from naturalspeech2_pytorch_b200 import Model
    m = Model(**kwargs)
    fill_module(m, seed)
    return m.to(device).eval()


def numpy_params(model: torch.nn.Module):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def oracle_config(kwargs: dict):
    from oracle.denoiser_oracle import ModelConfig
    allowed = ModelConfig.__init__.__code__.co_varnames
    return ModelConfig(**{k: v for k, v in kwargs.items() if k in allowed})


def err_stats(got: np.ndarray, ref: np.ndarray):
    d = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref, dtype=np.float64))
    return float(d.max()), float(np.sqrt((d ** 2).mean()))
