"""`NaturalSpeech2`: the latent-diffusion wrapper around the denoiser, with the call signatures of
naturalspeech2_pytorch.NaturalSpeech2 (ns2.py:1160-1684): `forward` (training loss), `sample`, `ddim_sample`.

Scope: the per-timestep path — schedules, q-sample, v-target, per-sample MSE, min-SNR
weight, the DDIM update and classifier-free guidance — runs on the sm_100a kernels (`ops.q_sample`,
`Model.forward`, `ops.mse_rows`, `ops.ddim_step`, `ops.cfg_combine`).  The once-per-sample conditioning encoders
of the reference (PhonemeEncoder, SpeechPromptEncoder, DurationPitchPredictor, Aligner; ns2.py:228-527,
aligner.py) are out of scope: for a conditional model pass their outputs directly (`prompt_enc=`, `cond=`) or
give a `conditioner` callable that produces them (e.g. the reference modules).

`forward` is forward-only in this round (no autograd graph; backward kernels are row f1 of SURVEY 8f).
"""
from __future__ import annotations

import math
from functools import partial
from typing import Callable, Optional

import torch
from torch import nn

from . import ops
from .model import Model


def _exists(v):
    return v is not None


# ---- noise schedules (ns2.py:1133-1156); tiny (B,)-sized host-side torch math ----
def simple_linear_schedule(t, clip_min=1e-9):
    return (1 - t).clamp(min=clip_min)


def cosine_schedule(t, start=0, end=1, tau=1, clip_min=1e-9):
    power = 2 * tau
    v_start = math.cos(start * math.pi / 2) ** power
    v_end = math.cos(end * math.pi / 2) ** power
    output = torch.cos((t * (end - start) + start) * math.pi / 2) ** power
    output = (v_end - output) / (v_end - v_start)
    return output.clamp(min=clip_min)


def sigmoid_schedule(t, start=-3, end=3, tau=1, clamp_min=1e-9):
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    gamma = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    return gamma.clamp_(min=clamp_min, max=1.)


def gamma_to_alpha_sigma(gamma, scale=1):
    return torch.sqrt(gamma) * scale, torch.sqrt(1 - gamma)


class NaturalSpeech2(nn.Module):
    def __init__(self, model: Model, codec=None, *, tokenizer=None, target_sample_hz=None, timesteps=1000,
                 use_ddim=True, noise_schedule="sigmoid", objective="v", schedule_kwargs: dict = dict(),
                 time_difference=0., min_snr_loss_weight=True, min_snr_gamma=5, train_prob_self_cond=0.9,
                 rvq_cross_entropy_loss_weight=0., scale=1.,
                 conditioner: Optional[Callable] = None, cuda_graphs: bool = True, **conditioning_kwargs):
        super().__init__()
        if not isinstance(model, Model):
            raise TypeError("model must be a naturalspeech2_pytorch_b200.Model")
        self.conditional = model.condition_on_prompt
        self.model = model
        self.codec = codec
        assert _exists(codec) or _exists(target_sample_hz)  # ns2.py:1207
        self.target_sample_hz = target_sample_hz
        self.seq_len_multiple_of = None
        if _exists(codec):
            self.target_sample_hz = codec.target_sample_hz
            self.seq_len_multiple_of = codec.seq_len_multiple_of
        assert not _exists(codec) or model.dim == codec.codebook_dim, \
            f"transformer model dimension {model.dim} must be equal to codec dimension {codec.codebook_dim}"
        self.dim = codec.codebook_dim if _exists(codec) else model.dim
        assert objective in {"x0", "eps", "v"}
        self.objective = objective
        sched = {"linear": simple_linear_schedule, "cosine": cosine_schedule, "sigmoid": sigmoid_schedule}
        if noise_schedule not in sched:
            raise ValueError(f"invalid noise schedule {noise_schedule}")
        assert scale <= 1, "scale must be less than or equal to 1"
        self.scale = scale
        self.gamma_schedule = partial(sched[noise_schedule], **schedule_kwargs)
        self.timesteps = timesteps
        self.use_ddim = use_ddim
        self.time_difference = time_difference
        self.train_prob_self_cond = train_prob_self_cond
        self.min_snr_loss_weight = min_snr_loss_weight
        self.min_snr_gamma = min_snr_gamma
        self.rvq_cross_entropy_loss_weight = rvq_cross_entropy_loss_weight
        if rvq_cross_entropy_loss_weight != 0:
            raise NotImplementedError("codec.rq cross-entropy head (SURVEY a17) is optional and not built")
        self.conditioner = conditioner
        self.cuda_graphs = cuda_graphs  # sampling loop: replay one captured CUDA graph per denoiser step
        self.conditioning_kwargs = conditioning_kwargs  # accepted for signature parity (encoder hyper-parameters)

    @property
    def device(self):
        return next(self.model.parameters()).device

    def get_sampling_timesteps(self, batch, *, device):
        """ns2.py:1303-1308."""
        times = torch.linspace(1., 0., self.timesteps + 1, device=device)
        times = times[None].expand(batch, -1)
        times = torch.stack((times[:, :-1], times[:, 1:]), dim=0)
        return times.unbind(dim=-1)

    # ------------------------------------------------------------------------------------------
    # sampling
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def ddim_sample(self, shape, prompt=None, time_difference=None, cond_scale=1., cond=None, *, noise=None):
        """ns2.py:1379-1431.  `noise` (optional) fixes the initial latent instead of drawing it."""
        batch, device = shape[0], self.device
        time_difference = self.time_difference if time_difference is None else time_difference
        time_pairs = self.get_sampling_timesteps(batch, device=device)
        audio = torch.randn(shape, device=device) if noise is None else noise.to(device).float().clone()
        conditioning = None
        if self.conditional:
            assert _exists(prompt) and _exists(cond)
            # timestep-invariant work (perceiver, prompt FiLM vector, aligned-condition projection) once
            conditioning = self.model.precompute_conditioning(prompt, cond, shape[1])
        graphs_before = self.model.use_cuda_graphs
        self.model.use_cuda_graphs = graphs_before or self.cuda_graphs
        for times, times_next in time_pairs:
            gamma = self.gamma_schedule(times)
            gamma_next = self.gamma_schedule(times_next)
            alpha, sigma = gamma_to_alpha_sigma(gamma, self.scale)
            alpha_next, sigma_next = gamma_to_alpha_sigma(gamma_next, self.scale)
            times_next = (times_next - time_difference).clamp(min=0.)
            if self.conditional:
                v = self.model.forward_with_cond_scale(audio, times, cond_scale=cond_scale,
                                                       _conditioning=conditioning)
            else:
                v = self.model.forward_with_cond_scale(audio, times, cond_scale=cond_scale)
            ops.ddim_step(audio, v, alpha.contiguous(), sigma.contiguous(), alpha_next.contiguous(),
                          sigma_next.contiguous(), objective=self.objective)
        self.model.use_cuda_graphs = graphs_before
        return audio

    def process_prompt(self, prompt=None):
        """ns2.py:1433-1447."""
        if not _exists(prompt):
            return None
        assert self.model.condition_on_prompt
        if prompt.ndim == 2:
            assert _exists(self.codec), "codec must be passed in if one were to train on raw prompt"
            with torch.no_grad():
                prompt, _, _ = self.codec(prompt, curtail_from_left=True, return_encoded=True)
        return prompt

    @torch.no_grad()
    def sample(self, *, length, prompt=None, batch_size=1, cond_scale=1., text=None, text_lens=None,
               prompt_enc=None, cond=None, noise=None):
        """ns2.py:1457-1501.  Conditional models need (`prompt_enc`, `cond`) or a `conditioner`."""
        if self.use_ddim is False:
            raise NotImplementedError("ddpm_sample is dead code in the reference (NameError: expm1, SURVEY T8)")
        if self.conditional:
            if not (_exists(prompt_enc) and _exists(cond)):
                if not _exists(self.conditioner):
                    raise NotImplementedError(
                        "conditional sampling needs prompt_enc= and cond= (outputs of the reference's "
                        "SpeechPromptEncoder / duration-pitch expansion) or a `conditioner` callable")
                prompt_enc, cond = self.conditioner(prompt=self.process_prompt(prompt), text=text,
                                                    text_lens=text_lens, mode="sample")
            batch_size = prompt_enc.shape[0]
        audio = self.ddim_sample((batch_size, length, self.dim), prompt=prompt_enc, cond=cond,
                                 cond_scale=cond_scale, noise=noise)
        if _exists(self.codec):
            audio = self.codec.decode(audio)
            if audio.ndim == 3 and audio.shape[1] == 1:
                audio = audio[:, 0]
        return audio

    # ------------------------------------------------------------------------------------------
    # training loss (forward only)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, audio, text=None, text_lens=None, mel=None, mel_lens=None, codes=None, prompt=None,
                pitch=None, *args, prompt_enc=None, cond=None, times=None, noise=None, **kwargs):
        """ns2.py:1503-1684 -> scalar diffusion loss (the only term the reference returns).
        Extra keyword-only arguments: `prompt_enc`/`cond` (precomputed conditioning) and `times`/`noise`
        (inject the two random draws of ns2.py:1621,1625 — used by the parity tests)."""
        is_raw_audio = audio.ndim == 2
        if self.conditional and not (_exists(prompt_enc) and _exists(cond)):
            if not _exists(self.conditioner):
                raise NotImplementedError(
                    "conditional training needs prompt_enc= and cond= or a `conditioner` callable (the "
                    "reference's encoders + aligner are outside the accelerated path)")
            prompt_enc, cond = self.conditioner(audio=audio, text=text, text_lens=text_lens, mel=mel,
                                                mel_lens=mel_lens, prompt=self.process_prompt(prompt),
                                                pitch=pitch, mode="train")
        assert not (is_raw_audio and not _exists(self.codec)), \
            "codec must be passed in if one were to train on raw audio"
        if is_raw_audio:
            audio, codes, _ = self.codec(audio, return_encoded=True)
        audio = audio.float().contiguous()
        batch, n, d = audio.shape
        device = self.device
        assert d == self.dim, f"codec codebook dimension {d} must match model dimensions {self.dim}"
        if times is None:
            times = torch.zeros((batch,), device=device).float().uniform_(0, 1.)
        if noise is None:
            noise = torch.randn_like(audio)
        times = times.to(device).float()
        noise = noise.to(device).float().contiguous()
        gamma = self.gamma_schedule(times)
        alpha, sigma = gamma_to_alpha_sigma(gamma, self.scale)
        alpha, sigma = alpha.contiguous(), sigma.contiguous()
        noised = torch.empty_like(audio)
        target = torch.empty_like(audio)
        ops.q_sample(audio, noise, alpha, sigma, noised, target, objective=self.objective)  # ns2.py:1631-1644
        pred = self.model(noised, times, prompt=prompt_enc, cond=cond)  # ns2.py:1635
        loss = ops.mse_rows(pred, target, torch.empty(batch, device=device))  # ns2.py:1646-1647
        # min-SNR weight on (B,)-sized tensors, with the reference's exact broadcasting (ns2.py:1651-1666):
        # loss is (B,), loss_weight is (B,1,1) -> the product is (B,1,B) before .mean()
        a3, s3 = alpha.view(-1, 1, 1), sigma.view(-1, 1, 1)
        snr = (a3 * a3) / (s3 * s3)
        clipped = snr.clone()
        if self.min_snr_loss_weight:
            clipped.clamp_(max=self.min_snr_gamma)
        if self.objective == "eps":       # ns2.py:1657-1664
            loss_weight = clipped / snr
        elif self.objective == "x0":
            loss_weight = clipped
        else:
            loss_weight = clipped / (snr + 1)
        return (loss * loss_weight).mean()

    p_losses = forward  # the name BASELINE.json's north_star uses; the reference inlines it in forward
