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
from collections import OrderedDict
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
        self.conditioner = conditioner
        self.cuda_graphs = cuda_graphs  # sampling loop: replay one captured CUDA graph per sampling step
        self._sampler_graphs = OrderedDict()
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
    def _schedule_tables(self, batch, device):
        """(times (T, B), coef (T, 4, B) = alpha, sigma, alpha_next, sigma_next) for every sampling step, computed with
        the reference's own element-wise formulas (ns2.py:1303-1308, 1396-1404), so the values are bit-identical to
        the per-step tensors the reference builds."""
        t_all = torch.linspace(1., 0., self.timesteps + 1, device=device)
        gamma = self.gamma_schedule(t_all)
        alpha, sigma = gamma_to_alpha_sigma(gamma, self.scale)
        coef = torch.stack((alpha[:-1], sigma[:-1], alpha[1:], sigma[1:]), dim=1)      # (T, 4)
        times = t_all[:-1, None].expand(-1, batch).contiguous()
        return times, coef[:, :, None].expand(-1, -1, batch).contiguous()

    def _sampler_entry(self, shape, conditioning, cond_scale, device):
        """One captured CUDA graph = one whole sampling step: denoiser forward(s), guidance combine, DDIM update of the
        static latent buffer.  Keyed on shapes only; conditioning is copied into static buffers."""
        guided = self.conditional and cond_scale != 1.
        cond_sig = None
        if conditioning is not None:
            cond_sig = tuple(tuple(v.shape) for v in conditioning.values() if torch.is_tensor(v))
        key = (tuple(shape), cond_sig, float(cond_scale) if guided else None, self.objective, str(device))
        entry = self._sampler_graphs.get(key)
        if entry is not None and entry["packed"] is self.model.packed():
            self._sampler_graphs.move_to_end(key)
            if conditioning is not None:
                for k, v in conditioning.items():
                    if torch.is_tensor(v):
                        entry["cond"][k].copy_(v)
            return entry
        while len(self._sampler_graphs) >= 4:
            self._sampler_graphs.popitem(last=False)
        B = shape[0]
        model = self.model
        x = torch.empty(shape, device=device, dtype=torch.float32)
        ts = torch.zeros(B, device=device, dtype=torch.float32)
        coef = torch.ones(4, B, device=device, dtype=torch.float32)
        v0 = torch.empty_like(x)
        v1 = torch.empty_like(x) if guided else None
        static_cond = None
        if conditioning is not None:
            static_cond = type(conditioning)({k: (v.clone() if torch.is_tensor(v) else v)
                                              for k, v in conditioning.items()})
        p_cond = 0. if self.conditional else None

        def step():
            model._forward_impl(x, ts, None, None, None, p_cond, static_cond, v0)
            if guided:   # classifier-free guidance (ns2.py:914-927): conditional + null forward, lerp
                model._forward_impl(x, ts, None, None, None, 1., static_cond, v1)
                ops.cfg_combine(v0, v1, cond_scale, v0)
            ops.ddim_step(x, v0, coef[0], coef[1], coef[2], coef[3], objective=self.objective)

        x.normal_()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up outside capture (workspaces, packing)
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            step()
        entry = {"graph": graph, "x": x, "ts": ts, "coef": coef, "cond": static_cond, "packed": model.packed()}
        self._sampler_graphs[key] = entry
        return entry

    @torch.no_grad()
    def ddim_sample(self, shape, prompt=None, time_difference=None, cond_scale=1., cond=None, *, noise=None):
        """ns2.py:1379-1431.  `noise` (optional) fixes the initial latent instead of drawing it.
        (`time_difference` only shifts a value the reference never reads again, ns2.py:1404-1406.)"""
        batch, device = shape[0], self.device
        audio = torch.randn(shape, device=device) if noise is None else noise.to(device).float().clone()
        conditioning = None
        if self.conditional:
            assert _exists(prompt) and _exists(cond)
            # timestep-invariant work (perceiver, prompt FiLM vector, aligned-condition projection) once
            conditioning = self.model.precompute_conditioning(prompt, cond, shape[1])
        times_tab, coef_tab = self._schedule_tables(batch, device)
        if self.cuda_graphs and audio.is_cuda and self.model._prof is None:
            entry = self._sampler_entry(shape, conditioning, cond_scale, device)
            entry["x"].copy_(audio)
            for i in range(self.timesteps):
                entry["ts"].copy_(times_tab[i])
                entry["coef"].copy_(coef_tab[i])
                entry["graph"].replay()
            return entry["x"].clone()
        for i in range(self.timesteps):
            if self.conditional:
                v = self.model.forward_with_cond_scale(audio, times_tab[i], cond_scale=cond_scale,
                                                       _conditioning=conditioning)
            else:
                v = self.model.forward_with_cond_scale(audio, times_tab[i], cond_scale=cond_scale)
            c = coef_tab[i]
            ops.ddim_step(audio, v, c[0], c[1], c[2], c[3], objective=self.objective)
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
    # training loss (differentiable: `loss.backward()` runs the hand-written backward kernels)
    # ------------------------------------------------------------------------------------------
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
        if pred.requires_grad:
            from .training import MseRowsFunction
            loss = MseRowsFunction.apply(pred, target)                  # ns2.py:1646-1647, with a backward kernel
        else:
            loss = ops.mse_rows(pred, target, torch.empty(batch, device=device))
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
        loss = (loss * loss_weight).mean()
        if self.rvq_cross_entropy_loss_weight == 0 or not _exists(codes):   # ns2.py:1670-1671
            return loss
        # cross entropy of the predicted x_start against the codec's codes (ns2.py:1673-1684)
        x_start = torch.empty_like(audio)
        ops.x_start_from_pred(audio, pred.detach(), alpha, sigma, x_start, objective=self.objective)
        _, ce_loss = self.codec.rq(x_start, codes)
        return loss + self.rvq_cross_entropy_loss_weight * ce_loss

    p_losses = forward  # the name BASELINE.json's north_star uses; the reference inlines it in forward
