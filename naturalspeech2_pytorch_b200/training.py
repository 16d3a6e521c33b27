"""Training step of the denoiser: forward with saved activations + hand-written backward.

`loss.backward()` is how the reference is used (README.md:60-63, ns2.py:1886).  Here `Model.forward` records ONE autograd
node (`DenoiserFunction`) when gradients are enabled; its backward walks the network in reverse and launches, per layer,
  * dgrad GEMMs     ns2_gemm on transposed weight packs (anti-causal shifts for the causal convs),
  * wgrad GEMMs     ns2_wgrad (tcgen05, MN-major operands, fp32 reduce-add into the packed gradient),
  * attention bwd   ns2_attn_bwd (tcgen05 flash backward from the saved log-sum-exp),
  * the element-wise backward kernels of csrc/backward.cu (RMSNorm+FiLM, GEGLU, Wavenet gate, bias column sums).
Pre-activations that the fused forward epilogues never materialise (GEGLU's value/gate pair, the Wavenet conv output
before FiLM) are recomputed with plain-epilogue GEMMs instead of being stored.  Gradients come out in the packed bf16
layouts' fp32 twins and are scattered back to the reference's parameter shapes (same keys as the state_dict).

Scope: unconditional and conditional denoisers (BASELINE configs[1], configs[2]/[4]): perceiver resampler, cross
attention, prompt FiLM vector, aligned-condition projection and the classifier-free-guidance null parameters included.
The (B,)-sized conditioning vectors — timestep embedding (LearnedSinusoidalPosEmb + Linear + SiLU, ns2.py:108-120,
839-843) and prompt vector (mean-pool + Linear + SiLU, ns2.py:858-862) — are differentiated with torch autograd on a
recomputation: 32 x 2048 values each, host-side glue like the noise schedules.  `prompt_mask` is unsupported (as in
the inference path).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import ops

bf = torch.bfloat16


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def pack_transposed(model) -> Dict[str, torch.Tensor]:
    """bf16 transposed twins of `Model.packed()` for the dgrad GEMMs (rebuilt whenever the forward packs are)."""
    P = model.packed()
    D, G = model.dim, model.wavenet_layers
    T: Dict[str, torch.Tensor] = {}
    t = lambda w: w.t().contiguous()
    T["film_w"] = t(P["film_w"])                                        # (dim_cond, rows)
    for s in range(model.wavenet_stacks):
        w = P[f"wn{s}_w"].view(G, D, 4, D)                              # [group][out][tap0,tap1,tap2,res][in]
        T[f"wn{s}_w"] = w.permute(0, 3, 2, 1).reshape(G * D, 4 * D).contiguous()   # [group][in][tap][out]
    T["wn_skip_w"] = t(P["wn_skip_w"])                                  # (G*D, D)
    T["wn_final_w"] = t(P["wn_final_w"])
    for l in range(model.depth):
        T[f"l{l}_qkv"] = t(P[f"l{l}_qkv"])                              # (D, 3*inner)
        T[f"l{l}_o"] = t(P[f"l{l}_o"])                                  # (inner, D)
        T[f"l{l}_ff_w1"] = t(P[f"l{l}_ff_w1"])                          # (D, 2*Dp)
        wc = P[f"l{l}_ff_wc"]                                           # (Dp, 3*Dp) tap-major columns
        Dp = wc.shape[0]
        T[f"l{l}_ff_wc"] = wc.view(Dp, 3, Dp).permute(2, 1, 0).reshape(Dp, 3 * Dp).contiguous()   # [in][tap][out]
        T[f"l{l}_ff_w2"] = t(P[f"l{l}_ff_w2"])                          # (Dp, D)
    T["pred_w"] = t(P["pred_w"])
    T["wn_init_w"] = P["wn_init_w"].view(D, 3, D).permute(2, 1, 0).reshape(D, 3 * D).contiguous()   # [in][tap][out]
    if model.condition_on_prompt:
        T["x_kv_all"] = t(P["x_kv_all"])                               # (D, depth*2*inner)
        for l in range(model.depth):
            T[f"l{l}_xq"] = t(P[f"l{l}_xq"])
            T[f"l{l}_xo"] = t(P[f"l{l}_xo"])
        for i in range(len(model.perceiver_resampler.layers)):
            for k in ("q", "kv", "o", "ff_w1", "ff_w2"):
                T[f"pr{i}_{k}"] = t(P[f"pr{i}_{k}"])
    return T


def _perceiver_forward(model, prompt_f, S):
    """PerceiverResampler.forward (ns2.py:568-579) keeping per-layer activations -> tokens (B, M, D) fp32."""
    P, D, M, inner, H = model.packed(), model.dim, model.num_latents_m, model.inner, model.heads
    pr = model.perceiver_resampler
    B, Np, _ = prompt_f.shape
    dev = prompt_f.device
    e = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)
    ctx = M + Np
    Dp = _round_up(model.ff_inner, 128)
    p_bf = ops.cast_bf16(prompt_f, e(B, Np, model.dim_prompt))
    if "pr_proj_w" in P:
        proj = ops.gemm(p_bf, P["pr_proj_w"], e(B, Np, D), n=D, epilogue=ops.EPI_BF16, bias=P["pr_proj_b"])
    else:
        proj = p_bf
    lat = pr.latents.detach().float().unsqueeze(0).expand(B, M, D).contiguous()
    layers = []
    for i in range(len(pr.layers)):
        L = {}
        L["lat_bf"] = ops.cast_bf16(lat, e(B, M, D))
        cat = e(B, ctx, D)
        cat[:, :M].copy_(L["lat_bf"])     # cross_attn_include_queries: keys = cat(latents, context) (ns2.py:1060-1061)
        cat[:, M:].copy_(proj)
        L["cat"] = cat
        L["q"] = ops.gemm(L["lat_bf"], P[f"pr{i}_q"], e(B, M, inner), n=inner, epilogue=ops.EPI_BF16)
        L["kv"] = ops.gemm(cat, P[f"pr{i}_kv"], e(B, ctx, 2 * inner), n=2 * inner, epilogue=ops.EPI_BF16)
        L["lse"] = e(B, H, M, dt=torch.float32)
        L["o"] = ops.attention(L["q"], L["kv"][:, :, :inner], L["kv"][:, :, inner:], e(B, M, inner), heads=H, lse=L["lse"])
        ops.gemm(L["o"], P[f"pr{i}_o"], lat, n=D, epilogue=ops.EPI_F32, resid=lat)
        L["lat_bf2"] = ops.cast_bf16(lat, e(B, M, D))
        L["g"] = ops.gemm(L["lat_bf2"], P[f"pr{i}_ff_w1"], e(B, M, Dp), n=2 * Dp, epilogue=ops.EPI_GEGLU, bias=P[f"pr{i}_ff_b1"])
        ops.gemm(L["g"], P[f"pr{i}_ff_w2"], lat, n=D, epilogue=ops.EPI_F32, bias=P[f"pr{i}_ff_b2"], resid=lat)
        layers.append(L)
    S.update(pr_layers=layers, pr_lat=lat, pr_p_bf=p_bf, pr_Np=Np)
    return ops.rmsnorm_f32(lat, e(B, M, D, dt=torch.float32), pr.norm.gamma.detach().float().contiguous())


def train_forward(model, x: torch.Tensor, times: torch.Tensor, prompt=None, cond=None, cond_drop_prob=None):
    """Same arithmetic as `Model._forward_impl` (ns2.py:929-1000), keeping what the backward needs."""
    from .model import _prob_mask_like
    B, N, D = x.shape
    dev = x.device
    P = model.packed()
    G, inner, H = model.wavenet_layers, model.inner, model.heads
    Dp = _round_up(model.ff_inner, 128)
    e = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)
    S: Dict[str, object] = {"B": B, "N": N}
    tc = model.to_time_cond
    t = e(B, model.dim_cond, dt=torch.float32)
    ops.time_cond(times.float().contiguous(), tc[0].weights.detach().float().contiguous(),
                  tc[1].weight.detach().float().contiguous(), tc[1].bias.detach().float().contiguous(),
                  t[:, :model.dim_time])
    conditional = model.condition_on_prompt
    c_bf = None
    if conditional:
        assert prompt is not None and cond is not None, "prompt and cond are required when condition_on_prompt=True"
        M = model.num_latents_m
        p_eff = model.cond_drop_prob if cond_drop_prob is None else cond_drop_prob
        drop = _prob_mask_like((B,), p_eff, dev)          # same two draws, same order as the reference (ns2.py:950, 980)
        cdrop = _prob_mask_like((B,), p_eff, dev)
        prompt_f = prompt.float().contiguous()
        mean = ops.mean_rows(prompt_f, e(B, model.dim_prompt, dt=torch.float32))
        lin = model.to_prompt_cond[1]
        raw_pc = ops.small_linear(mean, lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous(),
                                  e(B, model.dim_time, dt=torch.float32), act=1)
        ops.select_rows(drop, model.null_prompt_cond.detach().float().contiguous(), raw_pc, t[:, model.dim_time:])
        tokens = _perceiver_forward(model, prompt_f, S)
        c_bf = ops.select_rows(drop, model.null_prompt_tokens.detach().float().contiguous(), tokens, e(B, M, D))
        Lc = cond.shape[-1]
        cond_bf = ops.transpose_cast(cond.float().contiguous(), e(B, Lc, model.dim_prompt))
        cond_proj = ops.gemm(cond_bf, P["cond_w"], e(B, Lc, D, dt=torch.float32), n=D, epilogue=ops.EPI_F32, bias=P["cond_b"])
        S.update(drop=drop, cdrop=cdrop, prompt_mean=mean, c_bf=c_bf, cond_bf=cond_bf, Lc=Lc)
    t_bf = ops.cast_bf16(t, e(1, B, model.dim_cond))
    film = ops.gemm(t_bf, P["film_w"], e(1, B, P["film_w"].shape[0], dt=torch.float32), n=P["film_w"].shape[0],
                    epilogue=ops.EPI_F32, bias=P["film_b"])[0]
    S.update(times=times.float().contiguous(), t=t, film=film)
    # ---- wavenet ----
    if conditional:
        x_bf = ops.cond_inject(x.float().contiguous(), cond_proj, e(B, N, D), drop_mask=cdrop,
                               null_cond=model.null_cond.detach().float().reshape(-1))
    else:
        x_bf = ops.cast_bf16(x.float().contiguous(), e(B, N, D))
    h0 = ops.gemm(x_bf, P["wn_init_w"], e(B, N, D), n=D, epilogue=ops.EPI_BF16, bias=P["wn_init_b"], segs=ops.conv3_segs(D))
    segs = ops.conv3_segs(D) + [(0, 3 * D, D, 0, 1)]
    dil = [2 ** i for i in range(G)]
    src, stack_out = h0, []
    for s in range(model.wavenet_stacks):
        dst = e(B, N, G * D)
        ops.gemm(src, P[f"wn{s}_w"], dst, n=D, epilogue=ops.EPI_WAVENET, bias=P[f"wn{s}_b"], bias1_off=G * D, segs=segs,
                 film=film[:, s * G * 2 * D:], film_group_stride=2 * D, groups=G,
                 a_group_col_stride=0 if s == 0 else D, b_group_row_stride=D, out_group_col_stride=D, dil=dil)
        stack_out.append(dst)
        src = dst
    skip = ops.gemm(src, P["wn_skip_w"], e(B, N, D), n=D, epilogue=ops.EPI_BF16, bias=P["wn_skip_b"])
    xr = ops.gemm(skip, P["wn_final_w"], e(B, N, D, dt=torch.float32), n=D, epilogue=ops.EPI_F32, bias=P["wn_final_b"])
    S.update(x_bf=x_bf, h0=h0, stack_out=stack_out, skip=skip)
    # ---- transformer ----
    layers: List[dict] = []
    npl = model._norms_per_layer
    if conditional:
        S["xkv"] = ops.gemm(c_bf, P["x_kv_all"], e(B, model.num_latents_m, model.depth * 2 * inner),
                            n=model.depth * 2 * inner, epilogue=ops.EPI_BF16)
    for l in range(model.depth):
        fo = model._film_tr_off + l * npl * 2 * D
        L: Dict[str, torch.Tensor] = {"x_in": xr.clone()}
        L["h1"] = ops.rmsnorm_film(xr, e(B, N, D), film=film[:, fo:fo + 2 * D])
        L["qkv"] = ops.gemm(L["h1"], P[f"l{l}_qkv"], e(B, N, 3 * inner), n=3 * inner, epilogue=ops.EPI_BF16)
        L["lse"] = e(B, H, N, dt=torch.float32)
        qkv = L["qkv"]
        L["ao"] = ops.attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], e(B, N, inner),
                                heads=H, lse=L["lse"])
        ops.gemm(L["ao"], P[f"l{l}_o"], xr, n=D, epilogue=ops.EPI_F32, resid=xr)
        if conditional:   # cross attention over the perceiver latents (ns2.py:800-803)
            fo2 = fo + 2 * D
            L["x_c"] = xr.clone()
            L["h_x"] = ops.rmsnorm_film(xr, e(B, N, D), film=film[:, fo2:fo2 + 2 * D])
            L["xq"] = ops.gemm(L["h_x"], P[f"l{l}_xq"], e(B, N, inner), n=inner, epilogue=ops.EPI_BF16)
            kv = S["xkv"][:, :, l * 2 * inner:(l + 1) * 2 * inner]
            L["lse2"] = e(B, H, N, dt=torch.float32)
            L["ao2"] = ops.attention(L["xq"], kv[:, :, :inner], kv[:, :, inner:], e(B, N, inner), heads=H, lse=L["lse2"])
            ops.gemm(L["ao2"], P[f"l{l}_xo"], xr, n=D, epilogue=ops.EPI_F32, resid=xr)
        L["x_mid"] = xr.clone()
        fo3 = fo + (npl - 1) * 2 * D
        L["h2"] = ops.rmsnorm_film(xr, e(B, N, D), film=film[:, fo3:fo3 + 2 * D])
        L["ff_g"] = ops.gemm(L["h2"], P[f"l{l}_ff_w1"], e(B, N, Dp), n=2 * Dp, epilogue=ops.EPI_GEGLU, bias=P[f"l{l}_ff_b1"])
        L["ff_c"] = ops.gemm(L["ff_g"], P[f"l{l}_ff_wc"], e(B, N, Dp), n=Dp, epilogue=ops.EPI_BF16, bias=P[f"l{l}_ff_bc"],
                             segs=ops.conv3_segs(Dp))
        ops.gemm(L["ff_c"], P[f"l{l}_ff_w2"], xr, n=D, epilogue=ops.EPI_F32, bias=P[f"l{l}_ff_b2"], resid=xr)
        layers.append(L)
    S["x_final"] = xr
    S["hf"] = ops.rmsnorm_film(xr, e(B, N, D), gamma=P["pred_gamma"])
    out = ops.gemm(S["hf"], P["pred_w"], e(B, N, D, dt=torch.float32), n=D, epilogue=ops.EPI_F32)
    S["layers"] = layers
    return out, S


def train_backward(model, S: dict, d_out: torch.Tensor, reducer=None) -> Dict[str, torch.Tensor]:
    """Gradients of every parameter (keys of `model.named_parameters()`), given d(loss)/d(prediction).
    `reducer` (parallel.GradReducer): finished gradient buffers are handed over layer by layer, so that their
    all-reduce overlaps the rest of the backward pass."""
    flush = (lambda: reducer.reduce_all(grads)) if reducer is not None else (lambda: None)
    B, N = S["B"], S["N"]
    D, G, inner, H = model.dim, model.wavenet_layers, model.inner, model.heads
    Di = model.ff_inner
    Dp = _round_up(Di, 128)
    dev = d_out.device
    P = model.packed()
    T = model.packed_transposed()
    film = S["film"]
    e = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)
    z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
    grads: Dict[str, torch.Tensor] = {}
    dfilm = z(B, film.shape[1])
    t_cond = S["t"].contiguous()
    dil = [2 ** i for i in range(G)]

    # ---- to_pred: Linear (no bias) after RMSNorm(gamma) ----
    dout_bf = ops.cast_bf16(d_out.float().contiguous(), e(B, N, D))
    grads["transformer.to_pred.1.weight"] = ops.wgrad(dout_bf, S["hf"], z(D, D), n=D, k=D)
    dhf = ops.gemm(dout_bf, T["pred_w"], e(B, N, D), n=D, epilogue=ops.EPI_BF16)
    dxr = z(B, N, D)                       # fp32 gradient of the residual stream
    dxr_bf = e(B, N, D)
    dgam = z(D)
    ops.rmsnorm_film_bwd(S["x_final"], dhf, dxr, dxr_bf, rows_per_batch=N, gamma=P["pred_gamma"], dgamma=dgam)
    grads["transformer.to_pred.0.gamma"] = dgam

    npl = model._norms_per_layer
    conditional = model.condition_on_prompt
    if conditional:
        M = model.num_latents_m
        d_xkv = e(B, M, model.depth * 2 * inner)
    pre = e(B, N, 2 * Dp)
    for l in reversed(range(model.depth)):
        L = S["layers"][l]
        pfx = f"transformer.layers.{l}."
        fo = model._film_tr_off + l * npl * 2 * D
        fo3 = fo + (npl - 1) * 2 * D
        # ---- feed-forward branch: x += W2 conv(GEGLU(W1 h2)) ----
        dW2 = ops.wgrad(dxr_bf, L["ff_c"], z(D, Dp), n=D, k=Dp)
        grads[pfx + "5.3.weight"] = dW2[:, :Di]
        grads[pfx + "5.3.bias"] = ops.colsum(dxr_bf, z(D))
        d_c = ops.gemm(dxr_bf, T[f"l{l}_ff_w2"], e(B, N, Dp), n=Dp, epilogue=ops.EPI_BF16)
        dWc = z(Dp, 3 * Dp)
        for tap in range(3):   # tap t multiplies g[n - (2 - t)]
            ops.wgrad(d_c, L["ff_g"], dWc[:, tap * Dp:(tap + 1) * Dp], n=Dp, k=Dp, shift_units=2 - tap)
        grads[pfx + "5.2.1.weight"] = dWc.view(Dp, 3, Dp)[:Di, :, :Di].permute(0, 2, 1)
        grads[pfx + "5.2.1.bias"] = ops.colsum(d_c, z(Dp))[:Di]
        d_g = ops.gemm(d_c, T[f"l{l}_ff_wc"], e(B, N, Dp), n=Dp, epilogue=ops.EPI_BF16,
                       segs=[(0, tap * Dp, Dp, -(2 - tap), 0) for tap in range(3)])
        ops.gemm(L["h2"], P[f"l{l}_ff_w1"], pre, n=2 * Dp, epilogue=ops.EPI_BF16, bias=P[f"l{l}_ff_b1"])   # recompute
        ops.geglu_bwd(pre, d_g)                                                                              # pre <- d pre
        dW1 = ops.wgrad(pre, L["h2"], z(2 * Dp, D), n=2 * Dp, k=D).view(Dp // 128, 2, 128, D)
        db1 = ops.colsum(pre, z(2 * Dp)).view(Dp // 128, 2, 128)
        grads[pfx + "5.0.weight"] = torch.cat((dW1[:, 0].reshape(Dp, D)[:Di], dW1[:, 1].reshape(Dp, D)[:Di]), dim=0)
        grads[pfx + "5.0.bias"] = torch.cat((db1[:, 0].reshape(Dp)[:Di], db1[:, 1].reshape(Dp)[:Di]), dim=0)
        dh2 = ops.gemm(pre, T[f"l{l}_ff_w1"], e(B, N, D), n=D, epilogue=ops.EPI_BF16)
        ops.rmsnorm_film_bwd(L["x_mid"], dh2, dxr, dxr_bf, rows_per_batch=N, film=film[:, fo3:fo3 + 2 * D],
                             dfilm=dfilm[:, fo3:fo3 + 2 * D])
        # ---- cross-attention branch: x += Wxo attn(Wxq h_x, Wxkv c) ----
        if conditional:
            fo2 = fo + 2 * D
            grads[pfx + "3.to_out.weight"] = ops.wgrad(dxr_bf, L["ao2"], z(D, inner), n=D, k=inner)
            d_ao2 = ops.gemm(dxr_bf, T[f"l{l}_xo"], e(B, N, inner), n=inner, epilogue=ops.EPI_BF16)
            kv = S["xkv"][:, :, l * 2 * inner:(l + 1) * 2 * inner]
            dkv = d_xkv[:, :, l * 2 * inner:(l + 1) * 2 * inner]
            dq2 = z(B, N, inner)
            ops.attention_bwd(L["xq"], kv[:, :, :inner], kv[:, :, inner:], L["ao2"], d_ao2, L["lse2"], dq2,
                              dkv[:, :, :inner], dkv[:, :, inner:], heads=H)
            dq2_bf = ops.cast_bf16(dq2, e(B, N, inner))
            grads[pfx + "3.to_q.weight"] = ops.wgrad(dq2_bf, L["h_x"], z(inner, D), n=inner, k=D)
            dh_x = ops.gemm(dq2_bf, T[f"l{l}_xq"], e(B, N, D), n=D, epilogue=ops.EPI_BF16)
            ops.rmsnorm_film_bwd(L["x_c"], dh_x, dxr, dxr_bf, rows_per_batch=N, film=film[:, fo2:fo2 + 2 * D],
                                 dfilm=dfilm[:, fo2:fo2 + 2 * D])
        # ---- attention branch: x += Wo attn(Wqkv h1) ----
        grads[pfx + "1.to_out.weight"] = ops.wgrad(dxr_bf, L["ao"], z(D, inner), n=D, k=inner)
        d_ao = ops.gemm(dxr_bf, T[f"l{l}_o"], e(B, N, inner), n=inner, epilogue=ops.EPI_BF16)
        qkv = L["qkv"]
        d_qkv = e(B, N, 3 * inner)
        dq_acc = z(B, N, inner)
        ops.attention_bwd(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], L["ao"], d_ao, L["lse"],
                          dq_acc, d_qkv[:, :, inner:2 * inner], d_qkv[:, :, 2 * inner:], heads=H)
        d_qkv[:, :, :inner].copy_(dq_acc)   # fp32 accumulator -> bf16 slot (layout glue)
        dWqkv = ops.wgrad(d_qkv, L["h1"], z(3 * inner, D), n=3 * inner, k=D)
        grads[pfx + "1.to_q.weight"] = dWqkv[:inner]
        grads[pfx + "1.to_kv.weight"] = dWqkv[inner:]
        dh1 = ops.gemm(d_qkv, T[f"l{l}_qkv"], e(B, N, D), n=D, epilogue=ops.EPI_BF16)
        ops.rmsnorm_film_bwd(L["x_in"], dh1, dxr, dxr_bf, rows_per_batch=N, film=film[:, fo:fo + 2 * D],
                             dfilm=dfilm[:, fo:fo + 2 * D])
        # FiLM projections of this layer's norms: their rows of dfilm are final now, so the weight gradient (the largest
        # gradient buffers of the model) joins this layer's all-reduce instead of trailing the whole backward
        dWl = ops.film_wgrad(dfilm[:, fo:fo + npl * 2 * D], t_cond, torch.empty(npl * 2 * D, model.dim_cond, device=dev),
                             accumulate=False)
        for k, idx in enumerate((0, 2, 4) if conditional else (0, 4)):
            grads[pfx + f"{idx}.to_gamma_beta.weight"] = dWl[k * 2 * D:(k + 1) * 2 * D]
        flush()   # this layer's gradients are final: their all-reduce overlaps the layers still to come

    if conditional:
        _conditioning_backward_tokens(model, S, T, d_xkv, grads)
    # ---- wavenet: final 1x1 conv, skip sum, 4 stacks of 8 dilation columns, init conv ----
    grads["wavenet.final_conv.weight"] = ops.wgrad(dxr_bf, S["skip"], z(D, D), n=D, k=D).unsqueeze(-1)
    grads["wavenet.final_conv.bias"] = ops.colsum(dxr_bf, z(D))
    d_skip = ops.gemm(dxr_bf, T["wn_final_w"], e(B, N, D), n=D, epilogue=ops.EPI_BF16)
    last = S["stack_out"][-1]
    dWskip = ops.wgrad(d_skip, last, z(D, G * D), n=D, k=G * D)
    dbskip = ops.colsum(d_skip, z(D))
    nst = model.wavenet_stacks
    for g in range(G):
        grads[f"wavenet.stacks.{nst - 1}.blocks.{g}.skip_conv.weight"] = dWskip[:, g * D:(g + 1) * D].unsqueeze(-1)
        grads[f"wavenet.stacks.{nst - 1}.blocks.{g}.skip_conv.bias"] = dbskip
    # dcy: [dc | dy] halves, so that one grouped dgrad GEMM reads the conv taps from dc and the 1x1 res conv from dy
    dcy = e(B, N, 2 * G * D)
    ops.gemm(d_skip, T["wn_skip_w"], dcy[:, :, G * D:], n=G * D, epilogue=ops.EPI_BF16)     # d y of the last stack
    c_pre = e(B, N, G * D)
    for s in reversed(range(nst)):
        x_in = S["stack_out"][s - 1] if s > 0 else S["h0"]
        gcs = D if s > 0 else 0
        fo_s = s * G * 2 * D
        # recompute the conv output (incl. bias) that FiLM + the gate consumed
        ops.gemm(x_in, P[f"wn{s}_w"], c_pre, n=D, epilogue=ops.EPI_BF16, bias=P[f"wn{s}_b"], segs=ops.conv3_segs(D),
                 groups=G, a_group_col_stride=gcs, b_group_row_stride=D, out_group_col_stride=D, dil=dil)
        dy = dcy[:, :, G * D:]
        dc = dcy[:, :, :G * D]
        ops.wavenet_gate_bwd(c_pre, dy, dc, film[:, fo_s:], dfilm[:, fo_s:], dim=D, groups=G, film_group_stride=2 * D)
        dWs = ops.film_wgrad(dfilm[:, fo_s:fo_s + G * 2 * D], t_cond, torch.empty(G * 2 * D, model.dim_cond, device=dev),
                             accumulate=False)   # this stack's FiLM projections (see the transformer loop)
        for g in range(G):
            grads[f"wavenet.stacks.{s}.blocks.{g}.to_time_cond.weight"] = dWs[g * 2 * D:(g + 1) * 2 * D]
        dWp = z(G * D, 4 * D)
        for tap in range(3):
            ops.wgrad(dc, x_in, dWp[:, tap * D:(tap + 1) * D], n=D, k=D, shift_units=2 - tap, groups=G,
                      dy_group_col_stride=D, x_group_col_stride=gcs, dw_group_row_stride=D, dil=dil)
        ops.wgrad(dy, x_in, dWp[:, 3 * D:], n=D, k=D, groups=G, dy_group_col_stride=D, x_group_col_stride=gcs,
                  dw_group_row_stride=D)
        dbc, dbr = ops.colsum(dc, z(G * D)), ops.colsum(dy, z(G * D))
        for g in range(G):
            blk = f"wavenet.stacks.{s}.blocks.{g}."
            w = dWp[g * D:(g + 1) * D]
            grads[blk + "conv.weight"] = w[:, :3 * D].view(D, 3, D).permute(0, 2, 1)
            grads[blk + "res_conv.weight"] = w[:, 3 * D:].unsqueeze(-1)
            grads[blk + "conv.bias"] = dbc[g * D:(g + 1) * D]
            grads[blk + "res_conv.bias"] = dbr[g * D:(g + 1) * D]
        # d(input of every column): anti-causal taps on dc + the transposed 1x1 on dy
        segs = [(0, tap * D, D, -(2 - tap), 0) for tap in range(3)] + [(G * D, 3 * D, D, 0, 0)]
        d_in = e(B, N, G * D)
        ops.gemm(dcy, T[f"wn{s}_w"], d_in, n=D, epilogue=ops.EPI_BF16, segs=segs, groups=G, a_group_col_stride=D,
                 b_group_row_stride=D, out_group_col_stride=D, dil=dil)
        flush()
        if s > 0:
            dcy[:, :, G * D:].copy_(d_in)       # becomes d y of the previous stack
        else:
            d_h0 = ops.group_sum(d_in, e(B, N, D), dim=D, groups=G)   # h0 feeds all columns
    dWi = z(D, 3 * D)
    for tap in range(3):
        ops.wgrad(d_h0, S["x_bf"], dWi[:, tap * D:(tap + 1) * D], n=D, k=D, shift_units=2 - tap)
    grads["wavenet.init_conv.weight"] = dWi.view(D, 3, D).permute(0, 2, 1)
    grads["wavenet.init_conv.bias"] = ops.colsum(d_h0, z(D))
    if conditional:
        # x_in = x + pad_or_curtail(where(cdrop, null_cond, cond_proj)) (ns2.py:978-992): d x_in from the init conv's dgrad
        d_xin = ops.gemm(d_h0, T["wn_init_w"], e(B, N, D), n=D, epilogue=ops.EPI_BF16,
                         segs=[(0, tap * D, D, -(2 - tap), 0) for tap in range(3)])
        Lc = S["Lc"]
        n_used = min(Lc, N)
        keep = (~S["cdrop"])[:, None, None]
        d_cp = torch.zeros(B, Lc, D, device=dev, dtype=bf)
        d_cp[:, :n_used] = torch.where(keep, d_xin[:, :n_used], torch.zeros((), device=dev, dtype=bf))   # masking glue
        grads["null_cond"] = (d_xin[:, :n_used].float() * S["cdrop"][:, None, None]).sum((0, 1)).unsqueeze(-1)
        grads["cond_to_model_dim.weight"] = ops.wgrad(d_cp, S["cond_bf"], z(D, model.dim_prompt), n=D,
                                                      k=model.dim_prompt).unsqueeze(-1)
        grads["cond_to_model_dim.bias"] = ops.colsum(d_cp, z(D))

    # ---- FiLM projections (one stacked matrix) and the timestep embedding ----
    rows = film.shape[1]
    dbf = dfilm.sum(0)
    dfilm_bf = ops.cast_bf16(dfilm, e(1, B, rows))
    dt = ops.gemm(dfilm_bf, T["film_w"], e(1, B, model.dim_cond, dt=torch.float32), n=model.dim_cond, epilogue=ops.EPI_F32)[0]
    if conditional:
        # prompt FiLM vector: where(drop, null_prompt_cond, silu(Linear(mean(prompt)))) (ns2.py:952-962); (B,)-sized glue
        d_pc = dt[:, model.dim_time:]
        grads["null_prompt_cond"] = (d_pc * S["drop"][:, None]).sum(0)
        lin = model.to_prompt_cond[1]
        with torch.enable_grad():
            lw = lin.weight.detach().float().requires_grad_(True)
            lb = lin.bias.detach().float().requires_grad_(True)
            F.silu(F.linear(S["prompt_mean"], lw, lb)).backward(d_pc * (~S["drop"])[:, None])
        grads["to_prompt_cond.1.weight"], grads["to_prompt_cond.1.bias"] = lw.grad, lb.grad
        dt = dt[:, :model.dim_time]
    off = 0
    for s in range(nst):
        for g in range(G):
            key = f"wavenet.stacks.{s}.blocks.{g}.to_time_cond."
            grads[key + "bias"] = dbf[off:off + 2 * D]
            off += 2 * D
    for l in range(model.depth):
        for idx in ((0, 2, 4) if conditional else (0, 4)):
            key = f"transformer.layers.{l}.{idx}.to_gamma_beta."
            grads[key + "bias"] = dbf[off:off + 2 * D]
            off += 2 * D
    # (B,)-sized timestep embedding: torch autograd on a recomputation (ns2.py:108-120, 839-843)
    tc = model.to_time_cond
    with torch.enable_grad():
        wts = tc[0].weights.detach().float().requires_grad_(True)
        lw = tc[1].weight.detach().float().requires_grad_(True)
        lb = tc[1].bias.detach().float().requires_grad_(True)
        tt = S["times"][:, None]
        freqs = tt * wts[None] * 2 * math.pi
        emb = torch.cat((tt, freqs.sin(), freqs.cos()), dim=-1)
        F.silu(F.linear(emb, lw, lb)).backward(dt.contiguous())
    grads["to_time_cond.0.weights"], grads["to_time_cond.1.weight"], grads["to_time_cond.1.bias"] = wts.grad, lw.grad, lb.grad
    if reducer is not None:
        flush()
        reducer.finish()
    return grads


def _conditioning_backward_tokens(model, S, T, d_xkv, grads):
    """Backward of everything that produced the cross-attention context: the stacked K/V projection of all layers,
    the null-token substitution and the PerceiverResampler (ns2.py:532-579, 964-968)."""
    P = model.packed()
    B, D, M, inner, H = S["B"], model.dim, model.num_latents_m, model.inner, model.heads
    Di = model.ff_inner
    Dp = _round_up(Di, 128)
    dev = d_xkv.device
    e = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)
    z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
    dWkv = ops.wgrad(d_xkv, S["c_bf"], z(model.depth * 2 * inner, D), n=model.depth * 2 * inner, k=D)
    for l in range(model.depth):
        grads[f"transformer.layers.{l}.3.to_kv.weight"] = dWkv[l * 2 * inner:(l + 1) * 2 * inner]
    d_c = ops.gemm(d_xkv, T["x_kv_all"], e(B, M, D), n=D, epilogue=ops.EPI_BF16).float()
    drop = S["drop"]
    grads["null_prompt_tokens"] = (d_c * drop[:, None, None]).sum(0)
    d_tok = ops.cast_bf16((d_c * (~drop)[:, None, None]).contiguous(), e(B, M, D))
    # ---- perceiver: final RMSNorm(gamma), then the layers in reverse ----
    pr = model.perceiver_resampler
    dlat, dlat_bf = z(B, M, D), e(B, M, D)
    dgam = z(D)
    ops.rmsnorm_film_bwd(S["pr_lat"], d_tok, dlat, dlat_bf, rows_per_batch=M, gamma=pr.norm.gamma.detach().float().contiguous(),
                         dgamma=dgam)
    grads["perceiver_resampler.norm.gamma"] = dgam
    Np = S["pr_Np"]
    ctx = M + Np
    d_proj = z(B, Np, D)
    pre = e(B, M, 2 * Dp)
    for i in reversed(range(len(pr.layers))):
        L = S["pr_layers"][i]
        pfx = f"perceiver_resampler.layers.{i}."
        # feed-forward (no conv, no pre-norm): lat += W2 GEGLU(W1 lat)
        grads[pfx + "1.2.weight"] = ops.wgrad(dlat_bf, L["g"], z(D, Dp), n=D, k=Dp)[:, :Di]
        grads[pfx + "1.2.bias"] = ops.colsum(dlat_bf, z(D))
        d_g = ops.gemm(dlat_bf, T[f"pr{i}_ff_w2"], e(B, M, Dp), n=Dp, epilogue=ops.EPI_BF16)
        ops.gemm(L["lat_bf2"], P[f"pr{i}_ff_w1"], pre, n=2 * Dp, epilogue=ops.EPI_BF16, bias=P[f"pr{i}_ff_b1"])
        ops.geglu_bwd(pre, d_g)
        dW1 = ops.wgrad(pre, L["lat_bf2"], z(2 * Dp, D), n=2 * Dp, k=D).view(Dp // 128, 2, 128, D)
        db1 = ops.colsum(pre, z(2 * Dp)).view(Dp // 128, 2, 128)
        grads[pfx + "1.0.weight"] = torch.cat((dW1[:, 0].reshape(Dp, D)[:Di], dW1[:, 1].reshape(Dp, D)[:Di]), dim=0)
        grads[pfx + "1.0.bias"] = torch.cat((db1[:, 0].reshape(Dp)[:Di], db1[:, 1].reshape(Dp)[:Di]), dim=0)
        ops.accum_bf16(dlat, ops.gemm(pre, T[f"pr{i}_ff_w1"], e(B, M, D), n=D, epilogue=ops.EPI_BF16), dlat_bf)
        # attention over cat(latents, projected prompt): lat += Wo attn(Wq lat, Wkv cat)
        grads[pfx + "0.to_out.weight"] = ops.wgrad(dlat_bf, L["o"], z(D, inner), n=D, k=inner)
        d_o = ops.gemm(dlat_bf, T[f"pr{i}_o"], e(B, M, inner), n=inner, epilogue=ops.EPI_BF16)
        dq = z(B, M, inner)
        d_kv = e(B, ctx, 2 * inner)
        ops.attention_bwd(L["q"], L["kv"][:, :, :inner], L["kv"][:, :, inner:], L["o"], d_o, L["lse"], dq, d_kv[:, :, :inner],
                          d_kv[:, :, inner:], heads=H)
        dq_bf = ops.cast_bf16(dq, e(B, M, inner))
        grads[pfx + "0.to_q.weight"] = ops.wgrad(dq_bf, L["lat_bf"], z(inner, D), n=inner, k=D)
        grads[pfx + "0.to_kv.weight"] = ops.wgrad(d_kv, L["cat"], z(2 * inner, D), n=2 * inner, k=D)
        d_cat = ops.gemm(d_kv, T[f"pr{i}_kv"], e(B, ctx, D), n=D, epilogue=ops.EPI_BF16)
        ops.accum_bf16(dlat, ops.gemm(dq_bf, T[f"pr{i}_q"], e(B, M, D), n=D, epilogue=ops.EPI_BF16))
        ops.accum_bf16(dlat, d_cat[:, :M].contiguous(), dlat_bf)
        ops.accum_bf16(d_proj, d_cat[:, M:].contiguous())
    grads["perceiver_resampler.latents"] = dlat.sum(0)
    if "pr_proj_w" in P:
        d_proj_bf = ops.cast_bf16(d_proj, e(B, Np, D))
        grads["perceiver_resampler.proj_context.weight"] = ops.wgrad(d_proj_bf, S["pr_p_bf"], z(D, model.dim_prompt), n=D,
                                                                     k=model.dim_prompt)
        grads["perceiver_resampler.proj_context.bias"] = ops.colsum(d_proj_bf, z(D))


class DenoiserFunction(torch.autograd.Function):
    """One autograd node for the whole denoiser: forward saves activations, backward runs the kernels above."""

    @staticmethod
    def forward(ctx, model, x, times, prompt, cond, cond_drop_prob, *params):
        with torch.no_grad():
            out, saved = train_forward(model, x, times, prompt, cond, cond_drop_prob)
        ctx.model, ctx.saved = model, saved
        ctx.names = [n for n, _ in model.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, d_out):
        with torch.no_grad():
            grads = train_backward(ctx.model, ctx.saved, d_out, getattr(ctx.model, "grad_reducer", None))
        ctx.saved = None
        missing = [n for n in ctx.names if n not in grads]
        if missing:
            raise RuntimeError(f"backward produced no gradient for {missing[:4]}...")
        return (None, None, None, None, None, None, *[grads[n].reshape(p.shape).to(p.dtype)
                                    for n, p in ctx.model.named_parameters()])


class MseRowsFunction(torch.autograd.Function):
    """Per-sample mean squared error (ns2.py:1646-1647) with the hand-written forward / backward kernels."""

    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous(), target.contiguous()
        ctx.save_for_backward(pred, target)
        return ops.mse_rows(pred, target, torch.empty(pred.shape[0], device=pred.device))

    @staticmethod
    def backward(ctx, d_rows):
        pred, target = ctx.saved_tensors
        per = pred.numel() // pred.shape[0]
        coef = (d_rows.float() * (2.0 / per)).contiguous()
        return ops.mse_bwd(pred, target, coef, out_f32=torch.empty_like(pred)), None
