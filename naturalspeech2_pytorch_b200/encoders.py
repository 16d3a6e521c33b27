"""Per-sample conditioning encoders of NaturalSpeech2 on the sm_100a kernels.

`SpeechPromptEncoder` (ns2.py:289-341) and `PhonemeEncoder` (ns2.py:228-287) run once per sample BEFORE the denoiser
loop (`NaturalSpeech2.forward` ns2.py:1537-1539, `sample` 1474-1476).  Both are a stack of k=9 convolutions with SiLU
followed by the plain `Transformer` (ns2.py:1073-1117: RMSNorm -> Attention -> +res, RMSNorm -> GEGLU FeedForward ->
+res).  Same constructor arguments, same parameter names and shapes as the reference, so a reference state_dict loads
unchanged; the module tree only HOLDS parameters, the math goes through `ops` (libns2b200.so):

  Conv1d(k=9, padding=4) + SiLU   one segmented tcgen05 GEMM with nine shifted-row segments (TMA zero fill = the
                                  "same" padding), SiLU in the epilogue (NS2_GEMM_FLAG_SILU)
  CausalConv1d(k=9) + SiLU        the same GEMM with shifts 8..0 (left padding only, ns2.py:583-595)
  nn.Embedding                    ops.embedding_bf16 (gather + padding substitution)
  Transformer layer               RMSNorm kernel -> fused QKV GEMM -> flash attention -> out-proj GEMM (+residual,
                                  fp32 stream) -> RMSNorm -> GEGLU GEMM -> out GEMM (+residual)

Forward / inference only (dropout is the identity in eval mode; the reference's training-mode dropout and the
backward of these encoders are not restated).  Attention masks are not supported (`mask=None` is what
NaturalSpeech2.forward / .sample pass, ns2.py:1475-1476, 1538-1539).  Numerics follow the denoiser: bf16 tensor-core
operands, fp32 accumulation, fp32 residual stream and norm statistics.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib, ops
from .model import _AttentionParams, _NoParam, _RMSNormParams, _feedforward_params, _round_up

_SILU = _lib.NS2_GEMM_FLAG_SILU


class _PlainTransformerParams(nn.Module):
    """Parameter holder of `Transformer` (ns2.py:1073-1108): layers.{l} = [RMSNorm, Attention, RMSNorm, FeedForward]."""

    def __init__(self, dim: int, depth: int, dim_head: int, heads: int, ff_mult: int = 4, final_norm: bool = False):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([_RMSNormParams(dim), _AttentionParams(dim, dim_head, heads), _RMSNormParams(dim),
                           _feedforward_params(dim, ff_mult, causal_conv=False)])
            for _ in range(depth)])
        self.norm = _RMSNormParams(dim) if final_norm else nn.Identity()


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def _pack_conv(w: torch.Tensor) -> torch.Tensor:
    """(O, I, k) -> (O, k*I) bf16, tap t at columns [t*I, (t+1)*I)."""
    O, I, k = w.shape
    return _bf(w.detach().permute(0, 2, 1).reshape(O, k * I))


def _conv_segs(c_in: int, kernel: int, first_shift: int):
    """Segments of a stride-1 convolution: tap t reads position n - (first_shift - t)."""
    return [(0, t * c_in, c_in, first_shift - t, 0) for t in range(kernel)]


class _EncoderBase(nn.Module):
    """Packing cache + the shared transformer forward."""

    def _init_cache(self):
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_sig = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    def invalidate_packed(self) -> None:
        self._packed = None
        self._packed_sig = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_packed"):
            self.invalidate_packed()
        return out

    def packed(self) -> Dict[str, torch.Tensor]:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or sig != self._packed_sig:
            with torch.no_grad():
                self._packed = self._pack()
            self._packed_sig = sig
        return self._packed

    # ---- transformer ----
    def _pack_transformer(self, P: Dict[str, torch.Tensor], tr: _PlainTransformerParams, dim: int) -> None:
        for l, (n1, attn, n2, ff) in enumerate(tr.layers):
            P[f"l{l}_g1"] = n1.gamma.detach().float().contiguous()
            P[f"l{l}_g2"] = n2.gamma.detach().float().contiguous()
            P[f"l{l}_qkv"] = _bf(torch.cat((attn.to_q.weight, attn.to_kv.weight), dim=0))
            P[f"l{l}_o"] = _bf(attn.to_out.weight)
            lin1, lin2 = ff[0], ff[-1]
            Di = lin2.weight.shape[1]
            Dp = _round_up(Di, 128)
            dev = lin1.weight.device
            wv, wg = torch.zeros(Dp, dim, device=dev), torch.zeros(Dp, dim, device=dev)
            wv[:Di], wg[:Di] = lin1.weight[:Di], lin1.weight[Di:]   # first half = value, second = gate (ns2.py:1006)
            bv, bg = torch.zeros(Dp, device=dev), torch.zeros(Dp, device=dev)
            bv[:Di], bg[:Di] = lin1.bias[:Di], lin1.bias[Di:]
            P[f"l{l}_w1"] = _bf(torch.stack((wv.view(-1, 128, dim), wg.view(-1, 128, dim)), dim=1).reshape(2 * Dp, dim))
            P[f"l{l}_b1"] = torch.stack((bv.view(-1, 128), bg.view(-1, 128)), dim=1).reshape(2 * Dp).float().contiguous()
            w2 = torch.zeros(dim, Dp, device=dev)
            w2[:, :Di] = lin2.weight
            P[f"l{l}_w2"] = _bf(w2)
            P[f"l{l}_b2"] = lin2.bias.detach().float().contiguous()
        if isinstance(tr.norm, _RMSNormParams):
            P["final_g"] = tr.norm.gamma.detach().float().contiguous()

    def _transformer(self, x: torch.Tensor, tr: _PlainTransformerParams, P, heads: int) -> torch.Tensor:
        """Transformer.forward (ns2.py:1110-1115) on the fp32 residual stream x (B, N, D), updated in place."""
        B, N, D = x.shape
        dev, bf = x.device, torch.bfloat16
        inner = heads * 64
        Dp = P["l0_w2"].shape[1] if len(tr.layers) else 0
        h = torch.empty(B, N, D, device=dev, dtype=bf)
        qkv = torch.empty(B, N, 3 * inner, device=dev, dtype=bf)
        o = torch.empty(B, N, inner, device=dev, dtype=bf)
        g = torch.empty(B, N, Dp, device=dev, dtype=bf)
        for l in range(len(tr.layers)):
            ops.rmsnorm_film(x, h, gamma=P[f"l{l}_g1"])
            ops.gemm(h, P[f"l{l}_qkv"], qkv, n=3 * inner, epilogue=ops.EPI_BF16)
            ops.attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], o, heads=heads)
            ops.gemm(o, P[f"l{l}_o"], x, n=D, epilogue=ops.EPI_F32, resid=x)
            ops.rmsnorm_film(x, h, gamma=P[f"l{l}_g2"])
            ops.gemm(h, P[f"l{l}_w1"], g, n=2 * Dp, epilogue=ops.EPI_GEGLU, bias=P[f"l{l}_b1"])
            ops.gemm(g, P[f"l{l}_w2"], x, n=D, epilogue=ops.EPI_F32, bias=P[f"l{l}_b2"], resid=x)
        if "final_g" in P:
            out = torch.empty_like(x)
            ops.rmsnorm_f32(x, out, P["final_g"])
            return out
        return x


def _check_transformer_dims(dim: int, dim_head: int):
    if dim_head != 64:
        raise NotImplementedError("the sm_100a attention kernel is specialised for dim_head=64")
    if dim % 128 != 0 or dim > 1024:
        raise NotImplementedError("transformer dim must be a multiple of 128 (<= 1024) for the sm_100a kernels")


class SpeechPromptEncoder(_EncoderBase):
    """ns2.py:289-341.  forward(x: (B, Np, dim_codebook)) -> (B, Np, dims[-1]) fp32."""

    def __init__(self, dim_codebook, dims: Tuple[int, ...] = (256, 2048, 2048, 2048, 2048, 512, 512, 512), *,
                 depth=6, heads=8, dim_head=64, dropout=0.2, kernel_size=9, padding=4, use_flash_attn=True):
        super().__init__()
        dims = [dim_codebook, *dims]
        self.dim, self.dim_out = dims[0], dims[-1]
        if kernel_size > _lib.NS2_GEMM_MAX_SEGS:
            raise NotImplementedError(f"kernel_size must be <= {_lib.NS2_GEMM_MAX_SEGS}")
        if 2 * padding != kernel_size - 1:
            raise NotImplementedError("only 'same' padding (2*padding == kernel_size-1) keeps the sequence length")
        if any(d % 64 for d in dims):
            raise NotImplementedError("channel counts must be multiples of 64 (tensor-core K blocks)")
        _check_transformer_dims(dims[-1], dim_head)
        self.kernel_size, self.padding, self.heads = kernel_size, padding, heads
        mods = [_NoParam()]                                  # Rearrange('b n c -> b c n')
        for d_in, d_out in zip(dims[:-1], dims[1:]):
            mods.extend([nn.Conv1d(d_in, d_out, kernel_size, padding=padding), _NoParam()])   # conv, SiLU
        mods.append(_NoParam())                              # Rearrange back
        self.conv = nn.Sequential(*mods)
        self.transformer = _PlainTransformerParams(dims[-1], depth, dim_head, heads)
        self._init_cache()

    def _convs(self):
        return [m for m in self.conv if isinstance(m, nn.Conv1d)]

    def _pack(self) -> Dict[str, torch.Tensor]:
        P: Dict[str, torch.Tensor] = {}
        for i, c in enumerate(self._convs()):
            P[f"c{i}_w"] = _pack_conv(c.weight)
            P[f"c{i}_b"] = c.bias.detach().float().contiguous()
        self._pack_transformer(P, self.transformer, self.dim_out)
        return P

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == self.dim
        if not x.is_cuda:
            raise ValueError("SpeechPromptEncoder: input must be a CUDA tensor (the ns2_b200 ops have no CPU path)")
        P = self.packed()
        B, N, _ = x.shape
        dev, bf = x.device, torch.bfloat16
        h = ops.cast_bf16(x.float().contiguous(), torch.empty(B, N, self.dim, device=dev, dtype=bf))
        convs = self._convs()
        for i, c in enumerate(convs):
            last = i == len(convs) - 1
            out = torch.empty(B, N, c.out_channels, device=dev, dtype=torch.float32 if last else bf)
            ops.gemm(h, P[f"c{i}_w"], out, n=c.out_channels, epilogue=ops.EPI_F32 if last else ops.EPI_BF16,
                     segs=_conv_segs(c.in_channels, self.kernel_size, self.padding), bias=P[f"c{i}_b"], flags=_SILU)
            h = out
        return self._transformer(h, self.transformer, P, self.heads)


class PhonemeEncoder(_EncoderBase):
    """ns2.py:228-287.  forward(x: (B, T) int64 phoneme ids, negative = padding) -> (B, T, dim_hidden) fp32.
    A tokenizer (List[str] input) is used exactly like the reference when one is given."""

    def __init__(self, *, tokenizer=None, num_tokens=None, dim=512, dim_hidden=512, kernel_size=9, depth=6,
                 dim_head=64, heads=8, conv_dropout=0.2, attn_dropout=0., use_flash=False):
        super().__init__()
        self.tokenizer = tokenizer
        if num_tokens is None and tokenizer is not None:
            num_tokens = tokenizer.vocab_size
        if num_tokens is None:
            raise NotImplementedError("PhonemeEncoder without a token table (nn.Identity embedding) is not supported")
        if kernel_size > _lib.NS2_GEMM_MAX_SEGS:
            raise NotImplementedError(f"kernel_size must be <= {_lib.NS2_GEMM_MAX_SEGS}")
        if dim % 64:
            raise NotImplementedError("dim must be a multiple of 64 (tensor-core K blocks)")
        _check_transformer_dims(dim_hidden, dim_head)
        self.dim, self.dim_hidden, self.kernel_size, self.heads = dim, dim_hidden, kernel_size, heads
        self.token_emb = nn.Embedding(num_tokens + 1, dim)
        self.pad_id = num_tokens
        self.conv = nn.Sequential(_NoParam(), nn.Conv1d(dim, dim_hidden, kernel_size), _NoParam(), _NoParam(), _NoParam())
        self.transformer = _PlainTransformerParams(dim_hidden, depth, dim_head, heads)
        self._init_cache()

    def _pack(self) -> Dict[str, torch.Tensor]:
        c = self.conv[1]
        P = {"emb": self.token_emb.weight.detach().float().contiguous(), "c_w": _pack_conv(c.weight),
             "c_b": c.bias.detach().float().contiguous()}
        self._pack_transformer(P, self.transformer, self.dim_hidden)
        return P

    @torch.no_grad()
    def forward(self, x, mask=None) -> torch.Tensor:
        if mask is not None:
            raise NotImplementedError("PhonemeEncoder: attention masks are not supported by the sm_100a attention kernel")
        if isinstance(x, (list, tuple)):
            assert self.tokenizer is not None
            x = self.tokenizer.texts_to_tensor_ids(x).to(self.token_emb.weight.device)
        if not x.is_cuda:
            raise ValueError("PhonemeEncoder: input must be a CUDA tensor (the ns2_b200 ops have no CPU path)")
        P = self.packed()
        B, T = x.shape
        dev, bf = x.device, torch.bfloat16
        e = ops.embedding_bf16(x.long().contiguous(), P["emb"], torch.empty(B, T, self.dim, device=dev, dtype=bf),
                               self.pad_id)
        h = torch.empty(B, T, self.dim_hidden, device=dev, dtype=torch.float32)
        # CausalConv1d: left padding dilation*(k-1) (ns2.py:592-595) -> tap t reads position n - (k-1-t)
        ops.gemm(e, P["c_w"], h, n=self.dim_hidden, epilogue=ops.EPI_F32,
                 segs=_conv_segs(self.dim, self.kernel_size, self.kernel_size - 1), bias=P["c_b"], flags=_SILU)
        return self._transformer(h, self.transformer, P, self.heads)


# --------------------------------------------------------------------------------------------------
# duration / pitch predictor (ns2.py:345-527)
# --------------------------------------------------------------------------------------------------
class _BlockParams(nn.Module):
    """Block (ns2.py:345-365): proj = Conv1d(k, padding k//2), norm = GroupNorm(groups, dim_out)."""

    def __init__(self, dim, dim_out, kernel, groups):
        super().__init__()
        self.proj = nn.Conv1d(dim, dim_out, kernel, padding=kernel // 2)
        self.norm = nn.GroupNorm(groups, dim_out)


class _ResnetBlockParams(nn.Module):
    """ResnetBlock (ns2.py:367-401) with dim == dim_out (res_conv = Identity, the only shape the trunk builds)."""

    def __init__(self, dim, kernel, groups=8, num_convs=2):
        super().__init__()
        self.blocks = nn.Sequential(*[_BlockParams(dim, dim, kernel, groups) for _ in range(num_convs)])
        self.res_conv = nn.Identity()


class _TrunkParams(nn.Module):
    """DurationPitchPredictorTrunk (ns2.py:412-456)."""

    def __init__(self, dim, depth, kernel_size, dim_context, heads, dim_head, num_convs_per_resnet_block,
                 num_convolutions_per_block):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            attn = _AttentionParams(dim, dim_head, heads)
            if dim_context is not None and dim_context != dim:
                attn.to_kv = nn.Linear(dim_context, dim_head * heads * 2, bias=False)
            self.layers.append(nn.ModuleList([
                nn.Sequential(*[_ResnetBlockParams(dim, kernel_size, num_convs=num_convs_per_resnet_block)
                                for _ in range(num_convolutions_per_block)]),
                _RMSNormParams(dim),
                attn]))
        self.to_pred = nn.Sequential(nn.Linear(dim, 1), _NoParam(), _NoParam())


class DurationPitchPredictor(_EncoderBase):
    """ns2.py:468-527.  forward(x: (B, T, dim_hidden) phoneme encodings [or (B, T) ids with a token table],
    encoded_prompts: (B, Np, dim_encoded_prompts)) -> (duration_pred (B, T), pitch_pred (B, T)), both fp32 >= 0."""

    def __init__(self, *, dim, num_phoneme_tokens=None, tokenizer=None, dim_encoded_prompts=None,
                 num_convolutions_per_block=3, use_resnet_block=True, num_convs_per_resnet_block=2, depth=10,
                 kernel_size=3, heads=8, dim_head=64, dim_hidden=512, dropout=0.2, use_flash_attn=False):
        super().__init__()
        self.tokenizer = tokenizer
        if num_phoneme_tokens is None and tokenizer is not None:
            num_phoneme_tokens = tokenizer.vocab_size
        dim_encoded_prompts = dim if dim_encoded_prompts is None else dim_encoded_prompts
        if not use_resnet_block:
            raise NotImplementedError("only the ResnetBlock trunk (the reference default) is built")
        if kernel_size % 2 != 1 or kernel_size > _lib.NS2_GEMM_MAX_SEGS:
            raise NotImplementedError("kernel_size must be odd and <= NS2_GEMM_MAX_SEGS")
        _check_transformer_dims(dim_hidden, dim_head)
        if dim_encoded_prompts != dim_hidden:
            # keys = cat(norm(x), encoded_prompts) along the sequence (cross_attn_include_queries, ns2.py:1060-1061)
            raise NotImplementedError("dim_encoded_prompts must equal dim_hidden (the reference concatenates them)")
        if num_phoneme_tokens is not None and dim != dim_hidden:
            raise NotImplementedError("the token table width must equal dim_hidden")
        self.dim_hidden, self.heads, self.kernel_size = dim_hidden, heads, kernel_size
        self.phoneme_token_emb = nn.Embedding(num_phoneme_tokens, dim) if num_phoneme_tokens is not None else nn.Identity()
        mk = lambda: _TrunkParams(dim_hidden, depth, kernel_size, dim_encoded_prompts, heads, dim_head,  # noqa: E731
                                  num_convs_per_resnet_block, num_convolutions_per_block)
        self.to_pitch_pred = mk()
        self.to_duration_pred = mk()
        self._init_cache()

    def _pack(self) -> Dict[str, torch.Tensor]:
        P: Dict[str, torch.Tensor] = {}
        for name, trunk in (("p", self.to_pitch_pred), ("d", self.to_duration_pred)):
            for l, (convs, norm, attn) in enumerate(trunk.layers):
                for r, rb in enumerate(convs):
                    for c, blk in enumerate(rb.blocks):
                        k = f"{name}{l}_{r}_{c}"
                        P[k + "_w"] = _pack_conv(blk.proj.weight)
                        P[k + "_b"] = blk.proj.bias.detach().float().contiguous()
                        P[k + "_gw"] = blk.norm.weight.detach().float().contiguous()
                        P[k + "_gb"] = blk.norm.bias.detach().float().contiguous()
                P[f"{name}{l}_g"] = norm.gamma.detach().float().contiguous()
                P[f"{name}{l}_q"] = _bf(attn.to_q.weight)
                P[f"{name}{l}_kv"] = _bf(attn.to_kv.weight)
                P[f"{name}{l}_o"] = _bf(attn.to_out.weight)
            P[f"{name}_pw"] = trunk.to_pred[0].weight.detach().float().reshape(-1).contiguous()
            P[f"{name}_pb"] = trunk.to_pred[0].bias.detach().float().contiguous()
        if isinstance(self.phoneme_token_emb, nn.Embedding):
            P["emb"] = self.phoneme_token_emb.weight.detach().float().contiguous()
        return P

    def _trunk(self, name: str, trunk: _TrunkParams, P, x0: torch.Tensor, prompts_bf: torch.Tensor) -> torch.Tensor:
        B, T, D = x0.shape
        Np = prompts_bf.shape[1]
        dev, bf, H = x0.device, torch.bfloat16, self.heads
        inner = H * 64
        groups = trunk.layers[0][0][0].blocks[0].norm.num_groups if len(trunk.layers) else 8
        eps = trunk.layers[0][0][0].blocks[0].norm.eps if len(trunk.layers) else 1e-5
        segs = _conv_segs(D, self.kernel_size, self.kernel_size // 2)
        x = x0.clone()                                               # fp32 stream of this trunk
        x_bf = ops.cast_bf16(x, torch.empty(B, T, D, device=dev, dtype=bf))
        c = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        h_bf = torch.empty(B, T, D, device=dev, dtype=bf)
        ctx = torch.empty(B, T + Np, D, device=dev, dtype=bf)        # [norm(x) ; encoded prompts] (ns2.py:1060-1061)
        ctx[:, T:].copy_(prompts_bf)
        nx = torch.empty(B, T, D, device=dev, dtype=bf)
        q = torch.empty(B, T, inner, device=dev, dtype=bf)
        kv = torch.empty(B, T + Np, 2 * inner, device=dev, dtype=bf)
        o = torch.empty(B, T, inner, device=dev, dtype=bf)
        for l, (convs, _, _) in enumerate(trunk.layers):
            for r, rb in enumerate(convs):
                nb = len(rb.blocks)
                src = x_bf
                for ci in range(nb):
                    k = f"{name}{l}_{r}_{ci}"
                    ops.gemm(src, P[k + "_w"], c, n=D, epilogue=ops.EPI_F32, segs=segs, bias=P[k + "_b"])
                    if ci < nb - 1:
                        ops.groupnorm_silu(c, P[k + "_gw"], P[k + "_gb"], groups, eps=eps, out_bf16=h_bf)
                        src = h_bf
                    else:   # out = blocks(x) + res_conv(x), res_conv = Identity (ns2.py:399-401)
                        ops.groupnorm_silu(c, P[k + "_gw"], P[k + "_gb"], groups, eps=eps, resid=x, out_f32=x,
                                           out_bf16=x_bf)
            ops.rmsnorm_film(x, nx, gamma=P[f"{name}{l}_g"])
            ctx[:, :T].copy_(nx)
            ops.gemm(nx, P[f"{name}{l}_q"], q, n=inner, epilogue=ops.EPI_BF16)
            ops.gemm(ctx, P[f"{name}{l}_kv"], kv, n=2 * inner, epilogue=ops.EPI_BF16)
            ops.attention(q, kv[:, :, :inner], kv[:, :, inner:], o, heads=H)
            ops.gemm(o, P[f"{name}{l}_o"], x, n=D, epilogue=ops.EPI_F32, resid=x)   # attn(norm(x), prompts) + x
            ops.cast_bf16(x, x_bf)
        pred = torch.empty(B, T, device=dev, dtype=torch.float32)
        ops.rowdot(x, P[f"{name}_pw"], P[f"{name}_pb"], pred, relu=True)            # Linear(dim, 1) + ReLU
        return pred

    @torch.no_grad()
    def forward(self, x, encoded_prompts: torch.Tensor, prompt_mask=None):
        if prompt_mask is not None:
            raise NotImplementedError("DurationPitchPredictor: prompt masks are not supported by the sm_100a attention kernel")
        if isinstance(x, (list, tuple)):
            assert self.tokenizer is not None
            x = self.tokenizer.texts_to_tensor_ids(x).to(encoded_prompts.device)
        if not (x.is_cuda and encoded_prompts.is_cuda):
            raise ValueError("DurationPitchPredictor: inputs must be CUDA tensors (the ns2_b200 ops have no CPU path)")
        P = self.packed()
        dev, bf = x.device, torch.bfloat16
        if "emb" in P:
            B, T = x.shape
            e = ops.embedding_bf16(x.long().contiguous(), P["emb"],
                                   torch.empty(B, T, self.dim_hidden, device=dev, dtype=bf), 0)
            x = e.float()
        x = x.float().contiguous()
        B, Np, Dp = encoded_prompts.shape
        assert x.shape[-1] == self.dim_hidden and Dp == self.dim_hidden
        prompts_bf = ops.cast_bf16(encoded_prompts.float().contiguous(), torch.empty(B, Np, Dp, device=dev, dtype=bf))
        duration = self._trunk("d", self.to_duration_pred, P, x, prompts_bf)
        pitch = self._trunk("p", self.to_pitch_pred, P, x, prompts_bf)
        return duration, pitch


# --------------------------------------------------------------------------------------------------
# the conditional front end of NaturalSpeech2.sample (ns2.py:1472-1483)
# --------------------------------------------------------------------------------------------------
def f0_to_coarse(f0: torch.Tensor, f0_bin: int = 256, f0_max: float = 1100.0, f0_min: float = 50.0) -> torch.Tensor:
    """ns2.py:164-177 — (B, T)-sized host-side glue kept in torch like the noise schedules."""
    f0_mel_max = 1127 * torch.log(1 + torch.tensor(f0_max) / 700)
    f0_mel_min = 1127 * torch.log(1 + torch.tensor(f0_min) / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    pos = f0_mel > 0
    f0_mel = torch.where(pos, (f0_mel - f0_mel_min) * (f0_bin - 2) / (f0_mel_max - f0_mel_min) + 1, f0_mel)
    f0_mel = f0_mel.clamp(min=1, max=f0_bin - 1)
    return (f0_mel + 0.5).int()


def frames_to_text_index(duration: torch.Tensor) -> torch.Tensor:
    """The hard alignment of generate_mask_from_repeats (ns2.py:87-104) as one text index per frame: (B, L) int32,
    L = max total duration, -1 past a sample's own length.  mask[b, i, n] of the reference == (idx[b, n] == i)."""
    repeats = duration.int()
    cumsum = repeats.cumsum(dim=-1)
    lengths = cumsum[:, -1]
    L = int(lengths.amax().item())
    seq = torch.arange(L, device=duration.device).unsqueeze(0).expand(duration.shape[0], L).contiguous()
    idx = torch.searchsorted(cumsum, seq, right=True)          # first i with cumsum[i] > n
    idx = torch.where(seq < lengths.unsqueeze(-1), idx, torch.full_like(idx, -1))
    return idx.int().contiguous()


def expand_encodings(phoneme_enc: torch.Tensor, duration: torch.Tensor, pitch: torch.Tensor,
                     pitch_table: torch.Tensor) -> torch.Tensor:
    """cond (B, D, L) of ns2.py:1478-1483: phoneme encodings + coarse-pitch embeddings repeated `duration` frames."""
    idx = frames_to_text_index(duration)
    coarse = f0_to_coarse(pitch.float()).contiguous()
    return ops.expand_encodings(phoneme_enc.float().contiguous(), coarse, pitch_table.detach().float().contiguous(), idx)


class Conditioner(nn.Module):
    """The per-sample conditional front end of `NaturalSpeech2.sample` (ns2.py:1472-1483) as the `conditioner`
    callable of `naturalspeech2_pytorch_b200.NaturalSpeech2`: prompt latents + phoneme ids -> (prompt_enc, cond).
    Sub-module names follow the reference's NaturalSpeech2 attributes (ns2.py:1231-1236), so the matching slices of a
    reference checkpoint load with `load_state_dict(..., strict=False)`.  The training-time front end (mel, pitch
    extraction, aligner network, ns2.py:1537-1583) is not built: `mode="train"` raises."""

    def __init__(self, *, dim_codebook=128, num_phoneme_tokens=None, tokenizer=None, duration_pitch_dim=512,
                 pitch_emb_dim=256, pitch_emb_pp_hidden_dim=512):
        super().__init__()
        self.phoneme_enc = PhonemeEncoder(tokenizer=tokenizer, num_tokens=num_phoneme_tokens)
        self.prompt_enc = SpeechPromptEncoder(dim_codebook=dim_codebook)
        self.duration_pitch = DurationPitchPredictor(dim=duration_pitch_dim)
        self.pitch_emb = nn.Embedding(pitch_emb_dim, pitch_emb_pp_hidden_dim)

    @torch.no_grad()
    def forward(self, prompt=None, text=None, text_lens=None, mode="sample", **unused):
        if mode != "sample":
            raise NotImplementedError("Conditioner: only the sampling front end (ns2.py:1472-1483) is built")
        assert prompt is not None and text is not None
        prompt_enc = self.prompt_enc(prompt)
        phoneme_enc = self.phoneme_enc(text)
        duration, pitch = self.duration_pitch(phoneme_enc, prompt_enc)
        cond = expand_encodings(phoneme_enc, duration, pitch, self.pitch_emb.weight)
        return prompt_enc, cond
