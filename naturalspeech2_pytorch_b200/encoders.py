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
