"""`EncodecRVQ`: the residual-VQ step of the Encodec codec on sm_100a, behind the duck-type that
`NaturalSpeech2` expects from `audiolm_pytorch.EncodecWrapper` (ns2.py:1213-1214, 1244-1246, 1445, 1496, 1611).

In scope: nearest-codeword search over Q sequential residual stages (`ops.rvq_encode`, tcgen05
distance filter + exact fp64 re-score => bit-exact indices) and the sum-of-codewords decode (`ops.rvq_decode`).
Out of scope: Encodec's SEANet conv/LSTM encoder and decoder (pretrained weights are not available offline and
the north star does not name them).  They plug in as callables:
    encoder(raw_audio (B, T)) -> frames (B, N, 128)        decoder(emb (B, N, 128)) -> audio (B, 1, T)
Without an encoder the codec accepts encoder-output frames (B, N, 128) directly.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from . import ops


class EncodecRVQ(nn.Module):
    def __init__(self, codebooks: torch.Tensor, *, target_sample_hz: int = 24000, strides=(2, 4, 5, 8),
                 encoder: Optional[Callable] = None, decoder: Optional[Callable] = None):
        """codebooks: (Q, K, 128) fp32 — `model.quantizer.vq.layers[q]._codebook.embed` of an Encodec model."""
        super().__init__()
        if codebooks.dim() != 3 or codebooks.shape[-1] != 128:
            raise ValueError("codebooks must be (num_quantizers, codebook_size, 128)")
        self.register_buffer("codebooks", codebooks.detach().float().contiguous())
        self.target_sample_hz = target_sample_hz
        self.seq_len_multiple_of = 1
        for s in strides:
            self.seq_len_multiple_of *= s  # 320 for the 24 kHz model
        self.codebook_dim = codebooks.shape[-1]
        self.num_quantizers = codebooks.shape[0]
        self.encoder = encoder
        self.decoder = decoder
        self._prepared = None
        self._prepared_key = None

    def _prep(self):
        key = (self.codebooks.data_ptr(), self.codebooks._version, str(self.codebooks.device))
        if self._prepared is None or key != self._prepared_key:
            self._prepared = ops.rvq_prepare(self.codebooks)
            self._prepared_key = key
        return self._prepared

    @torch.no_grad()
    def quantize(self, frames: torch.Tensor, stats: Optional[torch.Tensor] = None):
        """frames (..., 128) fp32 -> (codes (..., Q) int64, emb (..., 128) fp32 = sum of the chosen codewords)."""
        shp = frames.shape[:-1]
        if frames.numel() == 0:  # empty batch: nothing to launch (the reference returns empty tensors too)
            return (torch.empty(*shp, self.num_quantizers, dtype=torch.int64, device=frames.device),
                    torch.empty(*shp, 128, dtype=torch.float32, device=frames.device))
        flat = frames.reshape(-1, 128).float().contiguous()
        codes = ops.rvq_encode(flat, self.codebooks, self._prep(), stats=stats)
        emb = ops.rvq_decode(codes, self.codebooks)
        return codes.view(*shp, self.num_quantizers), emb.view(*shp, 128)

    @torch.no_grad()
    def forward(self, x, return_encoded: bool = False, curtail_from_left: bool = False, **kwargs):
        """Mirror of EncodecWrapper.forward's return convention: (emb (B,N,128), codes (B,N,Q), None)."""
        if x.ndim == 2:
            if self.encoder is None:
                raise NotImplementedError(
                    "raw audio needs an `encoder` callable (Encodec's SEANet encoder is outside the "
                    "accelerated path); pass encoder-output frames (B, N, 128) instead")
            m = self.seq_len_multiple_of
            T = x.shape[-1] // m * m
            x = x[..., -T:] if curtail_from_left else x[..., :T]
            x = self.encoder(x)
        codes, emb = self.quantize(x)
        if not return_encoded:
            return codes
        return emb, codes, None

    @torch.no_grad()
    def get_emb_from_indices(self, codes: torch.Tensor) -> torch.Tensor:
        shp = codes.shape[:-1]
        emb = ops.rvq_decode(codes.reshape(-1, self.num_quantizers), self.codebooks)
        return emb.view(*shp, 128)

    @torch.no_grad()
    def decode(self, emb: torch.Tensor) -> torch.Tensor:
        if self.decoder is None:
            return emb
        return self.decoder(emb)

    @torch.no_grad()
    def rq(self, x: torch.Tensor, codes: torch.Tensor):
        """The call `NaturalSpeech2.forward` makes when rvq_cross_entropy_loss_weight != 0 (ns2.py:1682):
        `_, ce_loss = codec.rq(x_start, codes)` — vector-quantize-pytorch's ResidualVQ.forward(x, indices=codes).
        Per stage: logits = -||r_q - c_k|| (Euclidean), cross-entropy against codes[..., q] (ignore_index -1),
        residual chain through the codec's own nearest codewords; returns (quantized, summed CE loss)."""
        shp = x.shape[:-1]
        flat = x.reshape(-1, 128).float().contiguous()
        tgt = codes.reshape(-1, self.num_quantizers).to(torch.int64).contiguous()
        prep = self._prep()
        own = ops.rvq_encode(flat, self.codebooks, prep)
        loss = ops.rvq_ce(flat, self.codebooks, prep[1], own, tgt)
        emb = ops.rvq_decode(own, self.codebooks)
        return emb.view(*shp, 128), loss
