"""torch-tensor front end of the C ABI (include/ns2_b200.h).

Every function takes CUDA tensors, validates what the kernels assume (dtype, contiguity of the channel
dimension, alignment) and enqueues ONE library call on the current torch CUDA stream.  Nothing here computes
on the host or falls back to PyTorch math: if the library is missing, `_lib.load()` raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (AttnArgs, GemmArgs, NS2_EPI_BF16, NS2_EPI_F32, NS2_EPI_GEGLU, NS2_EPI_WAVENET,
                   NS2_MSE_SCRATCH_PER_SAMPLE, check)

EPI_BF16, EPI_F32, EPI_GEGLU, EPI_WAVENET = NS2_EPI_BF16, NS2_EPI_F32, NS2_EPI_GEGLU, NS2_EPI_WAVENET

Seg = Tuple[int, int, int, int, int]  # (a_col_off, b_col_off, k_len, shift_units, acc)


def _stream(t: Optional[torch.Tensor] = None) -> int:
    """Raw handle of torch's current stream.  The library launches on the CURRENT device (tensor maps, kernel
    attributes and the SM count are per device), so a tensor that lives elsewhere is rejected instead of being
    launched on the wrong GPU — wrap the call in `torch.cuda.device(t.device)`."""
    if t is not None and t.device.index != torch.cuda.current_device():
        raise ValueError(f"tensor is on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}")
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the ns2_b200 ops have no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() > 0 and t.shape[-1] > 1 and t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")


def set_sm_limit(sms: int) -> int:
    """Size every kernel's persistent grid for at most `sms` SMs (0 = all).  Returns the previous limit."""
    return int(_lib.load().ns2_set_sm_limit(int(sms)))


def launch_count() -> int:
    return int(_lib.load().ns2_launch_count())


# --------------------------------------------------------------------------------------------------
# GEMM family
# --------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, n: int, epilogue: int,
         segs: Optional[Sequence[Seg]] = None, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, film: Optional[torch.Tensor] = None,
         film_group_stride: int = 0, bias1_off: int = 0, groups: int = 1,
         a_group_col_stride: int = 0, b_group_row_stride: int = 0, out_group_col_stride: int = 0,
         dil: Optional[Sequence[int]] = None, flags: int = 0,
         debug_timeline: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(segmented_gemm(a, w)).  `a`: (batches, rows, cols) bf16 (may be a strided view),
    `w`: packed bf16 weight (rows, K).  See include/ns2_b200.h section 1 for the exact semantics."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    if a.dim() != 3 or w.dim() != 2:
        raise ValueError("a must be (batches, rows, cols) and w (rows, K)")
    out_dtype = torch.float32 if epilogue == EPI_F32 else torch.bfloat16
    _req(out, out_dtype, "out")
    if out.dim() != 3 or out.shape[0] != a.shape[0] or out.shape[1] != a.shape[1]:
        raise ValueError(f"out must be (batches, rows, *), got {tuple(out.shape)} for a {tuple(a.shape)}")
    if out.shape[0] > 1 and out.stride(0) != out.shape[1] * out.stride(1):
        raise ValueError("out rows must be uniformly strided across batches")
    if segs is None:
        segs = [(0, 0, a.shape[2], 0, 0)]
    args = GemmArgs()
    args.A = a.data_ptr()
    args.a_row_stride, args.a_batch_stride = a.stride(1), a.stride(0)
    args.a_batches, args.a_rows, args.a_cols = a.shape
    args.B = w.data_ptr()
    args.b_row_stride = w.stride(0)
    args.b_rows, args.b_cols = w.shape
    args.n = n
    args.groups = groups
    args.a_group_col_stride = a_group_col_stride
    args.b_group_row_stride = b_group_row_stride
    args.out_group_col_stride = out_group_col_stride
    for g in range(_lib.NS2_GEMM_MAX_GROUPS):
        args.dil[g] = int(dil[g]) if dil is not None and g < len(dil) else 1
    args.num_segs = len(segs)
    for i, s in enumerate(segs):
        sg = args.segs[i]
        sg.a_col_off, sg.b_col_off, sg.k_len, sg.shift_units, sg.acc = (int(v) for v in s)
    args.epilogue = epilogue
    if bias is not None:
        _req(bias, torch.float32, "bias")
    args.bias = _ptr(bias)
    args.bias1_off = bias1_off
    args.out = out.data_ptr()
    args.out_row_stride = out.stride(1)
    if resid is not None:
        _req(resid, torch.float32, "resid")
        if resid.shape != out.shape or (resid.shape[0] > 1 and resid.stride(0) != resid.shape[1] * resid.stride(1)):
            raise ValueError("resid must match out's shape with uniformly strided rows")
        args.resid_row_stride = resid.stride(1)
    args.resid = _ptr(resid)
    if film is not None:
        _req(film, torch.float32, "film")
        args.film_batch_stride = film.stride(0)
    args.film = _ptr(film)
    args.film_group_stride = film_group_stride
    args.flags = int(flags)
    args.debug_timeline = _ptr(debug_timeline)
    check(lib.ns2_gemm(C.byref(args), _stream(out)), "ns2_gemm")
    return out


def wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, *, n: int, k: int, shift_units: int = 0,
          x_col_off: int = 0, groups: int = 1, dy_group_col_stride: int = 0, x_group_col_stride: int = 0,
          dw_group_row_stride: int = 0, dil: Optional[Sequence[int]] = None, splits: int = 0) -> torch.Tensor:
    """dw[g][:n, :k] += dy[..., g-th n columns]^T @ x[..., rows shifted by shift_units*dil[g], g-th k columns].
    dy, x: (batches, rows, cols) bf16 (strided views are fine); dw: fp32 2-D (rows >= groups' n rows, cols >= k)."""
    lib = _lib.load()
    _req(dy, torch.bfloat16, "dy")
    _req(x, torch.bfloat16, "x")
    _req(dw, torch.float32, "dw")
    if dy.dim() != 3 or x.dim() != 3 or dy.shape[:2] != x.shape[:2] or dw.dim() != 2:
        raise ValueError("wgrad: dy and x must be (batches, rows, cols) with equal leading dims; dw 2-D")
    if groups > 1 and n % 128 != 0:
        raise ValueError("wgrad: n must be a multiple of 128 for grouped weights")
    args = _lib.WgradArgs()
    args.dY, args.dy_row_stride, args.dy_batch_stride, args.dy_cols = dy.data_ptr(), dy.stride(1), dy.stride(0), dy.shape[2]
    args.X, args.x_row_stride, args.x_batch_stride, args.x_cols = x.data_ptr(), x.stride(1), x.stride(0), x.shape[2]
    args.batches, args.rows = dy.shape[0], dy.shape[1]
    args.n, args.k = n, k
    args.groups = groups
    args.dy_group_col_stride, args.x_group_col_stride, args.x_col_off = dy_group_col_stride, x_group_col_stride, x_col_off
    for g in range(_lib.NS2_GEMM_MAX_GROUPS):
        args.dil[g] = int(dil[g]) if dil is not None and g < len(dil) else 1
    args.shift_units = shift_units
    args.dW, args.dw_row_stride, args.dw_group_row_stride = dw.data_ptr(), dw.stride(0), dw_group_row_stride
    args.splits = splits
    check(lib.ns2_wgrad(C.byref(args), _stream(dw)), "ns2_wgrad")
    return dw


def conv3_segs(k_len: int, tap_stride: Optional[int] = None, acc: int = 0, a_col_off: int = 0,
               b_col_off: int = 0) -> list:
    """Segments of a causal k=3 conv whose packed weight holds tap t at columns [t*tap_stride, +k_len):
    tap t multiplies x[n - (2 - t) * dilation]   (CausalConv1d, ns2.py:583-595)."""
    ts = k_len if tap_stride is None else tap_stride
    return [(a_col_off, b_col_off + t * ts, k_len, 2 - t, acc) for t in range(3)]


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
ATTN_AUTO, ATTN_ONE_TILE, ATTN_TWO_TILE, ATTN_TWO_TILE_POLY2, ATTN_TWO_TILE_POLY4, ATTN_TWO_TILE_LOCKSTEP = 0, 1, 2, 3, 4, 5


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, heads: int,
              scale: Optional[float] = None, kernel: int = ATTN_AUTO,
              debug_timeline: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q: (B, Nq, heads*64), k/v: (B, Nk, heads*64) bf16 (strided views into a fused projection are fine)."""
    lib = _lib.load()
    for name, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _req(t, torch.bfloat16, name)
        if t.dim() != 3 or t.shape[2] != heads * 64:
            raise ValueError(f"{name} must be (B, N, heads*64), got {tuple(t.shape)}")
    args = AttnArgs()
    args.q, args.q_row_stride, args.q_batch_stride = q.data_ptr(), q.stride(1), q.stride(0)
    args.k, args.k_row_stride, args.k_batch_stride = k.data_ptr(), k.stride(1), k.stride(0)
    args.v, args.v_row_stride, args.v_batch_stride = v.data_ptr(), v.stride(1), v.stride(0)
    args.out, args.o_row_stride, args.o_batch_stride = out.data_ptr(), out.stride(1), out.stride(0)
    args.batches, args.heads = q.shape[0], heads
    args.q_len, args.kv_len, args.dim_head = q.shape[1], k.shape[1], 64
    args.scale = float(scale if scale is not None else 64 ** -0.5)
    args.kernel = int(kernel)
    args.debug_timeline = _ptr(debug_timeline)
    if lse is not None:
        _req(lse, torch.float32, "lse")
        if not lse.is_contiguous() or tuple(lse.shape) != (q.shape[0], heads, q.shape[1]):
            raise ValueError("lse must be a contiguous (B, heads, Nq) float tensor")
    args.lse = _ptr(lse)
    check(lib.ns2_attn_fwd(C.byref(args), _stream(out)), "ns2_attn_fwd")
    return out


# --------------------------------------------------------------------------------------------------
# norms, small layers, casts
# --------------------------------------------------------------------------------------------------
def rmsnorm_film(x: torch.Tensor, out: torch.Tensor, *, gamma: Optional[torch.Tensor] = None,
                 film: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: (B, N, D) f32 -> out (B, N, D) bf16.  film: (B, >=2D) f32 view whose row b holds [gamma_b | beta_b]."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    if not (x.is_contiguous() and out.is_contiguous()):
        raise ValueError("x and out must be contiguous")
    B, N, D = x.shape
    if gamma is not None:
        _req(gamma, torch.float32, "gamma")
    film_bs = 0
    if film is not None:
        _req(film, torch.float32, "film")
        film_bs = film.stride(0)
    check(lib.ns2_rmsnorm_film(x.data_ptr(), D, B * N, D, N, _ptr(gamma), _ptr(film), film_bs,
                               out.data_ptr(), D, _stream(out)), "ns2_rmsnorm_film")
    return out


def rmsnorm_f32(x: torch.Tensor, out: torch.Tensor, gamma: Optional[torch.Tensor]) -> torch.Tensor:
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(out, torch.float32, "out")
    D = x.shape[-1]
    rows = x.numel() // D
    check(lib.ns2_rmsnorm_f32(x.data_ptr(), D, rows, D, _ptr(gamma), out.data_ptr(), D, _stream()),
          "ns2_rmsnorm_f32")
    return out


_SMALL_BATCH_MAX = 64          # kMaxSmallBatch in csrc/elementwise.cu
_SMALL_SMEM_BYTES = 200 * 1024  # dynamic shared memory the small-layer kernel may use for its (batch, k) input tile


def _small_chunk(k: int) -> int:
    return max(1, min(_SMALL_BATCH_MAX, _SMALL_SMEM_BYTES // (4 * k)))


def time_cond(times: torch.Tensor, freqs: torch.Tensor, w: torch.Tensor, bias: torch.Tensor,
              out: torch.Tensor) -> torch.Tensor:
    """out[b] = silu(W @ [t_b, sin(2 pi t_b f), cos(2 pi t_b f)] + bias); out may be a column slice.
    Batches larger than the kernel's per-launch limit are processed in row chunks."""
    lib = _lib.load()
    for name, t in (("times", times), ("freqs", freqs), ("w", w), ("bias", bias), ("out", out)):
        _req(t, torch.float32, name)
    step = _small_chunk(2 * freqs.shape[0] + 1)
    for b0 in range(0, times.shape[0], step):
        tb, ob = times[b0:b0 + step], out[b0:b0 + step]
        check(lib.ns2_time_cond(tb.data_ptr(), tb.shape[0], freqs.data_ptr(), freqs.shape[0], w.data_ptr(),
                                bias.data_ptr(), w.shape[0], ob.data_ptr(), out.stride(0), _stream()),
              "ns2_time_cond")
    return out


def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor,
                 act: int = 0) -> torch.Tensor:
    lib = _lib.load()
    for name, t in (("x", x), ("w", w), ("out", out)):
        _req(t, torch.float32, name)
    step = _small_chunk(x.shape[1])
    for b0 in range(0, x.shape[0], step):
        xb, ob = x[b0:b0 + step], out[b0:b0 + step]
        check(lib.ns2_small_linear(xb.data_ptr(), x.stride(0), xb.shape[0], x.shape[1], w.data_ptr(), _ptr(bias),
                                   w.shape[0], act, ob.data_ptr(), out.stride(0), _stream()), "ns2_small_linear")
    return out


def cast_bf16(x: torch.Tensor, out: torch.Tensor, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    if not (x.is_contiguous() and out.is_contiguous()) or x.numel() != out.numel():
        raise ValueError("cast_bf16 needs contiguous tensors of equal size")
    if add is not None:
        _req(add, torch.float32, "add")
        if not add.is_contiguous() or add.numel() != x.numel():
            raise ValueError("add must be contiguous and the same size as x")
    check(lib.ns2_cast_bf16(x.data_ptr(), _ptr(add), x.numel(), out.data_ptr(), _stream()),
          "ns2_cast_bf16")
    return out


def cond_inject(x: torch.Tensor, cproj: torch.Tensor, out: torch.Tensor, drop_mask: Optional[torch.Tensor] = None,
                null_cond: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (B, N, D) bf16 = x (B, N, D) f32 + [padded / curtailed, null-substituted] cproj (B, L, D) f32."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(cproj, torch.float32, "cproj")
    _req(out, torch.bfloat16, "out")
    B, N, D = x.shape
    if not (x.is_contiguous() and cproj.is_contiguous() and out.is_contiguous()) or cproj.shape[0] != B \
            or cproj.shape[2] != D or out.shape != x.shape:
        raise ValueError("cond_inject: x/out (B, N, D) and cproj (B, L, D) must be contiguous and consistent")
    if drop_mask is not None:
        if drop_mask.dtype != torch.bool or drop_mask.numel() != B or not drop_mask.is_cuda:
            raise ValueError("drop_mask must be a CUDA bool tensor of B elements")
        _req(null_cond, torch.float32, "null_cond")
        if null_cond.numel() != D or not null_cond.is_contiguous():
            raise ValueError("null_cond must be a contiguous (D,) float tensor")
    check(lib.ns2_cond_inject(x.data_ptr(), cproj.data_ptr(), _ptr(drop_mask), _ptr(null_cond), B, N, cproj.shape[1], D,
                              out.data_ptr(), _stream(out)), "ns2_cond_inject")
    return out


def select_rows(drop_mask: torch.Tensor, null_row: torch.Tensor, src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[b] = null_row if drop_mask[b] else src[b]; src/out are (B, ...) with contiguous trailing dims; out may be a
    column slice of a wider f32 matrix (row stride > row length) or a bf16 tensor."""
    lib = _lib.load()
    _req(null_row, torch.float32, "null_row")
    _req(src, torch.float32, "src")
    B = src.shape[0]
    row_len = src.numel() // B
    if drop_mask.dtype != torch.bool or drop_mask.numel() != B or not drop_mask.is_cuda:
        raise ValueError("drop_mask must be a CUDA bool tensor of B elements")
    if null_row.numel() != row_len or not null_row.is_contiguous() or not src.is_contiguous():
        raise ValueError("null_row must hold one row; src must be contiguous")
    if out.dtype not in (torch.float32, torch.bfloat16) or out.shape[0] != B or out.numel() != B * row_len:
        raise ValueError("out must be (B, ...) f32/bf16 with src's row length")
    if out.dim() > 2 and not out.is_contiguous():
        raise ValueError("multi-dimensional out must be contiguous")
    check(lib.ns2_select_rows(drop_mask.data_ptr(), null_row.data_ptr(), src.data_ptr(), row_len, B, row_len,
                              out.data_ptr(), out.stride(0), int(out.dtype == torch.bfloat16), _stream(out)),
          "ns2_select_rows")
    return out


def mean_rows(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(out, torch.float32, "out")
    B, N, D = x.shape
    check(lib.ns2_mean_rows(x.contiguous().data_ptr(), B, N, D, out.data_ptr(), _stream()),
          "ns2_mean_rows")
    return out


def transpose_cast(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """(B, C, L) f32 channel-first -> (B, L, C) bf16."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    B, Cc, L = x.shape
    check(lib.ns2_transpose_cast(x.contiguous().data_ptr(), B, Cc, L, out.data_ptr(), _stream()),
          "ns2_transpose_cast")
    return out


# --------------------------------------------------------------------------------------------------
# diffusion element-wise
# --------------------------------------------------------------------------------------------------
OBJECTIVES = {"v": _lib.NS2_OBJ_V, "eps": _lib.NS2_OBJ_EPS, "x0": _lib.NS2_OBJ_X0}


def groupnorm_silu(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, groups: int, *, eps: float = 1e-5,
                   resid: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
                   out_bf16: Optional[torch.Tensor] = None):
    """silu(GroupNorm(groups)(x)) (+ resid) for token-major x (B, N, C) f32 -> out_f32 and/or out_bf16 (B, N, C)."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(weight, torch.float32, "weight")
    _req(bias, torch.float32, "bias")
    if x.dim() != 3 or not x.is_contiguous():
        raise ValueError("x must be a contiguous (B, N, C) tensor")
    for name, t, dt in (("resid", resid, torch.float32), ("out_f32", out_f32, torch.float32),
                        ("out_bf16", out_bf16, torch.bfloat16)):
        if t is not None:
            _req(t, dt, name)
            if t.shape != x.shape or not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous with x's shape")
    if out_f32 is None and out_bf16 is None:
        raise ValueError("groupnorm_silu needs at least one output")
    B, N, Cn = x.shape
    check(lib.ns2_groupnorm_silu(x.data_ptr(), B, N, Cn, int(groups), weight.data_ptr(), bias.data_ptr(), float(eps),
                                 _ptr(resid), _ptr(out_f32), _ptr(out_bf16), _stream(x)), "ns2_groupnorm_silu")
    return out_f32, out_bf16


def rowdot(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, relu: bool = False):
    """out[...] = (relu)(x[..., :] . w + bias): Linear(dim, 1) heads.  x f32 contiguous, w (dim,), out one value per row."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(w, torch.float32, "w")
    _req(out, torch.float32, "out")
    if not (x.is_contiguous() and w.is_contiguous() and out.is_contiguous()) or out.numel() * x.shape[-1] != x.numel():
        raise ValueError("rowdot needs contiguous x (..., dim), w (dim,) and one output per row")
    if bias is not None:
        _req(bias, torch.float32, "bias")
    check(lib.ns2_rowdot(x.data_ptr(), out.numel(), x.shape[-1], w.data_ptr(), _ptr(bias), int(relu), out.data_ptr(),
                         _stream(x)), "ns2_rowdot")
    return out


def expand_encodings(phon: torch.Tensor, coarse: torch.Tensor, pitch_table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(B, D, L) f32 channel-first: phon[b, idx[b, n], :] + pitch_table[coarse[b, idx[b, n]], :], 0 where idx < 0
    (expand_encodings, ns2.py:1449-1455).  phon (B, T, D) f32, coarse (B, T) int32, idx (B, L) int32."""
    lib = _lib.load()
    _req(phon, torch.float32, "phon")
    _req(pitch_table, torch.float32, "pitch_table")
    _req(coarse, torch.int32, "coarse")
    _req(idx, torch.int32, "idx")
    if not all(t.is_contiguous() for t in (phon, coarse, pitch_table, idx)):
        raise ValueError("expand_encodings needs contiguous tensors")
    B, T, D = phon.shape
    if coarse.shape != (B, T) or idx.dim() != 2 or idx.shape[0] != B or pitch_table.shape[1] != D:
        raise ValueError("expand_encodings: inconsistent shapes")
    L = idx.shape[1]
    out = torch.empty(B, D, L, device=phon.device, dtype=torch.float32)
    check(lib.ns2_expand_encodings(phon.data_ptr(), coarse.data_ptr(), pitch_table.data_ptr(), pitch_table.shape[0],
                                   idx.data_ptr(), B, T, D, L, out.data_ptr(), _stream(phon)), "ns2_expand_encodings")
    return out


def embedding_bf16(ids: torch.Tensor, table: torch.Tensor, out: torch.Tensor, pad_id: int) -> torch.Tensor:
    """out[..., :] = bf16(table[ids < 0 ? pad_id : ids]) — nn.Embedding + padding substitution (ns2.py:279-282)."""
    lib = _lib.load()
    _req(ids, torch.int64, "ids")
    _req(table, torch.float32, "table")
    _req(out, torch.bfloat16, "out")
    if not (ids.is_contiguous() and table.is_contiguous() and out.is_contiguous()):
        raise ValueError("embedding_bf16 needs contiguous tensors")
    if out.numel() != ids.numel() * table.shape[1]:
        raise ValueError("out must hold one table row per id")
    check(lib.ns2_embedding_bf16(ids.data_ptr(), ids.numel(), table.data_ptr(), table.shape[0], table.shape[1],
                                 int(pad_id), out.data_ptr(), _stream(out)), "ns2_embedding_bf16")
    return out


def q_sample(x0, noise, alpha, sigma, x_t, target=None, objective: str = "v"):
    """x_t = alpha x0 + sigma noise; target of the chosen parameterisation (ns2.py:1631-1644)."""
    lib = _lib.load()
    B = x0.shape[0]
    per = x0.numel() // B
    for name, t in (("x0", x0), ("noise", noise), ("alpha", alpha), ("sigma", sigma), ("x_t", x_t)):
        _req(t, torch.float32, name)
    if target is not None:
        _req(target, torch.float32, "target")
    check(lib.ns2_q_sample(x0.data_ptr(), noise.data_ptr(), alpha.data_ptr(), sigma.data_ptr(), B, per,
                           x_t.data_ptr(), _ptr(target), OBJECTIVES[objective], _stream()), "ns2_q_sample")
    return x_t, target


def mse_rows(pred, target, out, scratch=None, mean_out=None):
    """out[b] = mean((pred[b] - target[b])^2); `mean_out` (0-d / 1-element f32) additionally receives out.mean()."""
    lib = _lib.load()
    B = pred.shape[0]
    per = pred.numel() // B
    for name, t in (("pred", pred), ("target", target), ("out", out)):
        _req(t, torch.float32, name)
    if not (pred.is_contiguous() and target.is_contiguous()):
        raise ValueError("pred and target must be contiguous")
    if scratch is None:
        scratch = torch.empty(B * NS2_MSE_SCRATCH_PER_SAMPLE, device=pred.device, dtype=torch.float32)
    if mean_out is not None:
        _req(mean_out, torch.float32, "mean_out")
    check(lib.ns2_mse_rows(pred.data_ptr(), target.data_ptr(), B, per, scratch.data_ptr(),
                           out.data_ptr(), _ptr(mean_out), _stream()), "ns2_mse_rows")
    return out


def ddim_step(x, v, alpha, sigma, alpha_next, sigma_next, objective: str = "v"):
    """In-place DDIM update of x from the model output `v` (ns2.py:1412-1429)."""
    lib = _lib.load()
    B = x.shape[0]
    per = x.numel() // B
    check(lib.ns2_ddim_step(x.data_ptr(), v.data_ptr(), alpha.data_ptr(), sigma.data_ptr(),
                            alpha_next.data_ptr(), sigma_next.data_ptr(), B, per, OBJECTIVES[objective],
                            _stream()),
          "ns2_ddim_step")
    return x


def x_start_from_pred(x, pred, alpha, sigma, out, objective: str = "v"):
    """x_start implied by the model output under the chosen parameterisation (ns2.py:1673-1680)."""
    lib = _lib.load()
    B = x.shape[0]
    per = x.numel() // B
    for name, t in (("x", x), ("pred", pred), ("alpha", alpha), ("sigma", sigma), ("out", out)):
        _req(t, torch.float32, name)
    check(lib.ns2_x_start(x.data_ptr(), pred.data_ptr(), alpha.data_ptr(), sigma.data_ptr(), B, per, out.data_ptr(),
                          OBJECTIVES[objective], _stream(out)), "ns2_x_start")
    return out


def cfg_combine(cond, null, scale, out):
    lib = _lib.load()
    check(lib.ns2_cfg_combine(cond.data_ptr(), null.data_ptr(), float(scale), cond.numel(),
                              out.data_ptr(), _stream()), "ns2_cfg_combine")
    return out


# --------------------------------------------------------------------------------------------------
# RVQ
# --------------------------------------------------------------------------------------------------
def rvq_prepare(codebooks: torch.Tensor):
    """codebooks (Q, K, 128) f32 -> (fp16 copy, ||c||^2 (Q, K) f32, meta (Q, 2) f32)."""
    lib = _lib.load()
    _req(codebooks, torch.float32, "codebooks")
    cb = codebooks.contiguous()
    Q, K, D = cb.shape
    # fp16 copy (Q, K, D) followed by the (Q, K, 16) norm blocks: NS2_RVQ_PREPARED_HALFS
    cb16 = torch.empty((Q * K * (D + 16),), device=cb.device, dtype=torch.float16)
    cn2 = torch.empty((Q, K), device=cb.device, dtype=torch.float32)
    meta = torch.empty((Q, 2), device=cb.device, dtype=torch.float32)
    check(lib.ns2_rvq_prepare(cb.data_ptr(), Q, K, D, cb16.data_ptr(), cn2.data_ptr(), meta.data_ptr(),
                              _stream()), "ns2_rvq_prepare")
    return cb16, cn2, meta


def rvq_encode(frames: torch.Tensor, codebooks: torch.Tensor, prepared, codes: Optional[torch.Tensor] = None,
               stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frames (F, 128) f32 -> codes (F, Q) int64."""
    lib = _lib.load()
    _req(frames, torch.float32, "frames")
    cb16, cn2, meta = prepared
    cb = codebooks.contiguous()
    Q, K, D = cb.shape
    fr = frames.contiguous()
    F = fr.shape[0]
    if codes is None:
        codes = torch.empty((F, Q), device=fr.device, dtype=torch.int64)
    if stats is not None and not (stats.is_cuda and stats.dtype == torch.int64 and stats.is_contiguous()
                                  and stats.numel() >= _lib.NS2_RVQ_STATS_LEN):
        raise ValueError(f"stats must be a contiguous CUDA int64 tensor with >= {_lib.NS2_RVQ_STATS_LEN} elements")
    check(lib.ns2_rvq_encode(fr.data_ptr(), F, D, cb.data_ptr(), cb16.data_ptr(), cn2.data_ptr(),
                             meta.data_ptr(), Q, K, codes.data_ptr(), _ptr(stats), _stream(codes)),
          "ns2_rvq_encode")
    return codes


def rvq_decode(codes: torch.Tensor, codebooks: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    cb = codebooks.contiguous()
    Q, K, D = cb.shape
    cd = codes.contiguous()
    F = cd.shape[0]
    if out is None:
        out = torch.empty((F, D), device=cb.device, dtype=torch.float32)
    check(lib.ns2_rvq_decode(cd.data_ptr(), F, Q, K, D, cb.data_ptr(), out.data_ptr(), _stream()),
          "ns2_rvq_decode")
    return out


def rvq_ce(frames: torch.Tensor, codebooks: torch.Tensor, cn2: torch.Tensor, own_codes: torch.Tensor,
           target_codes: torch.Tensor) -> torch.Tensor:
    """Cross-entropy head of the residual VQ (`codec.rq`): frames (F, 128) f32, codes (F, Q) int64 -> 0-d loss."""
    lib = _lib.load()
    _req(frames, torch.float32, "frames")
    cb = codebooks.contiguous()
    Q, K, D = cb.shape
    fr = frames.contiguous()
    F = fr.shape[0]
    for name, t in (("own_codes", own_codes), ("target_codes", target_codes)):
        if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and tuple(t.shape) == (F, Q)):
            raise ValueError(f"{name} must be a contiguous CUDA int64 tensor of shape (F, Q)")
    scratch = torch.empty(F * Q, device=fr.device, dtype=torch.float32)
    loss = torch.empty((), device=fr.device, dtype=torch.float32)
    check(lib.ns2_rvq_ce(fr.data_ptr(), F, D, cb.data_ptr(), cn2.data_ptr(), Q, K, own_codes.data_ptr(),
                         target_codes.data_ptr(), scratch.data_ptr(), loss.data_ptr(), _stream(fr)), "ns2_rvq_ce")
    return loss


# --------------------------------------------------------------------------------------------------
# backward pass
# --------------------------------------------------------------------------------------------------
def attention_bwd(q, k, v, o, d_o, lse, dq_accum, dk, dv, *, heads: int, scale: Optional[float] = None,
                  delta: Optional[torch.Tensor] = None):
    """(dq_accum f32 (B, Nq, inner) += dQ, dk, dv bf16) of softmax(q k^T scale) v given d_o; dq_accum must be zeroed."""
    lib = _lib.load()
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o), ("d_o", d_o), ("dk", dk), ("dv", dv)):
        _req(t, torch.bfloat16, name)
        if t.dim() != 3 or t.shape[2] != heads * 64:
            raise ValueError(f"{name} must be (B, N, heads*64), got {tuple(t.shape)}")
    _req(lse, torch.float32, "lse")
    _req(dq_accum, torch.float32, "dq_accum")
    B, Nq, Nk = q.shape[0], q.shape[1], k.shape[1]
    if not dq_accum.is_contiguous() or tuple(dq_accum.shape) != (B, Nq, heads * 64):
        raise ValueError("dq_accum must be contiguous (B, Nq, heads*64) float32")
    if delta is None:
        delta = torch.empty(B, heads, Nq, device=q.device, dtype=torch.float32)
    a = _lib.AttnBwdArgs()
    a.q, a.q_row_stride, a.q_batch_stride = q.data_ptr(), q.stride(1), q.stride(0)
    a.k, a.k_row_stride, a.k_batch_stride = k.data_ptr(), k.stride(1), k.stride(0)
    a.v, a.v_row_stride, a.v_batch_stride = v.data_ptr(), v.stride(1), v.stride(0)
    a.o, a.o_row_stride, a.o_batch_stride = o.data_ptr(), o.stride(1), o.stride(0)
    a.d_o, a.do_row_stride, a.do_batch_stride = d_o.data_ptr(), d_o.stride(1), d_o.stride(0)
    a.lse, a.delta, a.dq_accum = lse.data_ptr(), delta.data_ptr(), dq_accum.data_ptr()
    a.dk, a.dk_row_stride, a.dk_batch_stride = dk.data_ptr(), dk.stride(1), dk.stride(0)
    a.dv, a.dv_row_stride, a.dv_batch_stride = dv.data_ptr(), dv.stride(1), dv.stride(0)
    a.batches, a.heads, a.q_len, a.kv_len, a.dim_head = B, heads, Nq, Nk, 64
    a.scale = float(scale if scale is not None else 64 ** -0.5)
    check(lib.ns2_attn_bwd(C.byref(a), _stream(dq_accum)), "ns2_attn_bwd")
    return dq_accum, dk, dv


def rmsnorm_film_bwd(x, dh, dxr, dxr_bf, *, rows_per_batch: int, gamma=None, film=None, dfilm=None, dgamma=None):
    """dxr (f32, in place) += d/dx of rmsnorm_film(x) given dh (bf16); dxr_bf = bf16(dxr); dfilm / dgamma accumulate."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(dh, torch.bfloat16, "dh")
    _req(dxr, torch.float32, "dxr")
    _req(dxr_bf, torch.bfloat16, "dxr_bf")
    for t in (x, dh, dxr, dxr_bf):
        if not t.is_contiguous():
            raise ValueError("rmsnorm_film_bwd needs contiguous tensors")
    D = x.shape[-1]
    rows = x.numel() // D
    film_bs = dfilm_bs = 0
    if film is not None:
        _req(film, torch.float32, "film")
        _req(dfilm, torch.float32, "dfilm")
        film_bs, dfilm_bs = film.stride(0), dfilm.stride(0)
    check(lib.ns2_rmsnorm_film_bwd(x.data_ptr(), dh.data_ptr(), rows, D, rows_per_batch, _ptr(gamma), _ptr(film), film_bs,
                                   _ptr(dfilm), dfilm_bs, _ptr(dgamma), dxr.data_ptr(), dxr_bf.data_ptr(), _stream(dxr)),
          "ns2_rmsnorm_film_bwd")
    return dxr


def geglu_bwd(pre, dg):
    """pre (rows, 2*Dp) bf16 packed [128 value | 128 gate] tiles -> overwritten by its gradient given dg (rows, Dp)."""
    lib = _lib.load()
    _req(pre, torch.bfloat16, "pre")
    _req(dg, torch.bfloat16, "dg")
    if not (pre.is_contiguous() and dg.is_contiguous()) or pre.shape[-1] != 2 * dg.shape[-1]:
        raise ValueError("geglu_bwd: pre (.., 2*Dp) and dg (.., Dp) must be contiguous")
    dp = dg.shape[-1]
    check(lib.ns2_geglu_bwd(pre.data_ptr(), dg.data_ptr(), dg.numel() // dp, dp, _stream(pre)), "ns2_geglu_bwd")
    return pre


def wavenet_gate_bwd(c, dy, dc, film, dfilm, *, dim: int, groups: int, film_group_stride: int):
    """c, dy, dc: (B, N, >= groups*dim) bf16 views (first groups*dim columns used); film/dfilm (B, ...) f32 views whose
    row b holds, for group g at g*film_group_stride, [gamma | beta]."""
    lib = _lib.load()
    for name, t in (("c", c), ("dy", dy), ("dc", dc)):
        _req(t, torch.bfloat16, name)
        if t.dim() != 3 or t.stride(0) != t.shape[1] * t.stride(1):
            raise ValueError(f"{name} must be (B, N, cols) with uniformly strided rows")
    _req(film, torch.float32, "film")
    _req(dfilm, torch.float32, "dfilm")
    B, N = c.shape[0], c.shape[1]
    check(lib.ns2_wavenet_gate_bwd(c.data_ptr(), c.stride(1), dy.data_ptr(), dy.stride(1), dc.data_ptr(), dc.stride(1), B,
                                   N, dim, groups, film.data_ptr(), film.stride(0), film_group_stride, dfilm.data_ptr(),
                                   dfilm.stride(0), _stream(dc)), "ns2_wavenet_gate_bwd")
    return dc


def colsum(t, out):
    """out[c] (f32) += sum over all leading dims of t[..., c] (bf16; last dim contiguous, uniform row stride)."""
    lib = _lib.load()
    _req(t, torch.bfloat16, "t")
    _req(out, torch.float32, "out")
    cols = t.shape[-1]
    rows = t.numel() // cols
    rs = t.stride(-2) if t.dim() >= 2 else cols
    if t.dim() == 3 and t.stride(0) != t.shape[1] * t.stride(1):
        raise ValueError("colsum: rows must be uniformly strided")
    check(lib.ns2_colsum_bf16(t.data_ptr(), rows, cols, rs, out.data_ptr(), _stream(out)), "ns2_colsum_bf16")
    return out


def group_sum(t, out, *, dim: int, groups: int):
    lib = _lib.load()
    _req(t, torch.bfloat16, "t")
    _req(out, torch.bfloat16, "out")
    if not (t.is_contiguous() and out.is_contiguous()):
        raise ValueError("group_sum needs contiguous tensors")
    check(lib.ns2_group_sum_bf16(t.data_ptr(), out.numel() // dim, dim, groups, out.data_ptr(), _stream(out)),
          "ns2_group_sum_bf16")
    return out


def mse_bwd(pred, target, coef, out_bf=None, out_f32=None):
    """coef[b] * (pred - target) as bf16 and/or f32: the seed of the backward pass."""
    lib = _lib.load()
    for name, t in (("pred", pred), ("target", target), ("coef", coef)):
        _req(t, torch.float32, name)
    if out_bf is not None:
        _req(out_bf, torch.bfloat16, "out_bf")
    if out_f32 is not None:
        _req(out_f32, torch.float32, "out_f32")
    B = pred.shape[0]
    check(lib.ns2_mse_bwd(pred.data_ptr(), target.data_ptr(), coef.data_ptr(), B, pred.numel() // B, _ptr(out_bf),
                          _ptr(out_f32), _stream(pred)), "ns2_mse_bwd")
    return out_bf if out_bf is not None else out_f32


def film_wgrad(dfilm, t, dw, accumulate: bool = True):
    """dw (rows, cols) f32 (+)= dfilm (B, rows)^T @ t (B, cols).  accumulate=False overwrites dw (which then need not be
    initialised: one pass over the gradient buffer instead of zero-fill + read-modify-write).  `dfilm` may be a column
    window of a wider (B, total_rows) buffer (unit column stride)."""
    lib = _lib.load()
    for name, x in (("dfilm", dfilm), ("t", t), ("dw", dw)):
        _req(x, torch.float32, name)
    if not (t.is_contiguous() and dw.is_contiguous()) or dfilm.dim() != 2 or (dfilm.shape[1] > 1 and dfilm.stride(1) != 1):
        raise ValueError("t and dw must be contiguous, dfilm (B, rows) with unit column stride")
    B, rows = dfilm.shape
    if tuple(dw.shape) != (rows, t.shape[1]) or t.shape[0] != B:
        raise ValueError("film_wgrad: inconsistent shapes")
    for b0 in range(0, B, 32):
        check(lib.ns2_film_wgrad(dfilm[b0:b0 + 32].data_ptr(), dfilm.stride(0), t[b0:b0 + 32].data_ptr(), min(32, B - b0),
                                 rows, t.shape[1], dw.data_ptr(), int(accumulate or b0 > 0), _stream(dw)),
              "ns2_film_wgrad")
    return dw


def accum_bf16(acc, t, acc_bf=None):
    """acc (f32, contiguous) += t (bf16, contiguous, same numel); acc_bf (optional) = bf16(acc)."""
    lib = _lib.load()
    _req(acc, torch.float32, "acc")
    _req(t, torch.bfloat16, "t")
    if not (acc.is_contiguous() and t.is_contiguous()) or acc.numel() != t.numel():
        raise ValueError("accum_bf16 needs contiguous tensors of equal size")
    if acc_bf is not None:
        _req(acc_bf, torch.bfloat16, "acc_bf")
    check(lib.ns2_accum_bf16(acc.data_ptr(), t.data_ptr(), acc.numel(), _ptr(acc_bf), _stream(acc)), "ns2_accum_bf16")
    return acc


# --------------------------------------------------------------------------------------------------
# Monotonic alignment search (aligner.py:88-122)
# --------------------------------------------------------------------------------------------------
def maximum_path(value: torch.Tensor, mask: torch.Tensor, neg_const: float = float("-inf"), *,
                 want_path: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(idx (b, t_y) int32, path (b, t_x, t_y) f32 or None) for value/mask (b, t_x, t_y) f32 contiguous."""
    lib = _lib.load()
    _req(value, torch.float32, "value")
    _req(mask, torch.float32, "mask")
    if value.dim() != 3 or value.shape != mask.shape:
        raise ValueError(f"value and mask must both be (b, t_x, t_y); got {tuple(value.shape)} / {tuple(mask.shape)}")
    if not (value.is_contiguous() and mask.is_contiguous()):
        raise ValueError("value and mask must be contiguous")
    b, t_x, t_y = value.shape
    idx = torch.empty((b, t_y), dtype=torch.int32, device=value.device)
    path = torch.empty_like(value) if want_path else None
    ws_bytes = int(lib.ns2_maximum_path_workspace_bytes(b, t_x, t_y))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=value.device)
    check(lib.ns2_maximum_path(value.data_ptr(), mask.data_ptr(), b, t_x, t_y, float(neg_const), ws.data_ptr(),
                               ws_bytes, idx.data_ptr(), _ptr(path), _stream(value)), "ns2_maximum_path")
    return idx, path
