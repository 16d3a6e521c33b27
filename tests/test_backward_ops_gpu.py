"""GPU: every kernel of the backward pass against torch autograd on the same (bf16-rounded) operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


def _close(got, ref, atol, rtol, name=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    assert torch.isfinite(got).all(), name
    assert bool((err <= tol).all()), f"{name}: max err {float(err.max()):.4e} (ref absmax {float(ref.abs().max()):.3e})"


@pytest.mark.parametrize("B,N,n,k,shift,dil,groups", [
    (2, 200, 128, 256, 0, 1, 1), (3, 333, 256, 192, 2, 4, 1), (2, 1024, 512, 512, 1, 128, 1),
    (2, 160, 128, 128, 2, None, 3), (1, 64, 1408, 1408, 1, 1, 1)])
def test_wgrad_matches_einsum(B, N, n, k, shift, dil, groups):
    from naturalspeech2_pytorch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    dy = torch.randn(B, N, groups * n, device="cuda", generator=g).to(bf)
    x = torch.randn(B, N, groups * k, device="cuda", generator=g).to(bf)
    dils = [2 ** i for i in range(groups)] if dil is None else [dil] * groups
    dw = torch.randn(groups * n, k, device="cuda", generator=g)
    before = dw.clone()
    ops.wgrad(dy, x, dw, n=n, k=k, shift_units=shift, groups=groups, dy_group_col_stride=n, x_group_col_stride=k,
              dw_group_row_stride=n, dil=dils)
    for gi in range(groups):
        s = shift * dils[gi]
        xs = F.pad(x[:, :, gi * k:(gi + 1) * k].float(), (0, 0, s, 0))[:, :N]      # x[m - s]
        ref = torch.einsum("bmn,bmk->nk", dy[:, :, gi * n:(gi + 1) * n].float(), xs)
        _close(dw[gi * n:(gi + 1) * n] - before[gi * n:(gi + 1) * n], ref, 2e-2 * (B * N) ** 0.5 / 10, 1e-2, f"wgrad g{gi}")


def test_conv_dgrad_with_negative_shifts_matches_autograd():
    """dgrad of CausalConv1d = the same segmented GEMM with transposed taps and shifts of the opposite sign."""
    from naturalspeech2_pytorch_b200 import ops
    B, N, C, O, d = 2, 300, 256, 384, 4
    g = torch.Generator(device="cuda").manual_seed(1)
    w = (torch.randn(O, C, 3, device="cuda", generator=g) * 0.05).to(bf).float()
    dy = torch.randn(B, N, O, device="cuda", generator=g).to(bf)
    x = torch.zeros(B, C, N, device="cuda", requires_grad=True)
    y = F.conv1d(F.pad(x, (2 * d, 0)), w, dilation=d)
    y.backward(dy.float().transpose(1, 2))
    wt = torch.cat([w[:, :, t].t() for t in range(3)], dim=1).contiguous().to(bf)     # (C, 3*O): tap t at cols [t*O, ..)
    dx = torch.empty(B, N, C, device="cuda", dtype=bf)
    segs = [(0, t * O, O, -(2 - t), 0) for t in range(3)]
    ops.gemm(dy, wt, dx, n=C, epilogue=ops.EPI_BF16, segs=segs, dil=[d])
    _close(dx, x.grad.transpose(1, 2), 3e-2, 2e-2, "conv dgrad")


@pytest.mark.parametrize("mode", ["film", "gamma"])
def test_rmsnorm_film_bwd_matches_autograd(mode):
    from naturalspeech2_pytorch_b200 import ops
    B, N, D = 3, 150, 512
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(B, N, D, device="cuda", generator=g, requires_grad=True)
    dh = torch.randn(B, N, D, device="cuda", generator=g).to(bf)
    film = (torch.randn(B, 2 * D, device="cuda", generator=g) * 0.5 + 1).requires_grad_(True)
    gamma = (torch.randn(D, device="cuda", generator=g) * 0.3 + 1).requires_grad_(True)
    u = F.normalize(x, dim=-1) * D ** 0.5
    h = u * film[:, None, :D] + film[:, None, D:] if mode == "film" else u * gamma
    h.backward(dh.float())
    dxr = torch.randn(B, N, D, device="cuda", generator=g)
    expect = dxr + x.grad
    dxr_bf = torch.empty(B, N, D, device="cuda", dtype=bf)
    dfilm = torch.zeros(B, 2 * D, device="cuda")
    dgamma = torch.zeros(D, device="cuda")
    if mode == "film":
        ops.rmsnorm_film_bwd(x.detach(), dh, dxr, dxr_bf, rows_per_batch=N, film=film.detach(), dfilm=dfilm)
        _close(dfilm, film.grad, 2e-2, 1e-3, "dfilm")
    else:
        ops.rmsnorm_film_bwd(x.detach(), dh, dxr, dxr_bf, rows_per_batch=N, gamma=gamma.detach(), dgamma=dgamma)
        _close(dgamma, gamma.grad, 5e-2, 1e-3, "dgamma")
    _close(dxr, expect, 1e-4, 1e-4, "dx")
    _close(dxr_bf, expect, 2e-2, 1e-2, "dx bf16")


def test_geglu_bwd_matches_autograd():
    from naturalspeech2_pytorch_b200 import ops
    R, Dp = 300, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    pre = torch.randn(R, 2 * Dp, device="cuda", generator=g).to(bf)
    dg = torch.randn(R, Dp, device="cuda", generator=g).to(bf)
    t = pre.float().view(R, Dp // 128, 2, 128).requires_grad_(True)      # tiles of [128 value | 128 gate]
    out = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(R, Dp)
    out.backward(dg.float())
    got = ops.geglu_bwd(pre.clone(), dg)
    _close(got, t.grad.reshape(R, 2 * Dp), 2e-2, 2e-2, "geglu bwd")


def test_wavenet_gate_bwd_matches_autograd():
    from naturalspeech2_pytorch_b200 import ops
    B, N, D, G = 2, 70, 128, 3
    g = torch.Generator(device="cuda").manual_seed(4)
    c = torch.randn(B, N, G * D, device="cuda", generator=g).to(bf)
    dy = torch.randn(B, N, G * D, device="cuda", generator=g).to(bf)
    film = (torch.randn(B, G * 2 * D, device="cuda", generator=g) * 0.5 + 0.5).requires_grad_(True)
    cf = c.float().requires_grad_(True)
    fv = film.view(B, G, 2, D)
    z = cf.view(B, N, G, D) * fv[:, None, :, 0] + fv[:, None, :, 1]
    y = (z.tanh() * z.sigmoid()).reshape(B, N, G * D)
    y.backward(dy.float())
    dc = torch.empty(B, N, G * D, device="cuda", dtype=bf)
    dfilm = torch.zeros(B, G * 2 * D, device="cuda")
    ops.wavenet_gate_bwd(c, dy, dc, film.detach(), dfilm, dim=D, groups=G, film_group_stride=2 * D)
    _close(dc, cf.grad, 2e-2, 2e-2, "dc")
    _close(dfilm, film.grad, 5e-2, 5e-3, "dfilm")


def test_small_reductions():
    from naturalspeech2_pytorch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    t = torch.randn(3, 500, 384, device="cuda", generator=g).to(bf)
    out = torch.ones(384, device="cuda")
    ops.colsum(t, out)
    _close(out, 1 + t.float().sum((0, 1)), 2e-2, 1e-3, "colsum")
    tg = torch.randn(2, 90, 4 * 128, device="cuda", generator=g).to(bf)
    gs = torch.empty(2, 90, 128, device="cuda", dtype=bf)
    ops.group_sum(tg, gs, dim=128, groups=4)
    _close(gs, tg.float().view(2, 90, 4, 128).sum(2), 3e-2, 1e-2, "group_sum")
    pred, target = torch.randn(4, 64, 128, device="cuda", generator=g), torch.randn(4, 64, 128, device="cuda", generator=g)
    coef = torch.rand(4, device="cuda", generator=g)
    ob = torch.empty(4, 64, 128, device="cuda", dtype=bf)
    ops.mse_bwd(pred, target, coef, ob)
    _close(ob, coef[:, None, None] * (pred - target), 2e-2, 1e-2, "mse_bwd")
    dfilm, tt = torch.randn(40, 700, device="cuda", generator=g), torch.randn(40, 300, device="cuda", generator=g)
    dw = torch.ones(700, 300, device="cuda")
    ops.film_wgrad(dfilm, tt, dw)
    _close(dw, 1 + dfilm.t() @ tt, 1e-3, 1e-4, "film_wgrad")
    dw2 = torch.full_like(dw, float("nan"))                    # overwrite mode never reads the buffer
    ops.film_wgrad(dfilm, tt, dw2, accumulate=False)
    _close(dw2, dfilm.t() @ tt, 1e-3, 1e-4, "film_wgrad overwrite")
    wide = torch.randn(40, 1000, device="cuda", generator=g)   # a column window of a wider gradient table (per-layer use)
    dw3 = torch.empty(600, 300, device="cuda")
    ops.film_wgrad(wide[:, 200:800], tt, dw3, accumulate=False)
    _close(dw3, wide[:, 200:800].t() @ tt, 1e-3, 1e-4, "film_wgrad window")


@pytest.mark.parametrize("B,H,Nq,Nk", [(1, 1, 128, 128), (2, 4, 1024, 1024), (2, 2, 200, 300), (2, 8, 256, 32), (1, 2, 32, 135)])
def test_attention_bwd_matches_autograd(B, H, Nq, Nk):
    from naturalspeech2_pytorch_b200 import ops
    inner = H * 64
    g = torch.Generator(device="cuda").manual_seed(6)
    q = torch.randn(B, Nq, inner, device="cuda", generator=g).to(bf)
    k = torch.randn(B, Nk, inner, device="cuda", generator=g).to(bf)
    v = torch.randn(B, Nk, inner, device="cuda", generator=g).to(bf)
    d_o = torch.randn(B, Nq, inner, device="cuda", generator=g).to(bf)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    sp = lambda t: t.reshape(B, -1, H, 64).transpose(1, 2)
    sim = (sp(qf) @ sp(kf).transpose(-1, -2)) * 64 ** -0.5
    ref = (sim.softmax(dim=-1) @ sp(vf)).transpose(1, 2).reshape(B, Nq, inner)
    ref.backward(d_o.float())
    o = torch.empty(B, Nq, inner, device="cuda", dtype=bf)
    lse = torch.empty(B, H, Nq, device="cuda")
    ops.attention(q, k, v, o, heads=H, lse=lse)
    ref_lse = torch.logsumexp(sim, dim=-1) * 1.4426950408889634
    _close(lse, ref_lse, 2e-2, 1e-3, "lse")
    dq = torch.zeros(B, Nq, inner, device="cuda")
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    ops.attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, heads=H)
    _close(dv, vf.grad, 3e-2, 3e-2, "dv")
    _close(dk, kf.grad, 3e-2, 3e-2, "dk")
    _close(dq, qf.grad, 3e-2, 3e-2, "dq")
