"""GPU parity of `Model.forward` (CUDA kernels through the C ABI) against the golden vectors generated from the
reference and against the numpy oracle.

Tolerance protocol (the north-star rtol=1e-3/atol=1e-5 is not reachable by ANY bf16-operand
implementation — the reference's own autocast-bf16 output misses it on >90% of elements):
  * ground truth = reference fp64 output (golden) or the fp64 oracle;
  * bar = our error must not exceed the error of the reference's own bf16-autocast run against the same ground
    truth (max-abs AND rms), both recorded in the golden file; we land ~1.5x below it because the residual
    stream, norm statistics, softmax and accumulators stay in fp32; an absolute bound is asserted as well.
"""
import numpy as np
import pytest
import torch

from helpers import (BIG_MODEL_CASES, MODEL_CASES, build_model, err_stats, golden_inputs, golden_rows,
                     load_model_golden, numpy_params, oracle_config)

pytestmark = pytest.mark.gpu


def _run(model, z, kwargs, **fw):
    dev = "cuda"
    inp = golden_inputs(z, kwargs)
    x, times = inp["x"].to(dev), inp["times"].to(dev)
    extra = {}
    if kwargs.get("condition_on_prompt"):
        extra = dict(prompt=inp["prompt"].to(dev), cond=inp["cond"].to(dev))
    return model, x, times, extra


@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_forward_vs_reference_golden(name):
    z, kwargs, seed = load_model_golden(name)
    model = build_model(kwargs, seed, device="cuda")
    model, x, times, extra = _run(model, z, kwargs)
    out = model(x, times, **extra).float().cpu().numpy()
    assert np.isfinite(out).all()
    emax, erms = err_stats(out, z["out_fp64"])
    ref_max, ref_rms = err_stats(z["out_bf16_autocast"], z["out_fp64"])
    print(f"{name}: ours max={emax:.3e} rms={erms:.3e} | reference bf16-autocast max={ref_max:.3e} rms={ref_rms:.3e}")
    assert emax <= ref_max and erms <= ref_rms, (emax, erms, ref_max, ref_rms)
    # absolute bound, independent of the reference's own error (output std is ~1): bf16-operand noise
    assert emax < 5e-2 and erms < 1e-2, (emax, erms)
    # determinism: same inputs -> bit-identical output
    out2 = model(x, times, **extra).float().cpu().numpy()
    np.testing.assert_array_equal(out, out2)


@pytest.mark.parametrize("name", BIG_MODEL_CASES)
def test_model_forward_at_benchmarked_dims(name):
    """Whole-model parity at the kernel instantiations the bench runs (dim 512, heads 8, seq 1024: 256-wide CTA-pair
    tiles, n=1408 partial last tile, two-accumulator wavenet tile, 148-SM persistent schedules), chained through the
    4 wavenet stacks + 2 transformer layers, against the reference's fp64 output.  Same protocol as above; also
    reports (not asserts — no bf16-operand implementation can meet it, see the module docstring) the fraction of
    elements inside the north-star rtol=1e-3/atol=1e-5 band against the reference's fp32 output."""
    z, kwargs, seed = load_model_golden(name)
    model = build_model(kwargs, seed, device="cuda")
    model, x, times, extra = _run(model, z, kwargs)
    full = model(x, times, **extra).float().cpu().numpy()
    assert np.isfinite(full).all()
    out = golden_rows(z, full)
    emax, erms = err_stats(out, z["out_fp64"])
    ref_max, ref_rms, ref_frac = (float(v) for v in z["stats_out_bf16_autocast"])   # whole-tensor statistics
    frac = float(np.isclose(out, z["out_fp32"], rtol=1e-3, atol=1e-5).mean())
    print(f"{name}: ours max={emax:.3e} rms={erms:.3e} strict-band frac vs ref fp32={frac:.4f} | reference "
          f"bf16-autocast max={ref_max:.3e} rms={ref_rms:.3e} strict-band frac={ref_frac:.4f}")
    assert emax <= ref_max and erms <= ref_rms, (emax, erms, ref_max, ref_rms)
    assert emax < 5e-2 and erms < 1e-2, (emax, erms)
    assert frac >= ref_frac, (frac, ref_frac)
    if kwargs.get("condition_on_prompt"):
        null = golden_rows(z, model(x, times, cond_drop_prob=1., **extra).float().cpu().numpy())
        nmax, nrms = err_stats(null, z["out_fp64_null"])
        assert nmax < 5e-2 and nrms < 1e-2, (nmax, nrms)
        cfg = golden_rows(z, model.forward_with_cond_scale(x, times, cond_scale=3., **extra).float().cpu().numpy())
        cmax, crms = err_stats(cfg, z["out_fp64_cfg3"])
        # guidance: out = null + 3 (cond - null) amplifies the independent errors of the two passes by <= 3 + 2
        assert cmax < 5 * 5e-2 and crms < 5 * erms + 1e-3, (cmax, crms)
    # CUDA-graph replay of the same step is bit-identical
    model.use_cuda_graphs = True
    kw = dict(_conditioning=model.precompute_conditioning(extra["prompt"], extra["cond"], x.shape[1])) if extra else {}
    g = model(x, times, **kw).float().cpu().numpy()
    np.testing.assert_array_equal(full, g)


@pytest.mark.parametrize("name", ["cond_small", "cond_samedim"])
def test_model_cfg_paths(name):
    z, kwargs, seed = load_model_golden(name)
    model = build_model(kwargs, seed, device="cuda")
    model, x, times, extra = _run(model, z, kwargs)
    null = model(x, times, cond_drop_prob=1., **extra).float().cpu().numpy()
    emax, _ = err_stats(null, z["out_fp64_null"])
    assert emax < 5e-2, emax
    cfg = model.forward_with_cond_scale(x, times, cond_scale=3., **extra).float().cpu().numpy()
    emax, erms = err_stats(cfg, z["out_fp64_cfg3"])
    # out = null + 3 (cond - null): the independent errors of the two passes are amplified by at most 3 + 2
    assert emax < 1.5e-1 and erms < 3e-2, (emax, erms)


def test_model_vs_oracle_other_shape():
    """A shape no golden covers (ragged N, B=3), checked against the fp64 numpy oracle on the same weights."""
    from oracle import denoiser_oracle
    kwargs = dict(dim=128, depth=1, heads=3, wavenet_layers=4, wavenet_stacks=3)
    model = build_model(kwargs, 99, device="cuda")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 77, 128, generator=g)
    times = torch.rand(3, generator=g)
    out = model(x.cuda(), times.cuda()).float().cpu().numpy()
    ref = denoiser_oracle.model_forward(numpy_params(model), oracle_config(kwargs), x.numpy(), times.numpy())
    emax, erms = err_stats(out, ref)
    print(f"oracle parity: max={emax:.3e} rms={erms:.3e}")
    assert emax < 5e-2 and erms < 1e-2, (emax, erms)


def test_state_dict_roundtrip_and_repack():
    """Loading new weights must invalidate the packed bf16 copies."""
    _, kwargs, seed = load_model_golden("uncond_small")
    m1 = build_model(kwargs, seed, device="cuda")
    m2 = build_model(kwargs, seed + 1, device="cuda")
    x = torch.randn(1, 128, 128, device="cuda")
    t = torch.rand(1, device="cuda")
    a = m1(x, t).clone()
    b = m2(x, t).clone()
    assert not torch.equal(a, b)
    m2.load_state_dict(m1.state_dict())
    c = m2(x, t).clone()
    assert torch.equal(a, c)


def test_cpu_tensor_is_rejected():
    _, kwargs, seed = load_model_golden("uncond_small")
    m = build_model(kwargs, seed, device="cuda")
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 128, 128), torch.rand(1))


@pytest.mark.parametrize("name", ["uncond_small", "cond_small"])
def test_cuda_graph_replay_matches_eager(name):
    """use_cuda_graphs: the captured step replays bit-identically to the eager launches, also on new inputs."""
    z, kwargs, seed = load_model_golden(name)
    model = build_model(kwargs, seed, device="cuda")
    model, x, times, extra = _run(model, z, kwargs)
    cached = model.precompute_conditioning(extra["prompt"], extra["cond"], x.shape[1]) if extra else None
    kw = dict(_conditioning=cached) if extra else {}
    eager = model(x, times, **kw).clone()
    x2, t2 = torch.randn_like(x), torch.rand_like(times)
    eager2 = model(x2, t2, **kw).clone()
    model.use_cuda_graphs = True
    g1 = model(x, times, **kw).clone()
    g2 = model(x2, t2, **kw).clone()
    g1b = model(x, times, **kw).clone()
    assert torch.equal(eager, g1) and torch.equal(eager2, g2) and torch.equal(g1, g1b)
    assert len(model._graphs) == 1


def test_large_batch_is_chunked_in_the_small_layers():
    """B > 64 exceeds the per-launch limit of the conditioning-vector kernels; the host splits it (same results)."""
    _, kwargs, seed = load_model_golden("uncond_small")
    model = build_model(kwargs, seed, device="cuda")
    x = torch.randn(96, 128, 128, device="cuda")
    t = torch.rand(96, device="cuda")
    full = model(x, t).clone()
    part = torch.cat([model(x[:48], t[:48]).clone(), model(x[48:], t[48:]).clone()])
    assert torch.equal(full, part)


def test_outputs_are_fresh_tensors_and_out_argument():
    """Reference semantics: every forward returns its own tensor (two predictions can be held at once); `out=` writes
    into a caller-owned buffer; both also under CUDA-graph replay."""
    _, kwargs, seed = load_model_golden("uncond_small")
    model = build_model(kwargs, seed, device="cuda")
    x1, x2 = torch.randn(2, 128, 128, device="cuda"), torch.randn(2, 128, 128, device="cuda")
    t = torch.rand(2, device="cuda")
    for graphs in (False, True):
        model.use_cuda_graphs = graphs
        a = model(x1, t)
        a_copy = a.clone()
        b = model(x2, t)
        assert a.data_ptr() != b.data_ptr() and torch.equal(a, a_copy) and not torch.equal(a, b)
        buf = torch.empty_like(x1)
        c = model(x1, t, out=buf)
        assert c.data_ptr() == buf.data_ptr() and torch.equal(c, a)


def test_data_updates_need_invalidate_packed():
    """`.data` updates (EMA-style lerp_) do not bump the version counter: `invalidate_packed()` makes them visible,
    also to captured graphs."""
    _, kwargs, seed = load_model_golden("uncond_small")
    m1 = build_model(kwargs, seed, device="cuda")
    m2 = build_model(kwargs, seed + 1, device="cuda")
    x, t = torch.randn(1, 128, 128, device="cuda"), torch.rand(1, device="cuda")
    m1.use_cuda_graphs = True
    before = m1(x, t)
    with torch.no_grad():
        for p, q in zip(m1.parameters(), m2.parameters()):
            p.data.copy_(q.data)
    m1.invalidate_packed()
    after = m1(x, t)
    assert torch.equal(after, m2(x, t)) and not torch.equal(after, before)


def test_partial_condition_dropout_matches_composition():
    """0 < cond_drop_prob < 1: the two Bernoulli masks are drawn in the reference's order (ns2.py:950, 980); the
    output equals, sample by sample, the fully-conditioned or fully-null forward selected by those masks."""
    z, kwargs, seed = load_model_golden("cond_small")
    model = build_model(kwargs, seed, device="cuda")
    inp = golden_inputs(z, kwargs)
    x, times = inp["x"].cuda().repeat(4, 1, 1), inp["times"].cuda().repeat(4)
    prompt, cond = inp["prompt"].cuda().repeat(4, 1, 1), inp["cond"].cuda().repeat(4, 1, 1)
    B = x.shape[0]
    torch.manual_seed(123)
    out = model(x, times, prompt=prompt, cond=cond, cond_drop_prob=0.5)
    torch.manual_seed(123)
    m_prompt = torch.zeros((B,), device="cuda").float().uniform_(0, 1) < 0.5
    m_cond = torch.zeros((B,), device="cuda").float().uniform_(0, 1) < 0.5
    keep = model(x, times, prompt=prompt, cond=cond, cond_drop_prob=0.)
    null = model(x, times, prompt=prompt, cond=cond, cond_drop_prob=1.)
    both = m_prompt == m_cond          # samples where both masks agree can be compared with the pure forwards
    assert bool(both.any()) and bool((~m_prompt & both).any()) and bool((m_prompt & both).any())
    ref = torch.where(m_prompt[:, None, None], null, keep)
    assert torch.equal(out[both], ref[both])
