"""CPU: host-side logic — weight packing layouts, schedules, data-parallel sharding (gloo, world_size 2)."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import build_model, load_model_golden


def test_conv3_and_geglu_packing_layouts():
    from naturalspeech2_pytorch_b200 import Model
    _, kwargs, seed = load_model_golden("uncond_small")
    m = build_model(kwargs, seed)
    P = m._pack()
    D, Di = m.dim, m.ff_inner
    Dp = (Di + 127) // 128 * 128
    ff = m.transformer.layers[0][5]
    # GEGLU: every 256-row tile = 128 value rows then the 128 matching gate rows; zero rows beyond Di
    w1 = P["l0_ff_w1"].float()
    lin1 = ff[0].weight.detach()
    for tile in range(Dp // 128):
        for j in (0, 5, 127):
            ch = tile * 128 + j
            val_row, gate_row = w1[tile * 256 + j], w1[tile * 256 + 128 + j]
            if ch < Di:
                assert torch.allclose(val_row, lin1[ch].bfloat16().float())
                assert torch.allclose(gate_row, lin1[Di + ch].bfloat16().float())
            else:
                assert not val_row.any() and not gate_row.any()
    # causal conv: tap t occupies columns [t*Dp, t*Dp + Di)
    wc = P["l0_ff_wc"].float()
    conv = ff[2][1].weight.detach()
    assert wc.shape == (Dp, 3 * Dp)
    for t in range(3):
        assert torch.allclose(wc[:Di, t * Dp:t * Dp + Di], conv[:, :, t].bfloat16().float())
        assert not wc[:, t * Dp + Di:(t + 1) * Dp].any()
    assert not wc[Di:].any()
    # wavenet stack: rows of block i = [tap0 | tap1 | tap2 | res_conv]
    blk = m.wavenet.stacks[0].blocks[1]
    w = P["wn0_w"].float()[D:2 * D]
    for t in range(3):
        assert torch.allclose(w[:, t * D:(t + 1) * D], blk.conv.weight[:, :, t].detach().bfloat16().float())
    assert torch.allclose(w[:, 3 * D:], blk.res_conv.weight[:, :, 0].detach().bfloat16().float())
    # FiLM stack: wavenet blocks first, then the transformer norms
    G = m.wavenet_layers
    assert P["film_w"].shape[0] == (m.wavenet_stacks * G + 2 * m.depth) * 2 * D
    assert m._film_tr_off == m.wavenet_stacks * G * 2 * D


def test_schedules_match_oracle():
    from naturalspeech2_pytorch_b200.diffusion import sigmoid_schedule, gamma_to_alpha_sigma
    from oracle import diffusion_oracle
    t = torch.linspace(0, 1, 17)
    g = sigmoid_schedule(t.clone())
    np.testing.assert_allclose(g.numpy(), diffusion_oracle.sigmoid_schedule(t.numpy()), rtol=1e-6, atol=1e-7)
    a, s = gamma_to_alpha_sigma(g)
    ao, so = diffusion_oracle.gamma_to_alpha_sigma(g.numpy())
    np.testing.assert_allclose(a.numpy(), ao, rtol=1e-6)
    np.testing.assert_allclose(s.numpy(), so, rtol=1e-6, atol=1e-7)


def test_sampling_time_pairs():
    from naturalspeech2_pytorch_b200 import Model, NaturalSpeech2
    from oracle import diffusion_oracle
    m = Model(dim=128, depth=1, heads=1, wavenet_layers=1, wavenet_stacks=1)
    ns = NaturalSpeech2(m, target_sample_hz=24000, timesteps=5)
    pairs = ns.get_sampling_timesteps(3, device="cpu")
    ref = diffusion_oracle.sampling_time_pairs(5)
    assert len(pairs) == 5
    for (t, tn), (rt, rtn) in zip(pairs, ref):
        assert t.shape == (3,) and float(t[0]) == pytest.approx(rt) and float(tn[2]) == pytest.approx(rtn)


def test_constructor_guards():
    from naturalspeech2_pytorch_b200 import Model, NaturalSpeech2, EncodecRVQ
    with pytest.raises(NotImplementedError):
        Model(dim=100, depth=1)
    with pytest.raises(NotImplementedError):
        Model(dim=128, depth=1, dim_head=32)
    m = Model(dim=128, depth=1, heads=1, wavenet_layers=1, wavenet_stacks=1)
    with pytest.raises(AssertionError):
        NaturalSpeech2(m)  # neither codec nor target_sample_hz (ns2.py:1207)
    codec = EncodecRVQ(torch.randn(8, 1024, 128))
    assert codec.seq_len_multiple_of == 320 and codec.codebook_dim == 128
    NaturalSpeech2(m, codec)
    m512 = Model(dim=512, depth=1, heads=1, wavenet_layers=1, wavenet_stacks=1)
    with pytest.raises(AssertionError):
        NaturalSpeech2(m512, codec)  # model.dim must equal codec.codebook_dim (ns2.py:1244)


def test_shard_bounds_cover_batch():
    from naturalspeech2_pytorch_b200.parallel import shard_bounds
    for gb in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_bounds(gb, r, world)
                covered += list(range(lo, hi))
            assert covered == list(range(gb))


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from naturalspeech2_pytorch_b200 import parallel
    r, w, _ = parallel.init_from_env(backend="gloo")
    per_sample = torch.arange(7, dtype=torch.float32) ** 2  # "loss" of every sample of the global batch
    mine = parallel.shard_batch(per_sample, r, w)
    out = parallel.global_mean_loss(mine.mean(), mine.numel())
    q.put((rank, float(out)))
    dist.destroy_process_group()


def test_scalar_loss_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = float((torch.arange(7, dtype=torch.float32) ** 2).mean())
    assert res[0] == pytest.approx(expect) and res[1] == pytest.approx(expect)


def _reducer_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from naturalspeech2_pytorch_b200 import parallel
    parallel.init_from_env(backend="gloo")
    red = parallel.GradReducer(coalesce_below=1024)
    big = torch.full((4, 512), float(rank + 1))            # 8 KB: its own (asynchronous) collective
    small = torch.full((7,), float(10 * (rank + 1)))       # coalesced into the flat tail message
    grads = {"a": big[:2], "b": big[2:], "c": small, "d": small[:3]}   # views of shared packed buffers
    red.reduce_all(grads)
    red.finish()
    q.put((rank, float(big.mean()), float(small.mean()), red.bytes_reduced))
    dist.destroy_process_group()


def test_grad_reducer_averages_packed_buffers_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, big, small, nbytes in res:
        assert big == pytest.approx(1.5) and small == pytest.approx(15.0)
        assert nbytes == 4 * 512 * 4 + 7 * 4       # every base buffer exactly once


def test_rvq_empty_input_returns_empty_tensors():
    from naturalspeech2_pytorch_b200 import EncodecRVQ
    codec = EncodecRVQ(torch.randn(8, 1024, 128))
    codes, emb = codec.quantize(torch.empty(0, 128))
    assert codes.shape == (0, 8) and codes.dtype == torch.int64 and emb.shape == (0, 128)
    codes, emb = codec.quantize(torch.empty(2, 0, 128))
    assert codes.shape == (2, 0, 8) and emb.shape == (2, 0, 128)


def test_conditional_packing_layouts():
    """Cross-attention K/V weights of all layers are stacked into one matrix; FiLM rows follow the layer order."""
    _, kwargs, seed = load_model_golden("cond_small")
    m = build_model(kwargs, seed)
    P = m._pack()
    D, inner, depth = m.dim, m.inner, m.depth
    assert P["x_kv_all"].shape == (depth * 2 * inner, D)
    for l in range(depth):
        ref = m.transformer.layers[l][3].to_kv.weight.detach().bfloat16().float()
        assert torch.allclose(P["x_kv_all"].float()[l * 2 * inner:(l + 1) * 2 * inner], ref)
        qkv = P[f"l{l}_qkv"].float()
        assert torch.allclose(qkv[:inner], m.transformer.layers[l][1].to_q.weight.detach().bfloat16().float())
        assert torch.allclose(qkv[inner:], m.transformer.layers[l][1].to_kv.weight.detach().bfloat16().float())
    # three adaptive norms per layer when conditional: attn, cross-attn, ff — in that order after the wavenet rows
    assert m._norms_per_layer == 3
    G = m.wavenet_layers
    rows = (m.wavenet_stacks * G + 3 * depth) * 2 * D
    assert P["film_w"].shape == (rows, m.dim_cond) and P["film_b"].shape == (rows,)
    off = m._film_tr_off + (1 * 3 + 1) * 2 * D  # layer 1, cross-attn norm
    ref = m.transformer.layers[1][2].to_gamma_beta.weight.detach().bfloat16().float()
    assert torch.allclose(P["film_w"].float()[off:off + 2 * D], ref)
    # perceiver + aligned-condition projection
    assert P["pr_proj_w"].shape == (D, m.dim_prompt) and P["cond_w"].shape == (D, m.dim_prompt)
    # skip convs concatenated along K, biases summed
    last = m.wavenet.stacks[-1]
    assert P["wn_skip_w"].shape == (D, G * D)
    assert torch.allclose(P["wn_skip_b"], torch.stack([b.skip_conv.bias for b in last.blocks]).sum(0).detach())
