"""GPU parity for the captured decode step (SURVEY.md §8f N2 / BASELINE.json configs[3]).

The graphed greedy decoder must emit exactly the ids of the un-graphed loop (`CT_DECODE_GRAPH=0`, the path the golden
vectors of tests/test_gpu_models.py pin against the real reference): both run the same kernels on the same values, only
the cache length / write column / alive flags move from host integers into device memory. The decode attention kernel
and the greedy-step kernel are also checked on their own against torch.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_decode(q, k, v, scale, kbias2, n):
    """fp32 softmax(q.k * scale + bias) . v over the first n keys, P rounded like the kernels do not matter here."""
    qf, kf, vf = q.float(), k[:, :, :n].float(), v[:, :, :n].float()
    s = torch.einsum("bhqd,bhkd->bhqk", qf, kf) * scale
    if kbias2 is not None:
        s = s + (kbias2[:, :, None, :n] / 1.4426950408889634)
    return torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, -1), vf)


@pytest.mark.parametrize("D", [32, 64, 128])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n_keys", [1, 5, 130, 777])
def test_decode_attention_kernel(D, dtype, n_keys):
    from cleantransformer_b200 import ops
    torch.manual_seed(D + n_keys)
    B, H, CAP = 3, 5, 1024
    q = torch.randn(B, H, 1, D, device=DEV).to(dtype)
    k = torch.randn(B, H, CAP, D, device=DEV).to(dtype)
    v = torch.randn(B, H, CAP, D, device=DEV).to(dtype)
    kb = torch.randn(B, H, CAP, device=DEV) * 2
    kb[1, :, :n_keys // 2] = -float("inf")  # left padding of one row
    scale = D ** -0.5
    want = _ref_decode(q, k, v, scale, kb, n_keys)
    # host-side key count (the un-graphed loop: a [b,h,n,d] view of the cache)
    o1, _ = ops.attn_fwd(q, k[:, :, :n_keys], v[:, :, :n_keys], scale, True, -1e4, kb, None, need_lse=False)
    # device-side key count over the whole capacity (the captured step)
    n_dev = torch.tensor([n_keys, 0, 0, 0, 0], dtype=torch.int32, device=DEV)
    o2, _ = ops.attn_fwd(q, k, v, scale, True, -1e4, kb, None, need_lse=False, seq_len_dev=n_dev)
    assert torch.equal(o1, o2), "device-side and host-side key counts must give the same bits"
    got = o1.view(B, 1, H, D).permute(0, 2, 1, 3).float()
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3  # output + P rounding to the activation dtype
    assert (got - want).abs().max() <= tol * max(1.0, float(want.abs().max())), float((got - want).abs().max())


def test_decode_attention_without_bias_and_unsupported_head_dim():
    from cleantransformer_b200 import ops
    torch.manual_seed(0)
    q = torch.randn(2, 4, 1, 64, device=DEV).bfloat16()
    k = torch.randn(2, 4, 40, 64, device=DEV).bfloat16()
    v = torch.randn(2, 4, 40, 64, device=DEV).bfloat16()
    o, _ = ops.attn_fwd(q, k, v, 0.125, False, 0.0, None, None, need_lse=False)
    want = _ref_decode(q, k, v, 0.125, None, 40)
    assert (o.view(2, 1, 4, 64).permute(0, 2, 1, 3).float() - want).abs().max() < 2e-2
    q48 = torch.randn(2, 4, 1, 48, device=DEV).bfloat16()
    k48 = torch.randn(2, 4, 40, 48, device=DEV).bfloat16()
    n_dev = torch.tensor([40], dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError):
        ops.attn_fwd(q48, k48, k48, 0.1, False, 0.0, None, None, need_lse=False, seq_len_dev=n_dev)


def test_kv_append_dev_writes_at_the_device_side_position():
    from cleantransformer_b200 import ops
    torch.manual_seed(1)
    base = torch.zeros(2, 3, 16, 64, device=DEV, dtype=torch.bfloat16)
    new = torch.randn(2, 3, 1, 64, device=DEV).bfloat16()
    n = torch.tensor([7], dtype=torch.int32, device=DEV)
    ops.kv_append_dev(base, new, n)
    assert torch.equal(base[:, :, 6], new[:, :, 0]) and float(base[:, :, :6].abs().sum()) == 0
    assert float(base[:, :, 7:].abs().sum()) == 0
    n.fill_(17)  # one past the capacity: the kernel must not write outside the buffer
    ops.kv_append_dev(base, new, n)
    torch.cuda.synchronize()
    assert float(base[:, :, 7:].abs().sum()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_greedy_step_kernel_follows_generation_util(dtype):
    """generation_util.py:86-101 with do_sample=False: argmax (first maximum), pad for finished rows, end ids."""
    from cleantransformer_b200 import ops
    torch.manual_seed(2)
    B, V, T = 6, 50257, 12
    logits = torch.randn(B, V, device=DEV).to(dtype)
    logits[0, 77] = logits[0, 4000] = 9.0     # tie: the first index wins
    logits[1, 13] = 50.0                      # row 1 emits end id 13 and dies
    logits[3, V - 1] = 60.0                   # last column
    alive = torch.ones(B, dtype=torch.long, device=DEV)
    alive[2] = 0                              # already finished: emits pad
    end_ids = torch.tensor([13, 99], device=DEV)
    ids_out = torch.full((B, T), -1, dtype=torch.long, device=DEV)
    cur = torch.zeros(B, dtype=torch.long, device=DEV)
    pos = torch.arange(B, device=DEV)
    state = torch.tensor([4, 5, 5, -1, 0], dtype=torch.int32, device=DEV)
    want = logits.argmax(-1) * alive + 7 * (1 - alive)
    ops.greedy_step(logits, alive, end_ids, 7, ids_out, cur, pos, state)
    assert torch.equal(ids_out[:, 5], want) and torch.equal(cur, want)
    assert int(want[0]) == 77 and int(want[2]) == 7 and int(want[3]) == V - 1
    assert alive.tolist() == [1, 0, 0, 1, 1, 1]
    assert torch.equal(pos, torch.arange(B, device=DEV) + 1)
    assert state.tolist() == [5, 6, 4, -1, 0]
    assert int((ids_out[:, :5] != -1).sum()) == 0 and int((ids_out[:, 6:] != -1).sum()) == 0
    # every remaining row hits an end id: done_at = the column after this step
    logits[:] = 0
    logits[:, 99] = 1
    ops.greedy_step(logits, alive, end_ids, 7, ids_out, cur, None, state)
    assert alive.tolist() == [0] * B and state.tolist() == [6, 7, 0, 7, 0]
    assert ids_out[:, 6].tolist() == [99, 7, 7, 99, 99, 99]


def _gen(model, ids, mask, graph, **cfg):
    old = os.environ.get("CT_DECODE_GRAPH")
    os.environ["CT_DECODE_GRAPH"] = "1" if graph else "0"
    try:
        base = {"beam_size": 1, "do_sample": False, "max_gen_len": 20, "end_ids": None, "pad_id": 0}
        base.update(cfg)
        return model.generate(ids, attention_mask=mask, generation_configs=base)
    finally:
        if old is None:
            os.environ.pop("CT_DECODE_GRAPH", None)
        else:
            os.environ["CT_DECODE_GRAPH"] = old


def _left_padded(B, P, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, vocab, (B, P), generator=g)
    mask = torch.ones(B, P, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(P // 2, P + 1, (1,), generator=g))
        mask[b, :P - n] = 0
        ids[b, :P - n] = 0
    return ids.to(DEV), mask.to(DEV)


def _init(model, std=0.08):
    g = torch.Generator(device="cpu").manual_seed(5)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif "bias" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)


@pytest.mark.parametrize("family", ["gpt2", "bloom"])
def test_graphed_greedy_decode_emits_the_ids_of_the_loop(family):
    if family == "gpt2":
        from cleantransformer_b200.models import modeling_gpt as mg
        cfg = mg.GPTConfig(vocab_size=1000, n_embd=256, n_positions=256, n_layer=3, n_head=4, n_ctx=256, afn="gelu_new")
        model = mg.GPTLMHeadModel(cfg, version="gpt2").to(DEV).eval()
        _init(model)
        model._tie_weights()
    else:
        from cleantransformer_b200.models import modeling_bloom as mb
        cfg = mb.BloomConfig(vocab_size=1000, hidden_size=256, n_layer=3, num_attention_heads=4, hidden_dropout=0.0,
                             attention_dropout=0.0)
        model = mb.BloomForCausalLM(cfg).to(DEV).eval()
        _init(model)
        model._tie_weight()
    ids, mask = _left_padded(5, 24, 1000, 11)
    want = _gen(model, ids, mask, graph=False)
    got = _gen(model, ids, mask, graph=True)
    assert model._ct_decode_graph_launches > 0, "the captured path did not run"
    assert want.shape == got.shape == (5, 1, 24 + 22)
    assert torch.equal(want, got)
    # end ids: pick tokens the un-graphed run emits, so rows die at different steps and the loop stops early
    gen = want[:, 0, 24:]
    end_ids = [int(gen[0, 3]), int(gen[1, 9])]
    want_e = _gen(model, ids, mask, graph=False, end_ids=end_ids, pad_id=3)
    got_e = _gen(model, ids, mask, graph=True, end_ids=end_ids, pad_id=3)
    assert want_e.shape == got_e.shape and torch.equal(want_e, got_e)
    # every row finishes at once: the loop ends after the step that killed the last row
    all_end = sorted(set(gen[:, 2].tolist()))
    want_a = _gen(model, ids, mask, graph=False, end_ids=all_end)
    got_a = _gen(model, ids, mask, graph=True, end_ids=all_end)
    assert want_a.shape == got_a.shape and want_a.shape[-1] <= 24 + 3 and torch.equal(want_a, got_a)
    # max_gen_len = 1 (three emitted tokens: prefill + eager step + one replay)
    assert torch.equal(_gen(model, ids, mask, graph=False, max_gen_len=1), _gen(model, ids, mask, graph=True, max_gen_len=1))


def test_right_padded_prompt_falls_back_to_the_loop():
    """A prompt whose last mask column holds a 0 keeps masking its generated positions (generation_util.py:111 repeats
    the last column): the captured step assumes valid keys, so generate() must take the un-graphed loop."""
    from cleantransformer_b200.models import modeling_gpt as mg
    cfg = mg.GPTConfig(vocab_size=500, n_embd=128, n_positions=128, n_layer=2, n_head=2, n_ctx=128, afn="gelu_new")
    model = mg.GPTLMHeadModel(cfg, version="gpt2").to(DEV).eval()
    _init(model)
    model._tie_weights()
    ids, mask = _left_padded(3, 12, 500, 3)
    mask[1, -1] = 0
    model._ct_decode_graph_launches = -1
    out = _gen(model, ids, mask, graph=True, max_gen_len=4)
    assert out.shape == (3, 1, 12 + 6) and model._ct_decode_graph_launches == -1


@pytest.mark.parametrize("M,N,K", [(32, 1024, 1024), (32, 3072, 1024), (32, 1024, 4096), (32, 50257, 1024),
                                   (1, 4096, 1024), (7, 40, 96), (19, 1000, 32), (32, 250880, 64)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_skinny_gemm_vs_fp32_and_vs_the_tcgen05_kernel(M, N, K, dtype):
    """csrc/gemm.cu: gemm_skinny_kernel (the decode step's weight-streaming GEMM, impl 4; auto for M <= 32) against an
    fp32 product of the same rounded operands, with the full epilogue (bias, activation, residual, both output dtypes),
    and against the 128-row tcgen05 kernel (impl 1) it replaces."""
    from cleantransformer_b200 import ops
    torch.manual_seed(M * 7 + N + K)
    x = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(dtype)
    bias = torch.randn(N, device=DEV) * 0.1
    res = torch.randn(M, N, device=DEV)
    ref_lin = x.float() @ w.float().t()
    # plain f32 output
    y = ops.gemm(x, w, M, N, K, out_dtype=torch.float32, impl=4)
    assert (y - ref_lin).abs().max() <= 2e-5 * K ** 0.5 * max(1.0, float(ref_lin.abs().max()))
    y_auto = ops.gemm(x, w, M, N, K, out_dtype=torch.float32)
    if K % 32 == 0 and N * K >= (1 << 16) and N < 131072:  # (Bloom's 250 880-wide head stays on the tcgen05 kernel)
        assert torch.equal(y, y_auto), "auto dispatch must pick the skinny kernel for M <= 32"
    y_tc = ops.gemm(x, w, M, N, K, out_dtype=torch.float32, impl=1)
    assert (y - y_tc).abs().max() <= 1e-4 * max(1.0, float(ref_lin.abs().max()))
    # bias + tanh-GELU + residual, activation-dtype output
    want = torch.nn.functional.gelu(ref_lin + bias, approximate="tanh") + res
    got = ops.gemm(x, w, M, N, K, out_dtype=dtype, bias=bias, act=ops.ACT_GELU_TANH, residual=res, impl=4)
    tol = (8e-3 if dtype == torch.bfloat16 else 2e-3) * max(1.0, float(want.abs().max()))
    assert (got.float() - want).abs().max() <= tol
    # deterministic: the K split is summed in a fixed order
    assert torch.equal(got, ops.gemm(x, w, M, N, K, out_dtype=dtype, bias=bias, act=ops.ACT_GELU_TANH, residual=res, impl=4))


def test_skinny_gemm_refuses_layouts_it_does_not_serve():
    from cleantransformer_b200 import ops
    x = torch.randn(33, 64, device=DEV).bfloat16()
    w = torch.randn(128, 64, device=DEV).bfloat16()
    with pytest.raises(RuntimeError):
        ops.gemm(x, w, 33, 128, 64, out_dtype=torch.float32, impl=4)      # M > 32
    x = torch.randn(8, 48, device=DEV).bfloat16()
    w = torch.randn(128, 48, device=DEV).bfloat16()
    with pytest.raises(RuntimeError):
        ops.gemm(x, w, 8, 128, 48, out_dtype=torch.float32, impl=4)       # K % 32 != 0


def test_conv1d_decode_step_uses_the_k_major_shadow_and_matches_the_full_path():
    """modeling_gpt.py:32-46 Conv1D ([in,out] weight) with <= 32 rows under no_grad streams a cached [out,in] copy;
    the result must equal the [in,out] tcgen05 path up to summation order, and follow weight updates."""
    from cleantransformer_b200 import functional as F
    from cleantransformer_b200.models.modeling_gpt import Conv1D
    torch.manual_seed(3)
    lin = Conv1D(3072, 1024).to(DEV)
    x = torch.randn(32, 1, 1024, device=DEV)
    with torch.no_grad():
        y_dec = lin(x, out_dtype=torch.float32)
    assert getattr(lin.weight, "_ct_shadow_t", None) is not None and lin.weight._ct_shadow_t.shape == (3072, 1024)
    y_full = lin(x, out_dtype=torch.float32)  # grad mode on: the training path ([in,out] operand read in place)
    assert (y_dec - y_full.detach()).abs().max() <= 1e-4 * float(y_full.abs().max())
    with torch.no_grad():
        lin.weight.mul_(2.0)
        y2 = lin(x, out_dtype=torch.float32)
        lin.bias.zero_()
    assert (y2 - 2 * y_dec).abs().max() <= 2e-2 * float(y_dec.abs().max())  # refreshed shadow (bias was zero-initialised)


@pytest.mark.parametrize("D", [32, 64, 128])
def test_decode_attention_appends_the_new_token_itself(D):
    """ct_attn_args.k_new / v_new: the q_len = 1 kernel stores the new key / value at cache row (count - 1) and attends
    in one launch — same output bits and same cache contents as ct_kv_append_dev followed by the attention."""
    from cleantransformer_b200 import ops
    torch.manual_seed(D)
    B, H, CAP, n = 3, 5, 300, 131
    q = torch.randn(B, H, 1, D, device=DEV).bfloat16()
    qkv_new = torch.randn(B, 1, 2, H, D, device=DEV).bfloat16()          # strided views, like a packed projection
    k_new, v_new = qkv_new[:, :, 0].permute(0, 2, 1, 3), qkv_new[:, :, 1].permute(0, 2, 1, 3)
    k0 = torch.randn(B, H, CAP, D, device=DEV).bfloat16()
    v0 = torch.randn(B, H, CAP, D, device=DEV).bfloat16()
    kb = torch.randn(B, H, CAP, device=DEV)
    n_dev = torch.tensor([n, 0, 0, 0, 0], dtype=torch.int32, device=DEV)
    ka, va = k0.clone(), v0.clone()
    ops.kv_append_dev(ka, k_new, n_dev)
    ops.kv_append_dev(va, v_new, n_dev)
    want, _ = ops.attn_fwd(q, ka, va, D ** -0.5, True, -1e4, kb, None, need_lse=False, seq_len_dev=n_dev)
    kb_, vb_ = k0.clone(), v0.clone()
    got, _ = ops.attn_fwd(q, kb_, vb_, D ** -0.5, True, -1e4, kb, None, need_lse=False, seq_len_dev=n_dev,
                          kv_new=(k_new, v_new))
    assert torch.equal(got, want) and torch.equal(kb_, ka) and torch.equal(vb_, va)
    assert torch.equal(kb_[:, :, n - 1], k_new[:, :, 0]) and torch.equal(kb_[:, :, n:], k0[:, :, n:])


def test_captured_decode_with_sampling():
    """do_sample=True (the reference's default) through the captured step: torch's processors and multinomial are
    captured with the step (their Philox offset advances per replay), ct_greedy_step takes the drawn tokens. The CUDA
    generator is consumed differently by a replayed graph than by the host loop, so ids are not comparable draw by
    draw; checked instead: top_k = 1 (a deterministic draw) reproduces the greedy ids of both paths, a flat
    distribution gives different samples for different seeds and the same sample for the same seed, every drawn token
    lies in the top-k set of its step (teacher-forced recomputation), finished rows emit pad."""
    from cleantransformer_b200.models import modeling_gpt as mg
    cfg = mg.GPTConfig(vocab_size=1000, n_embd=256, n_positions=256, n_layer=3, n_head=4, n_ctx=256, afn="gelu_new")
    model = mg.GPTLMHeadModel(cfg, version="gpt2").to(DEV).eval()
    _init(model, std=0.05)
    model._tie_weights()
    ids, mask = _left_padded(4, 16, 1000, 5)
    greedy_loop = _gen(model, ids, mask, graph=False)
    greedy_graph = _gen(model, ids, mask, graph=True)
    assert torch.equal(greedy_loop, greedy_graph)
    samp = dict(do_sample=True, temperature=1.5, top_k=1, top_p=1.0)
    assert torch.equal(_gen(model, ids, mask, graph=True, **samp), greedy_graph)
    assert model._ct_decode_graph_launches > 0, "sampling did not take the captured step"
    assert torch.equal(_gen(model, ids, mask, graph=False, **samp), greedy_graph)
    flat = dict(do_sample=True, temperature=30.0, top_k=50, top_p=1.0)
    torch.manual_seed(11)
    a = _gen(model, ids, mask, graph=True, **flat)
    torch.manual_seed(11)
    b = _gen(model, ids, mask, graph=True, **flat)
    torch.manual_seed(12)
    c = _gen(model, ids, mask, graph=True, **flat)
    assert a.shape == (4, 1, 16 + 22) and torch.equal(a, b) and not torch.equal(a, c)
    assert int(a.min()) >= 0 and int(a.max()) < 1000
    # every sampled token is one of the 50 most likely of its step given the sampled prefix
    seq = a[:, 0]
    full_mask = torch.cat([mask, mask[:, -1:].expand(4, 22)], dim=1)
    with torch.no_grad():
        (logits, _), _ = model(seq[:, :-1], attention_mask=full_mask[:, :-1])
    top = logits[:, 15:, :].float().topk(50, dim=-1).indices          # predictions for positions 16 ..
    hit = (top == seq[:, 16:, None]).any(-1)
    assert float(hit.float().mean()) > 0.97, float(hit.float().mean())  # (bf16 near-ties at the 50th place aside)
    # end ids: rows that drew one emit pad afterwards
    end = int(a[0, 0, 20])
    torch.manual_seed(11)
    e = _gen(model, ids, mask, graph=True, end_ids=[end], pad_id=7, **flat)
    row = e[0, 0, 16:]
    first = int((row == end).nonzero()[0])
    assert bool((row[first + 1:] == 7).all())


@pytest.mark.parametrize("family", ["gpt2", "bloom"])
def test_decode_plan_reuse_on_the_device(family):
    """The captured step is kept between generate() calls (generation._DecodePlan): a second generation with the same
    shapes but other tokens / padding re-initialises the plan's buffers in place, prefills into its KV buffers and
    replays the SAME graph — ids equal to the host loop's; a parameter update drops the plan."""
    if family == "gpt2":
        from cleantransformer_b200.models import modeling_gpt as mg
        cfg = mg.GPTConfig(vocab_size=1000, n_embd=256, n_positions=256, n_layer=3, n_head=4, n_ctx=256, afn="gelu_new")
        model = mg.GPTLMHeadModel(cfg, version="gpt2").to(DEV).eval()
        _init(model)
        model._tie_weights()
    else:
        from cleantransformer_b200.models import modeling_bloom as mb
        cfg = mb.BloomConfig(vocab_size=1000, hidden_size=256, n_layer=3, num_attention_heads=4, hidden_dropout=0.0,
                             attention_dropout=0.0)
        model = mb.BloomForCausalLM(cfg).to(DEV).eval()
        _init(model)
        model._tie_weight()
    outs = []
    for seed in (11, 12, 13):
        ids, mask = _left_padded(5, 24, 1000, seed)
        want = _gen(model, ids, mask, graph=False)
        got = _gen(model, ids, mask, graph=True)
        assert torch.equal(want, got), seed
        assert model._ct_decode_plan_reused == (seed != 11)
        outs.append(got)
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    with torch.no_grad():
        next(p for p in model.parameters() if p.dim() >= 2).mul_(1.01)
    ids, mask = _left_padded(5, 24, 1000, 14)
    assert torch.equal(_gen(model, ids, mask, graph=False), _gen(model, ids, mask, graph=True))
    assert model._ct_decode_plan_reused is False
    # sampling plans are cached too; the generator moves on between generations
    flat = dict(do_sample=True, temperature=30.0, top_k=50, top_p=1.0)
    torch.manual_seed(3)
    a = _gen(model, ids, mask, graph=True, **flat)
    b = _gen(model, ids, mask, graph=True, **flat)
    assert model._ct_decode_plan_reused is True and a.shape == b.shape and not torch.equal(a, b)
