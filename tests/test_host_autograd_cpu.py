"""Host layer (autograd Functions, residual wiring, gradient bookkeeping, tied weights, DDP hooks) executed on CPU
with `tests/mock_ops.py` standing in for the C ABI, against the golden vectors of the REAL reference
(tests/golden/*.pt, tools/make_golden.py). The kernels themselves are covered by the `-m gpu` suite; this file checks
the Python that strings them together — including paths that were written without a GPU at hand."""
import pytest
import torch

import mock_ops
from conftest import rel_err


def _bloom(golden_entry):
    from cleantransformer_b200.models import modeling_bloom as mb
    m = mb.BloomForCausalLM(mb.BloomConfig(**golden_entry["cfg"]))
    m.load_state_dict(golden_entry["sd"], strict=True)
    m._tie_weight()
    m.train()
    return m


def _check_against_golden(m, g, loss, logits, hidden, tol=2e-4):
    assert abs(float(loss) - float(g["loss"])) <= tol * abs(float(g["loss"]))
    assert rel_err(logits, g["logits"]) < tol and rel_err(hidden, g["hidden"]) < tol
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        assert rel_err(p.grad, g["grads"][name]) < 5 * tol, name


@pytest.mark.parametrize("save_act_grad", [False, True], ids=["save_h", "save_gelu_grad"])
def test_bloom_training_step_host_logic_vs_reference_golden(golden, save_act_grad):
    """Fused pre-LN block node, embedding / LM-head tie (two gradient contributions), shifted loss: fp32 on CPU equals
    the reference's own forward/backward. With `save_gelu_grad` the FFN forward stores GELU'(h) and the backward only
    multiplies (functional.SAVE_ACT_GRAD)."""
    from cleantransformer_b200 import functional as F
    g = golden("bloom_tiny")
    with mock_ops.patched():
        prev, F.SAVE_ACT_GRAD = F.SAVE_ACT_GRAD, save_act_grad
        try:
            m = _bloom(g)
            (loss, logits, hidden), kv = m(input_ids=g["ids"], attention_mask=g["mask"], labels=g["labels"])
            loss.backward()
        finally:
            F.SAVE_ACT_GRAD = prev
    _check_against_golden(m, g, loss, logits, hidden)


def test_bloom_fused_lm_head_loss_node_host_logic(golden):
    """functional.LMHeadLossFn (logits GEMM with row statistics + one-pass loss) wires the same gradients as the
    default Linear + cross-entropy nodes; the mock also checks the statistics layout against the stored logits."""
    from cleantransformer_b200 import functional as F
    g = golden("bloom_tiny")
    with mock_ops.patched():
        prev, F.FUSED_LM_STATS = F.FUSED_LM_STATS, True
        try:
            m = _bloom(g)
            (loss, logits, hidden), _ = m(input_ids=g["ids"], attention_mask=g["mask"], labels=g["labels"])
            assert not logits.requires_grad  # an output for the caller's tuple only
            loss.backward()
        finally:
            F.FUSED_LM_STATS = prev
    _check_against_golden(m, g, loss, logits, hidden)


def test_bloom_gradient_accumulation_and_second_step(golden):
    """Two backward passes without zeroing accumulate (every wgrad site honours `accumulate`), and zero_grad
    (set_to_none) followed by another step reproduces the single-step gradients."""
    g = golden("bloom_tiny")
    with mock_ops.patched():
        m = _bloom(g)
        for _ in range(2):
            (loss, _, _), _ = m(input_ids=g["ids"], attention_mask=g["mask"], labels=g["labels"])
            loss.backward()
        for name, p in m.named_parameters():
            assert rel_err(p.grad, 2 * g["grads"][name]) < 1e-3, name
        for p in m.parameters():
            p.grad = None
        (loss, logits, hidden), _ = m(input_ids=g["ids"], attention_mask=g["mask"], labels=g["labels"])
        loss.backward()
    _check_against_golden(m, g, loss, logits, hidden)


# ---------------------------------------------------------------------------------------------------------
# DistributedDataParallel around the Bloom mirror: gradients announced by the autograd Functions themselves
# (functional.grad_written), a tied table with two contributions, buckets launched during backward. gloo, world 2.
# ---------------------------------------------------------------------------------------------------------
def _bloom_ddp_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cleantransformer_b200.ddp import DistributedDataParallel
    from cleantransformer_b200.models import modeling_bloom as mb
    cfg = dict(vocab_size=96, hidden_size=32, n_layer=2, num_attention_heads=4)
    with mock_ops.patched():
        torch.manual_seed(50 + rank)  # different init per rank: the wrapper syncs from rank 0
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m._tie_weight(); m.train()
        ddp = DistributedDataParallel(m, bucket_cap_mb=0.01)
        assert len(ddp.buckets) >= 3
        torch.manual_seed(70 + rank)
        ids = torch.randint(3, 96, (2, 12))
        mask = torch.ones_like(ids)
        for _ in range(2):  # the second step re-arms the per-bucket counters
            for p in m.parameters():
                p.grad = None
            (loss, _, _), _ = ddp(input_ids=ids, attention_mask=mask, labels=ids)
            loss.backward()
        reduced = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        # local gradients of an identical, unwrapped copy
        ref = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        ref.load_state_dict({k[len("module."):]: v for k, v in ddp.state_dict().items()})
        ref._tie_weight(); ref.train()
        (l2, _, _), _ = ref(input_ids=ids, attention_mask=mask, labels=ids)
        l2.backward()
        local = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
    out[rank] = (reduced, local)
    dist.barrier()
    dist.destroy_process_group()


def test_bloom_ddp_world2_gloo_reduces_every_gradient_once():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bloom_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    (r0, l0), (r1, l1) = out[0], out[1]
    for n in r0:
        assert torch.allclose(r0[n], r1[n]), n                  # same reduced gradient on both ranks
        assert rel_err(r0[n], (l0[n] + l1[n]) / 2) < 1e-5, n      # = the mean of the local ones (tied table included)


# ---------------------------------------------------------------------------------------------------------
# The other mirrors, same idea: GPT (Conv1D weights [in,out], blocked QKV split, post-LN 'gpt' and pre-LN 'gpt2'
# wiring, greedy generation with the KV cache), BERT (separate q/k/v Linears, additive -1e4 mask, pooler), the
# generic TransformerBlock of transformer.py.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("version", ["gpt", "gpt2"])
def test_gpt_host_logic_vs_reference_golden(golden, version, monkeypatch):
    from cleantransformer_b200 import ops
    from cleantransformer_b200.models import modeling_gpt as mg
    g = golden("gpt_tiny")
    c, cfg = g[version], g["cfg"]
    with mock_ops.patched():
        model = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version=version)
        model.load_state_dict(c["sd"], strict=True)
        model._tie_weights()
        model.eval()
        with torch.no_grad():
            (logits, hidden), _ = model(c["ids"], attention_mask=c["mask"])
        assert rel_err(logits, c["logits"]) < 2e-4 and rel_err(hidden, c["hidden"]) < 2e-4
        gen = model.generate(c["ids"], attention_mask=c["mask"],
                             generation_configs={"beam_size": 1, "do_sample": False, "max_gen_len": 6,
                                                 "end_ids": None, "pad_id": 0, "no_repeat_ngram_size": 0})
        assert torch.equal(gen, c["generated"])
        blk = model.gpt.blocks[0]
        x = c["blk_x"].clone().requires_grad_(True)
        model.zero_grad()
        y, _ = blk(x)
        y.backward(c["blk_dy"])
    assert rel_err(y, c["blk_y"]) < 2e-4 and rel_err(x.grad, c["blk_dx"]) < 1e-3
    for name, p in blk.named_parameters():
        assert rel_err(p.grad, c["blk_grads"][name]) < 1e-3, name


def test_bert_and_generic_block_host_logic_vs_reference_golden(golden):
    from cleantransformer_b200 import transformer as T
    from cleantransformer_b200.models import modeling_bert as mbert
    g = golden("bert_tiny")
    with mock_ops.patched():
        model = mbert.BertForSequenceClassification(mbert.BertConfig(**dict(g["cfg"]))).eval()
        model.load_state_dict(g["sd"], strict=True)
        with torch.no_grad():
            logits = model(g["ids"], g["mask"], g["seg"], g["pos"])
            hidden, pooled = model.bert(g["ids"], g["mask"], g["seg"], g["pos"])
        assert rel_err(hidden, g["hidden"]) < 2e-4 and rel_err(pooled, g["pooled"]) < 2e-4
        assert rel_err(logits, g["logits"]) < 2e-4
        gb = golden("generic_block")
        blk = T.TransformerBlock(T.ExampleConfig()).eval()
        blk.load_state_dict(gb["sd"], strict=True)
        with torch.no_grad():
            y = blk(gb["x"])
            a = blk.attention(gb["x"], (1.0 - gb["mask"][:, None, None, :]) * -10000.0)
    assert rel_err(y, gb["y"]) < 2e-4 and rel_err(a, gb["att_masked"]) < 2e-4


def test_optimizer_host_logic_matches_reference_trajectories(golden):
    """The optimizer classes' host side — flat-arena creation (parameters re-pointed into one buffer), the choice
    between the flat and the per-tensor launch, step counters, generator input, momentum buffers — against the
    trajectories of the reference's own classes and of torch.optim.AdamW (tests/golden/optim.pt)."""
    from cleantransformer_b200 import optimizer as opt
    g = golden("optim")

    def run(make):
        ps = [p.clone().requires_grad_(True) for p in g["p0"]]
        o = make(ps)
        traj = []
        for step_g in g["grads"]:
            for p, gr in zip(ps, step_g):
                p.grad = gr.clone()
            o.step()
            traj.append([p.detach().clone() for p in ps])
        return traj, o

    def check(traj, ref, tol=1e-5):
        for a, b in zip(traj, ref):
            for x, y in zip(a, b):
                assert rel_err(x, y) < tol

    with mock_ops.patched():
        t, o = run(lambda ps: opt.AdamW(ps, lr=0.01, weight_decay=0.01))
        check(t, g["ref_adamw"])
        for m, mr in zip(o.momentum_buffer, g["ref_adamw_m"]):
            assert rel_err(m, mr) < 1e-5
        t, _ = run(lambda ps: opt.AdamW(iter(ps), lr=0.01))
        check(t, g["ref_adamw_nowd"])
        t, o = run(lambda ps: opt.TorchAdamW(ps, lr=0.01, weight_decay=0.01))
        check(t, g["torch_adamw"])
        assert o._arena is not None  # the per-tensor gradients here are not arena views: multi-tensor path
        t, _ = run(lambda ps: opt.SGD(ps, lr=0.01, weight_decay=0.01, momentum=0.9))
        check(t, g["ref_sgd"])
        t, _ = run(lambda ps: opt.SGD(ps, lr=0.01))
        check(t, g["ref_sgd_plain"])
        # flat-arena path: gradients written into the arena views (what the wgrad kernels do)
        ps = [p.clone().requires_grad_(True) for p in g["p0"]]
        o = opt.TorchAdamW(ps, lr=0.01, weight_decay=0.01)
        o._setup()
        for k, step_g in enumerate(g["grads"]):
            for p, gr in zip(ps, step_g):
                p._ct_grad_view.copy_(gr)
                p.grad = p._ct_grad_view
            assert o._arena.grads_complete()
            o.step()
            for p, ref in zip(ps, g["torch_adamw"][k]):
                assert rel_err(p.detach(), ref) < 1e-5


def test_bf16_shadow_weights_stay_coherent_across_optimizer_steps(golden):
    """The tensor-core operands are bf16 shadows of the fp32 parameters, cached per parameter and refreshed by the
    fused AdamW pass (arena.shadow, functional.shadow). Four training steps with the cache must equal, bit for bit,
    four steps where every shadow is thrown away after each optimizer step (a stale shadow would show up as a
    different trajectory). Also: from the second step on the gradients are arena views and the flat kernel path runs."""
    from cleantransformer_b200.optimizer import TorchAdamW
    g = golden("bloom_tiny")

    def run(invalidate):
        with mock_ops.patched(compute_dtype=torch.bfloat16):
            m = _bloom(g)
            o = TorchAdamW(m.parameters(), lr=1e-2)
            flat_steps = 0
            for _ in range(4):
                o.zero_grad()
                (loss, _, _), _ = m(input_ids=g["ids"], attention_mask=g["mask"], labels=g["labels"])
                loss.backward()
                flat_steps += int(o._arena is not None and o._arena.grads_complete())
                o.step()
                if invalidate:
                    for p in m.parameters():
                        p._ct_shadow_ver = -1
                        p._ct_shadow = None
            return [p.detach().clone() for p in m.parameters()], float(loss), flat_steps

    a, la, flat_a = run(False)
    b, lb, _ = run(True)
    assert flat_a >= 3
    assert la == lb
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_trainer_loop_trains_the_bloom_mirror(golden, tmp_path):
    """Trainer.train / evaluate / save_model (the reference's trainer surface, trainer/trainer.py:1303-1511) around
    the Bloom mirror + TorchAdamW: the loss of a memorisable toy set falls, evaluation runs without gradients, the
    checkpoint reloads into a fresh model with identical logits."""
    import types
    from cleantransformer_b200.models import modeling_bloom as mb
    from cleantransformer_b200.trainer import Trainer
    cfg = dict(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4)
    torch.manual_seed(3)
    data = [dict(input_ids=torch.randint(3, 64, (10,)), attention_mask=torch.ones(10, dtype=torch.long)) for _ in range(8)]
    for d in data:
        d["labels"] = d["input_ids"].clone()

    def collate(items):
        return {k: torch.stack([it[k] for it in items]) for k in items[0]}

    with mock_ops.patched():
        torch.manual_seed(4)
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m._tie_weight()
        args = types.SimpleNamespace(per_device_train_batch_size=4, per_device_eval_batch_size=4, learning_rate=5e-3,
                                     max_steps=12, logging_steps=2, output_dir=str(tmp_path), weight_decay=0.0)
        tr = Trainer(model=m, args=args, data_collator=collate, train_dataset=data, eval_dataset=data)
        before = tr.evaluate()["eval_loss"]
        out = tr.train()
        after = tr.evaluate()["eval_loss"]
        assert out.global_step == 12 and len(tr.state.log_history) == 6
        assert after < 0.8 * before, (before, after)
        tr.save_model()
        m2 = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m2.load_state_dict(torch.load(str(tmp_path / "pytorch_model.bin")), strict=True)
        m2._tie_weight(); m.eval(); m2.eval()
        with torch.no_grad():
            (lg1, _), _ = m(input_ids=data[0]["input_ids"][None], attention_mask=data[0]["attention_mask"][None])
            (lg2, _), _ = m2(input_ids=data[0]["input_ids"][None], attention_mask=data[0]["attention_mask"][None])
        assert torch.equal(lg1, lg2)


@pytest.mark.parametrize("pad", ["left", "right"])
def test_bloom_padding_semantics_and_greedy_generation_vs_oracle(golden, pad):
    """Left / right padded batches through the Bloom mirror (ALiBi positions from the mask's cumulative sum,
    masked_fill(finfo.min) on padded keys, fully masked query rows on left padding) against the oracle restatement
    of modeling_bloom.py, plus greedy decoding with the KV cache against the oracle's loop (token ids bit-exact)."""
    from oracle import ct_oracle as O
    g = golden("bloom_tiny")
    cfg = g["cfg"]
    torch.manual_seed(31)
    B, S = 3, 9
    ids = torch.randint(3, cfg["vocab_size"], (B, S))
    mask = torch.ones(B, S, dtype=torch.long)
    for b, n in enumerate((0, 3, 5)):
        if n:
            if pad == "left":
                mask[b, :n] = 0
            else:
                mask[b, S - n:] = 0
    sd = {k: v for k, v in g["sd"].items() if k != "lm_head.weight"}
    with torch.no_grad():
        (lg_ref, h_ref), _ = O.bloom_causal_lm(ids, mask, sd, cfg["n_layer"], cfg["num_attention_heads"],
                                               cfg["layer_norm_epsilon"])
    with mock_ops.patched():
        m = _bloom(g).eval()
        with torch.no_grad():
            (lg, h), _ = m(input_ids=ids, attention_mask=mask)
        valid = mask.bool()
        assert rel_err(lg[valid], lg_ref[valid]) < 2e-4
        assert rel_err(lg, lg_ref) < 2e-4  # padded positions too: the reference's finite-fill arithmetic
        if pad == "left":  # generation recipe of the reference (inference_bloom.py: left padding)
            gen = m.generate(ids, attention_mask=mask,
                             generation_configs={"beam_size": 1, "do_sample": False, "max_gen_len": 5,
                                                 "end_ids": None, "pad_id": 3, "no_repeat_ngram_size": 0})

            def step(x, am, caches):
                return O.bloom_causal_lm(x, am, sd, cfg["n_layer"], cfg["num_attention_heads"],
                                         cfg["layer_norm_epsilon"], k_v_pasts=caches)

            with torch.no_grad():
                ref = O.greedy_generate(step, ids, mask, cfg["n_layer"], max_gen_len=5, pad_id=3)
            assert torch.equal(gen.view(ref.shape), ref)


def test_bloom_post_layernorm_residual_switch_vs_oracle(golden):
    """BloomBlock's `apply_residual_connection_post_layernorm` switch (modeling_bloom.py:142-159): the residual is
    the LayerNorm OUTPUT instead of its input — the un-fused block path (the fused pre-LN node does not apply).
    Training step (loss, every gradient) against the oracle's autograd on the same weights."""
    from cleantransformer_b200.models import modeling_bloom as mb
    from oracle import ct_oracle as O
    g = golden("bloom_tiny")
    cfg = dict(g["cfg"]); cfg["apply_residual_connection_post_layernorm"] = True
    ids, mask, labels = g["ids"], g["mask"], g["labels"]
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["sd"].items() if k != "lm_head.weight"}
    (l_ref, lg_ref, _), _ = O.bloom_causal_lm(ids, mask, sd, cfg["n_layer"], cfg["num_attention_heads"],
                                              cfg["layer_norm_epsilon"], labels=labels, training=True,
                                              post_ln_residual=True)
    l_ref.backward()
    with mock_ops.patched():
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m.load_state_dict(g["sd"], strict=True)
        m._tie_weight(); m.train()
        (loss, logits, _), _ = m(input_ids=ids, attention_mask=mask, labels=labels)
        loss.backward()
    assert abs(float(loss) - float(l_ref)) <= 2e-4 * abs(float(l_ref))
    assert rel_err(logits, lg_ref) < 2e-4
    for name, p in m.named_parameters():
        key = "bloom.word_embeddings.weight" if name == "lm_head.weight" else name
        assert rel_err(p.grad, sd[key].grad) < 1e-3, name


def test_bert_classifier_gradients_vs_oracle(golden):
    """BERT fine-tuning direction (BASELINE config 5): external cross entropy on the classifier logits, backward
    through the pooler (tanh), erf-GELU FFNs, post-LN blocks, separate q/k/v Linears with the additive -1e4 mask and
    the three embedding tables — every gradient against the oracle's autograd (dropout off: eval())."""
    from cleantransformer_b200.models import modeling_bert as mbert
    from oracle import ct_oracle as O
    g = golden("bert_tiny")
    cfg = dict(g["cfg"])
    ids, mask, seg, pos = g["ids"], g["mask"], g["seg"], g["pos"]
    torch.manual_seed(5)
    labels = torch.randint(0, cfg["num_labels"], (ids.shape[0],))
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["sd"].items()}
    lg_ref, _, _ = O.bert_classifier(ids, mask, seg, pos, sd, cfg["num_hidden_layers"], cfg["num_attention_heads"],
                                     cfg["layer_norm_eps"])
    torch.nn.functional.cross_entropy(lg_ref, labels).backward()
    with mock_ops.patched():
        model = mbert.BertForSequenceClassification(mbert.BertConfig(**cfg)).eval()
        model.load_state_dict(g["sd"], strict=True)
        logits = model(ids, mask, seg, pos)
        torch.nn.functional.cross_entropy(logits.float(), labels).backward()
    assert rel_err(logits, lg_ref) < 2e-4
    checked = 0
    for name, p in model.named_parameters():
        ref = sd[name].grad
        if ref is None:
            continue
        assert p.grad is not None, name
        # (the key bias gradient is analytically zero — softmax is shift invariant — hence the absolute floor)
        assert float((p.grad - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-9, name
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("version", ["gpt", "gpt2"])
def test_gpt_lm_gradients_vs_oracle(golden, version):
    """GPT training direction: external shifted cross entropy on the LM logits, backward through the tied head, all
    blocks (post-LN 'gpt' / pre-LN 'gpt2'), Conv1D weights and BOTH embedding tables, left-padded batch — every
    gradient against the oracle's autograd. eval(): the reference's Dropout(0.5) (modeling_gpt.py:136) is off."""
    from cleantransformer_b200.models import modeling_gpt as mg
    from oracle import ct_oracle as O
    g = golden("gpt_tiny")
    c, cfg = g[version], g["cfg"]
    ids, mask = c["ids"], c["mask"]

    def lm_loss(logits):
        return torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), ids[:, 1:].reshape(-1))

    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "attn.bias" not in k) for k, v in c["sd"].items()}
    (lg_ref, _), _ = O.gpt_lm_head_model(ids, mask, sd, cfg["n_layer"], cfg["n_head"], cfg["n_ctx"],
                                         cfg["layer_norm_epsilon"], version=version)
    lm_loss(lg_ref).backward()
    with mock_ops.patched():
        model = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version=version)
        model.load_state_dict(c["sd"], strict=True)
        model._tie_weights()
        model.eval()
        (logits, _), _ = model(ids, attention_mask=mask)
        lm_loss(logits).backward()
    assert rel_err(logits, lg_ref) < 2e-4
    checked = 0
    for name, p in model.named_parameters():
        key = "gpt.tokens_embed.weight" if name == "lm_head.weight" else name
        ref = sd[key].grad if key in sd else None
        if ref is None:
            continue
        assert p.grad is not None, name
        assert float((p.grad - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-9, name
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("pad", ["none", "right"])
def test_bloom_block_accepts_the_reference_argument_tensors(golden, pad):
    """BloomBlock.forward(hidden, attention_mask=bool [b,1,q,k], alibi=[b*h,1,k]) — the reference's own calling
    convention (modeling_bloom.py:142-159, tensors built by BloomModel._attn_mask / build_alibi_tensor) — goes
    through AttnBias.from_reference_args; output and both cache tensors against the oracle's block."""
    from cleantransformer_b200.models import modeling_bloom as mb
    from oracle import ct_oracle as O
    g = golden("bloom_tiny")
    cfg = g["cfg"]
    nh = cfg["num_attention_heads"]
    torch.manual_seed(41)
    B, S, H = 2, 10, cfg["hidden_size"]
    x = torch.randn(B, S, H)
    mask = torch.ones(B, S, dtype=torch.long)
    if pad == "right":
        mask[1, 7:] = 0
    alibi = O.build_alibi_tensor(mask, nh, torch.float32)
    mask_bool = O.bloom_attn_mask(mask, (B, S))
    sd = g["sd"]
    with torch.no_grad():
        ref, (k_ref, v_ref) = O.bloom_block(x, mask_bool, alibi, sd, "bloom.blocks.0.", nh, cfg["layer_norm_epsilon"])
    with mock_ops.patched():
        m = _bloom(g).eval()
        with torch.no_grad():
            out, (k, v) = m.bloom.blocks[0](x, attention_mask=mask_bool, alibi=alibi, head_mask=None)
    valid = mask.bool()
    assert rel_err(out[valid], ref[valid]) < 2e-4
    assert rel_err(k, k_ref) < 2e-4 and rel_err(v, v_ref) < 2e-4


def _bloom_ddp_nosync_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cleantransformer_b200.ddp import DistributedDataParallel
    from cleantransformer_b200.models import modeling_bloom as mb
    cfg = dict(vocab_size=96, hidden_size=32, n_layer=1, num_attention_heads=4)
    with mock_ops.patched():
        torch.manual_seed(60)
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m._tie_weight(); m.train()
        ddp = DistributedDataParallel(m, bucket_cap_mb=0.01)
        torch.manual_seed(80 + rank)
        batches = [torch.randint(3, 96, (2, 8)) for _ in range(2)]
        local = []
        for ids in batches:  # per-micro-batch local gradients of the same weights
            for p in m.parameters():
                p.grad = None
            with ddp.no_sync():
                (l, _, _), _ = ddp(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids)
                l.backward()
            local.append({n: p.grad.detach().clone() for n, p in m.named_parameters()})
        for p in m.parameters():
            p.grad = None
        with ddp.no_sync():  # micro-batch 1: accumulate locally, nothing is exchanged
            (l, _, _), _ = ddp(input_ids=batches[0], attention_mask=torch.ones_like(batches[0]), labels=batches[0])
            l.backward()
        (l, _, _), _ = ddp(input_ids=batches[1], attention_mask=torch.ones_like(batches[1]), labels=batches[1])
        l.backward()               # micro-batch 2: accumulated gradients are averaged over the ranks
        got = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    out[rank] = (got, {n: local[0][n] + local[1][n] for n in local[0]})
    dist.barrier()
    dist.destroy_process_group()


def test_bloom_ddp_gradient_accumulation_with_no_sync():
    """ft_bloom_DDP-style gradient accumulation: micro-batches under `no_sync()` stay local, the synchronising
    backward reduces the ACCUMULATED gradients — mean over ranks of each rank's sum, tied table included."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bloom_ddp_nosync_worker, args=(2, port, out), nprocs=2, join=True)
    (g0, s0), (g1, s1) = out[0], out[1]
    for n in g0:
        assert torch.allclose(g0[n], g1[n]), n
        assert rel_err(g0[n], (s0[n] + s1[n]) / 2) < 1e-5, n


def _gpt_segment_ddp_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cleantransformer_b200.ddp import DistributedDataParallel
    from cleantransformer_b200.models import modeling_gpt as mg
    cfg = dict(vocab_size=80, n_embd=32, n_positions=16, n_layer=2, n_head=4, n_ctx=16, embd_pdrop=0.0,
               attn_pdrop=0.0, resid_pdrop=0.0)
    with mock_ops.patched():
        torch.manual_seed(11)
        m = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version="gpt2").eval()  # eval: Dropout(0.5) of the MLP (gpt:136)
        ddp = DistributedDataParallel(m, bucket_cap_mb=0.005)
        torch.manual_seed(90 + rank)
        ids = torch.randint(1, 80, (2, 10))
        seg = torch.randint(1, 80, (2, 10))       # modeling_gpt.py:186-188: segment ids index tokens_embed again
        mask = torch.ones_like(ids)

        def loss_of(model):
            (logits, _), _ = model(ids, attention_mask=mask, segment_ids=seg)
            return torch.nn.functional.cross_entropy(logits.float().view(-1, 80), ids.view(-1))

        for _ in range(2):
            for p in m.parameters():
                p.grad = None
            loss_of(ddp).backward()
        table = m.gpt.tokens_embed.weight
        assert table._ct_uses == 3                # lm_head + token lookup + segment lookup, counted in forward
        reduced = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        ref = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version="gpt2").eval()
        ref.load_state_dict({k[len("module."):]: v for k, v in ddp.state_dict().items()})
        ref._tie_weights()
        loss_of(ref).backward()
        local = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
    out[rank] = (reduced, local)
    dist.barrier()
    dist.destroy_process_group()


def test_gpt_segment_ids_ddp_world2_gloo_counts_three_writes_of_the_tied_table():
    """with `segment_ids` GPTModel looks tokens_embed up twice, so the tied table receives three gradient
    writes; the wrapper must reduce it after the third (a static `_ct_expected_writes = 2` reduced it one write
    early and the last contribution stayed local)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gpt_segment_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    (r0, l0), (r1, l1) = out[0], out[1]
    for n in r0:
        assert torch.allclose(r0[n], r1[n]), n
        assert rel_err(r0[n], (l0[n] + l1[n]) / 2) < 1e-5, n


@pytest.mark.parametrize("family", ["gpt2", "bloom"])
def test_captured_decode_host_logic_matches_the_loop_and_the_oracle(family, monkeypatch):
    """generation._graphed_greedy with the device pieces mocked: static KV buffers sized once, the
    full-capacity mask, device-side cache length / write column / alive flags / position ids, polling and trimming.
    Its ids must equal the un-captured loop's and the oracle's restatement of generation_util.py:57-119, with and
    without end ids (rows finishing at different steps; every row finishing at once)."""
    from cleantransformer_b200 import generation
    from oracle import ct_oracle as O
    torch.manual_seed(77)
    V, L, NH = 97, 2, 2
    B, P = 4, 7
    ids = torch.randint(3, V, (B, P))
    mask = torch.ones(B, P, dtype=torch.long)
    for b, n in enumerate((0, 2, 4, 1)):  # LEFT padding (examples/inference_gpt2.py:55-59)
        mask[b, :n] = 0
        ids[b, :n] = 0
    with mock_ops.patched():
        if family == "gpt2":
            from cleantransformer_b200.models import modeling_gpt as mg
            m = mg.GPTLMHeadModel(mg.GPTConfig(vocab_size=V, n_embd=64, n_positions=64, n_layer=L, n_head=NH, n_ctx=64,
                                               afn="gelu_new"), version="gpt2").eval()
            sd = {k: v.detach() for k, v in m.state_dict().items()}

            def step(x, am, caches):
                return O.gpt_lm_head_model(x, am, sd, L, NH, 64, 1e-5, version="gpt2", k_v_pasts=caches)
        else:
            from cleantransformer_b200.models import modeling_bloom as mb
            m = mb.BloomForCausalLM(mb.BloomConfig(vocab_size=V, hidden_size=64, n_layer=L, num_attention_heads=NH)).eval()
            m._tie_weight()
            with torch.no_grad():
                for p in m.parameters():
                    if p.dim() >= 2:
                        p.normal_(0, 0.2)
            sd = {k: v.detach() for k, v in m.state_dict().items() if k != "lm_head.weight"}

            def step(x, am, caches):
                return O.bloom_causal_lm(x, am, sd, L, NH, 1e-5, k_v_pasts=caches)

        def gen(graph, **cfg):
            monkeypatch.setenv("CT_DECODE_GRAPH", "1" if graph else "0")
            m._ct_decode_graph_launches = -1
            base = {"beam_size": 1, "do_sample": False, "max_gen_len": 9, "end_ids": None, "pad_id": 1}
            base.update(cfg)
            out = m.generate(ids, attention_mask=mask, generation_configs=base)
            assert (m._ct_decode_graph_launches >= 0) == graph
            return out

        loop, cap = gen(False), gen(True)
        assert cap.shape == (B, 1, P + 11) and torch.equal(loop, cap)
        with torch.no_grad():
            ref = O.greedy_generate(step, ids, mask, L, max_gen_len=9, pad_id=1)
        assert torch.equal(cap.view(ref.shape), ref)
        new = loop[:, 0, P:]
        for ends in ([int(new[0, 2]), int(new[2, 6])], sorted(set(new[:, 1].tolist())), [int(new[1, 0])]):
            a, b = gen(False, end_ids=ends), gen(True, end_ids=ends)
            assert a.shape == b.shape and torch.equal(a, b), ends
            with torch.no_grad():
                r = O.greedy_generate(step, ids, mask, L, max_gen_len=9, pad_id=1, end_ids=ends)
            assert torch.equal(b.view(r.shape), r), ends
        assert torch.equal(gen(False, max_gen_len=1), gen(True, max_gen_len=1))
        # POLL_EVERY smaller than the generation: the early stop is noticed at a poll, the result is trimmed at done_at
        monkeypatch.setattr(generation, "POLL_EVERY", 2)
        ends = sorted(set(new[:, 4].tolist()))
        assert torch.equal(gen(False, end_ids=ends), gen(True, end_ids=ends))


# ------------------------------------------------------------------------------------------------------------------
# Dropout in train mode: the host wiring — which site draws which stream of the counter-based
# generator, the fused residual add, the backward regenerating the same mask — against the oracle given the SAME masks
# (oracle.DropoutFeeder). The kernels themselves are checked on the GPU (tests/test_gpu_dropout.py).
# ------------------------------------------------------------------------------------------------------------------
def _grads_match(model, sd, rename=lambda n: n, tol=2e-3, at_least=10, floor=1e-9):
    checked = 0
    for name, p in model.named_parameters():
        key = rename(name)
        ref = sd[key].grad if key in sd else None
        if ref is None:
            continue
        assert p.grad is not None, name
        assert float((p.grad - ref).abs().max()) <= tol * float(ref.abs().max()) + floor, name
        checked += 1
    assert checked >= at_least


def test_bert_train_mode_dropout_vs_oracle_with_the_same_masks(golden):
    """BASELINE config 5 as SURVEY d2 specifies it: BERT with hidden / attention dropout p = 0.1 in train mode."""
    from cleantransformer_b200 import functional as F
    from cleantransformer_b200.models import modeling_bert as mbert
    from oracle import ct_oracle as O
    g = golden("bert_tiny")
    cfg = dict(g["cfg"])
    cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"] = 0.1, 0.2
    ids, mask, seg, pos = g["ids"], g["mask"], g["seg"], g["pos"]
    torch.manual_seed(5)
    labels = torch.randint(0, cfg["num_labels"], (ids.shape[0],))
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["sd"].items()}
    lg_ref, _, _ = O.bert_classifier(ids, mask, seg, pos, sd, cfg["num_hidden_layers"], cfg["num_attention_heads"],
                                     cfg["layer_norm_eps"], drop=O.DropoutFeeder(777), p_attn=0.2, p_hidden=0.1)
    torch.nn.functional.cross_entropy(lg_ref, labels).backward()
    lg_eval, _, _ = O.bert_classifier(ids, mask, seg, pos, g["sd"], cfg["num_hidden_layers"],
                                      cfg["num_attention_heads"], cfg["layer_norm_eps"])
    assert rel_err(lg_ref.detach(), lg_eval) > 1e-2  # the masks do something
    with mock_ops.patched():
        model = mbert.BertForSequenceClassification(mbert.BertConfig(**cfg)).train()
        model.load_state_dict(g["sd"], strict=True)
        F.manual_dropout_seed(777)
        logits = model(ids, mask, seg, pos)
        torch.nn.functional.cross_entropy(logits.float(), labels).backward()
        assert rel_err(logits, lg_ref) < 2e-4
        _grads_match(model, sd, at_least=20)
        # a second forward draws new masks (the stream counter moved on); eval() is the deterministic network again
        assert rel_err(model(ids, mask, seg, pos), lg_ref) > 1e-3
        assert rel_err(model.eval()(ids, mask, seg, pos), lg_eval) < 2e-4


@pytest.mark.parametrize("version", ["gpt", "gpt2"])
def test_gpt_train_mode_dropout_vs_oracle_with_the_same_masks(golden, version):
    """GPT in train mode as the reference constructs it: embd / attn / resid dropout from the config and the MLP's
    torch.nn.Dropout() with its default p = 0.5 (modeling_gpt.py:136)."""
    from cleantransformer_b200 import functional as F
    from cleantransformer_b200.models import modeling_gpt as mg
    from oracle import ct_oracle as O
    g = golden("gpt_tiny")
    c, cfg = g[version], dict(g["cfg"])
    cfg["embd_pdrop"], cfg["attn_pdrop"], cfg["resid_pdrop"] = 0.1, 0.15, 0.2
    ids, mask = c["ids"], c["mask"]

    def lm_loss(logits):
        return torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), ids[:, 1:].reshape(-1))

    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "attn.bias" not in k) for k, v in c["sd"].items()}
    (lg_ref, _), _ = O.gpt_lm_head_model(ids, mask, sd, cfg["n_layer"], cfg["n_head"], cfg["n_ctx"],
                                         cfg["layer_norm_epsilon"], version=version, drop=O.DropoutFeeder(4321),
                                         p_embd=0.1, p_attn=0.15, p_resid=0.2, p_mlp=0.5)
    lm_loss(lg_ref).backward()
    with mock_ops.patched():
        model = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version=version)
        model.load_state_dict(c["sd"], strict=True)
        model._tie_weights()
        model.train()
        F.manual_dropout_seed(4321)
        (logits, _), _ = model(ids, attention_mask=mask)
        lm_loss(logits).backward()
    assert rel_err(logits, lg_ref) < 2e-4
    _grads_match(model, sd, rename=lambda n: "gpt.tokens_embed.weight" if n == "lm_head.weight" else n, at_least=20)


def test_bloom_and_generic_block_train_mode_dropout_vs_oracle_with_the_same_masks(golden):
    """Bloom with hidden_dropout / attention_dropout > 0 (modeling_bloom.py:111-113, :121-123, :269: the un-fused
    block path) and transformer.py's TransformerBlock (:47-50, :109, :115)."""
    from cleantransformer_b200 import functional as F
    from cleantransformer_b200 import transformer as T
    from cleantransformer_b200.models import modeling_bloom as mb
    from oracle import ct_oracle as O
    g = golden("bloom_tiny")
    cfg = dict(g["cfg"])
    cfg["hidden_dropout"], cfg["attention_dropout"] = 0.1, 0.25
    ids, mask, labels = g["ids"], g["mask"], g["labels"]
    sd = {k: v.clone().requires_grad_(True) for k, v in g["sd"].items() if k != "lm_head.weight"}
    (l_ref, lg_ref, _), _ = O.bloom_causal_lm(ids, mask, sd, cfg["n_layer"], cfg["num_attention_heads"],
                                             cfg["layer_norm_epsilon"], labels=labels, training=True,
                                             drop=O.DropoutFeeder(99), p_attn=0.25, p_hidden=0.1)
    l_ref.backward()
    with mock_ops.patched():
        model = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        model.load_state_dict(g["sd"], strict=True)
        model._tie_weight()
        model.train()
        F.manual_dropout_seed(99)
        (loss, logits, _), _ = model(input_ids=ids, attention_mask=mask, labels=labels)
        loss.backward()
        assert abs(float(loss) - float(l_ref)) / float(l_ref) < 1e-4 and rel_err(logits, lg_ref) < 2e-4
        _grads_match(model, sd, rename=lambda n: "bloom.word_embeddings.weight" if n == "lm_head.weight" else n,
                     at_least=20)
        # transformer.py TransformerBlock (ExampleConfig: both probabilities 0.1)
        gb = golden("generic_block")
        cfg_b = T.ExampleConfig()
        sdb = {k: v.clone().requires_grad_(True) for k, v in gb["sd"].items()}
        xr = gb["x"].clone().requires_grad_(True)
        y_ref = O.generic_block(xr, sdb, cfg_b.num_attention_heads, cfg_b.layer_norm_epsilong, drop=O.DropoutFeeder(5),
                                p_attn=cfg_b.attention_probs_dropout_prob, p_hidden=cfg_b.hidden_dropout_prob)
        dy = torch.randn_like(y_ref)
        y_ref.backward(dy)
        blk = T.TransformerBlock(cfg_b).train()
        blk.load_state_dict(gb["sd"], strict=True)
        x = gb["x"].clone().requires_grad_(True)
        F.manual_dropout_seed(5)
        y = blk(x)
        y.backward(dy)
        assert rel_err(y, y_ref) < 2e-4 and rel_err(x.grad, xr.grad) < 2e-3
        _grads_match(blk, sdb, at_least=8, floor=1e-6)  # (the key bias gradient is analytically zero)


def test_generate_on_a_train_mode_model_takes_the_host_loop_with_dropout_active(golden, monkeypatch):
    """generation_util.py:57-119 does not switch the model to eval(): called on a model in train mode the reference
    decodes with its dropout active. The mirror then keeps the host loop (the captured step refuses dropout) and feeds
    the attention dropout into the cached-attention calls; same masks -> same ids as the oracle."""
    from cleantransformer_b200 import functional as F
    from cleantransformer_b200.models import modeling_gpt as mg
    from oracle import ct_oracle as O
    torch.manual_seed(5)
    V, L, NH = 61, 2, 2
    cfg = dict(vocab_size=V, n_embd=64, n_positions=64, n_layer=L, n_head=NH, n_ctx=64, afn="gelu_new",
               embd_pdrop=0.1, attn_pdrop=0.2, resid_pdrop=0.1)
    ids = torch.randint(3, V, (2, 5))
    mask = torch.ones(2, 5, dtype=torch.long)
    mask[1, :2] = 0
    with mock_ops.patched():
        m = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version="gpt2").train()
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        feeder = O.DropoutFeeder(2024)

        def step(x, am, caches):
            return O.gpt_lm_head_model(x, am, sd, L, NH, 64, 1e-5, version="gpt2", k_v_pasts=caches, drop=feeder,
                                       p_embd=0.1, p_attn=0.2, p_resid=0.1, p_mlp=0.5)

        with torch.no_grad():
            ref = O.greedy_generate(step, ids, mask, L, max_gen_len=4, pad_id=0)
        F.manual_dropout_seed(2024)
        m._ct_decode_graph_launches = -1
        out = m.generate(ids, attention_mask=mask, generation_configs={"beam_size": 1, "do_sample": False,
                                                                      "max_gen_len": 4, "end_ids": None, "pad_id": 0})
        assert m._ct_decode_graph_launches == -1, "train mode must not take the captured step"
        assert torch.equal(out.view(ref.shape), ref)


def test_captured_decode_with_sampling_host_logic(monkeypatch):
    """do_sample=True (the reference's default, generation_util.py:74-84) through the captured step: the draw (torch's
    processors + multinomial) happens inside the step and ct_greedy_step only does the bookkeeping. With the device
    pieces mocked both paths consume the CPU generator identically, so the same seed must give the same ids; top_k = 1
    makes the draw deterministic: the greedy ids."""
    from cleantransformer_b200.models import modeling_gpt as mg
    torch.manual_seed(123)
    V, L, NH = 83, 2, 2
    ids = torch.randint(3, V, (3, 6))
    mask = torch.ones(3, 6, dtype=torch.long)
    mask[2, :3] = 0
    with mock_ops.patched():
        m = mg.GPTLMHeadModel(mg.GPTConfig(vocab_size=V, n_embd=64, n_positions=64, n_layer=L, n_head=NH, n_ctx=64,
                                           afn="gelu_new"), version="gpt2").eval()
        with torch.no_grad():
            m.gpt.tokens_embed.weight.mul_(0.1)  # (N(0,1) tied embeddings make the LM head all but deterministic)

        def gen(graph, seed, **cfg):
            monkeypatch.setenv("CT_DECODE_GRAPH", "1" if graph else "0")
            m._ct_decode_graph_launches = -1
            base = {"beam_size": 1, "do_sample": True, "max_gen_len": 8, "end_ids": None, "pad_id": 0,
                    "temperature": 4.0, "top_k": 30, "top_p": 0.97}   # (flat enough for seeds to matter)
            base.update(cfg)
            torch.manual_seed(seed)
            out = m.generate(ids, attention_mask=mask, generation_configs=base)
            assert (m._ct_decode_graph_launches >= 0) == graph
            return out

        a, b = gen(False, 7), gen(True, 7)
        assert a.shape == (3, 1, 6 + 10) and torch.equal(a, b)
        assert not torch.equal(gen(True, 8), b)                     # another seed, another sample
        greedy = gen(True, 1, do_sample=False)
        assert torch.equal(gen(True, 5, top_k=1), greedy) and torch.equal(gen(False, 6, top_k=1), greedy)
        ends = [int(b[0, 0, 8])]
        assert torch.equal(gen(False, 7, end_ids=ends), gen(True, 7, end_ids=ends))


def test_decode_plan_is_reused_between_generations_and_dropped_when_parameters_change(monkeypatch):
    """generation._DecodePlan: the captured step, its KV buffers, mask bias, counters and output buffer are kept on the
    model; the next generate() with the same shapes / options re-initialises them in place (the prefill lands in the
    plan's buffers through ops.KV_PREALLOC) and replays. Different prompts of the same shape must still give the loop's
    ids; a parameter update, another shape or other options force a new plan; results never alias the plan's buffer."""
    from cleantransformer_b200 import generation
    from cleantransformer_b200.models import modeling_bloom as mb
    torch.manual_seed(9)
    V = 71
    with mock_ops.patched():
        m = mb.BloomForCausalLM(mb.BloomConfig(vocab_size=V, hidden_size=64, n_layer=2, num_attention_heads=2)).eval()
        m._tie_weight()
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() >= 2:
                    p.normal_(0, 0.2)
        cfg = {"beam_size": 1, "do_sample": False, "max_gen_len": 6, "end_ids": None, "pad_id": 1}

        def prompts(seed, P=7):
            g = torch.Generator().manual_seed(seed)
            ids = torch.randint(3, V, (3, P), generator=g)
            mask = torch.ones(3, P, dtype=torch.long)
            n = int(torch.randint(0, P - 2, (1,), generator=g))
            mask[1, :n] = 0
            ids[1, :n] = 0
            return ids, mask

        def both(ids, mask, **extra):
            c = dict(cfg, **extra)
            monkeypatch.setenv("CT_DECODE_GRAPH", "0")
            want = m.generate(ids, attention_mask=mask, generation_configs=c)
            monkeypatch.setenv("CT_DECODE_GRAPH", "1")
            got = m.generate(ids, attention_mask=mask, generation_configs=c)
            assert torch.equal(want, got)
            return got

        first = both(*prompts(1))
        assert m._ct_decode_plan_reused is False and m._ct_decode_plan is not None
        keep = first.clone()
        second = both(*prompts(2))                       # same shapes, other tokens and padding: the plan is reused
        assert m._ct_decode_plan_reused is True and not torch.equal(second, first)
        assert torch.equal(first, keep), "a returned result must not alias the plan's output buffer"
        both(*prompts(3), end_ids=[int(second[0, 0, 9])])
        assert m._ct_decode_plan_reused is False         # other options (an end id): new plan
        both(*prompts(4, P=9))
        assert m._ct_decode_plan_reused is False         # other prompt length
        both(*prompts(5, P=9))
        assert m._ct_decode_plan_reused is True
        with torch.no_grad():
            m.bloom.blocks[0].mlp.dense_h_to_4h.weight.mul_(1.5)   # a parameter update invalidates the cached step
        both(*prompts(5, P=9))
        assert m._ct_decode_plan_reused is False
        monkeypatch.setattr(generation, "DECODE_PLAN_CACHE", [False])
        both(*prompts(6, P=9))
        both(*prompts(7, P=9))
        assert m._ct_decode_plan_reused is False


def _toy_trainer(tmp_path, max_steps, **extra):
    import types
    from cleantransformer_b200.models import modeling_bloom as mb
    from cleantransformer_b200.trainer import Trainer
    cfg = dict(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4)
    g = torch.Generator().manual_seed(3)
    data = [dict(input_ids=torch.randint(3, 64, (10,), generator=g), attention_mask=torch.ones(10, dtype=torch.long))
            for _ in range(12)]
    for d in data:
        d["labels"] = d["input_ids"].clone()

    def collate(items):
        return {k: torch.stack([it[k] for it in items]) for k in items[0]}

    torch.manual_seed(4)
    m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
    m._tie_weight()
    a = dict(per_device_train_batch_size=4, learning_rate=5e-3, max_steps=max_steps, logging_steps=1,
             output_dir=str(tmp_path), weight_decay=0.01, save_steps=5)
    a.update(extra)
    return Trainer(model=m, args=types.SimpleNamespace(**a), data_collator=collate, train_dataset=data), m


def test_trainer_checkpoints_are_reference_shaped_and_a_resumed_run_is_bit_identical(golden, tmp_path):
    """SURVEY §8 N4 (trainer/trainer.py:1303-1342 save, :351-379 + :448-453 resume, :1465-1486 rotation): the folders
    hold what the reference's loaders expect (plain torch.save'd state_dicts that torch.optim.AdamW itself accepts),
    are written behind the step loop, and a run resumed from the middle of an epoch ends with the SAME bits —
    parameters, both AdamW moments, step counters, logged losses — as the uninterrupted one."""
    import json, os
    with mock_ops.patched():
        straight, m_a = _toy_trainer(tmp_path / "a", 12)
        out = straight.train()
        assert out.global_step == 12
        folders = sorted(os.listdir(tmp_path / "a"))
        assert folders == ["checkpoint-10", "checkpoint-5"], folders
        for f in folders:
            assert sorted(os.listdir(tmp_path / "a" / f)) == ["optimizer.pt", "pytorch_model.bin", "rng_state.pth",
                                                                "trainer_state.json"]
        st = json.load(open(tmp_path / "a" / "checkpoint-5" / "trainer_state.json"))
        assert st["global_step"] == 5 and len(st["log_history"]) == 5 and abs(st["epoch"] - 5 / 3) < 1e-9

        # the files load into the reference's own stack: strict state_dict load, torch.optim.AdamW.load_state_dict
        sd = torch.load(str(tmp_path / "a" / "checkpoint-5" / "pytorch_model.bin"))
        assert set(sd) == set(m_a.state_dict())
        assert sd["lm_head.weight"].data_ptr() == sd["bloom.word_embeddings.weight"].data_ptr()   # still tied
        osd = torch.load(str(tmp_path / "a" / "checkpoint-5" / "optimizer.pt"))
        ref_opt = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in m_a.parameters()], lr=1.0)
        ref_opt.load_state_dict(osd)
        assert all(float(s["step"]) == 5.0 for s in ref_opt.state.values())

        # interrupted at step 7 (in the middle of epoch 2: its last checkpoint is step 5), then resumed to 12
        first, _ = _toy_trainer(tmp_path / "b", 7)
        first.train()
        assert sorted(os.listdir(tmp_path / "b")) == ["checkpoint-5"]
        second, m_b = _toy_trainer(tmp_path / "b", 12)
        torch.manual_seed(12345)                 # whatever state the new process has must not matter
        out_b = second.train(resume_from_checkpoint=True)
        assert out_b.global_step == 12
        assert [h["step"] for h in second.state.log_history] == list(range(1, 13))
        for ha, hb in zip(straight.state.log_history, second.state.log_history):
            assert ha == hb, (ha, hb)
        for (n, pa), (_, pb) in zip(m_a.named_parameters(), m_b.named_parameters()):
            assert torch.equal(pa, pb), n
        for pa, pb in zip(m_a.parameters(), m_b.parameters()):
            sa, sb = straight.optimizer.state[pa], second.optimizer.state[pb]
            assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])
            assert float(sa["step"]) == float(sb["step"]) == 12.0
            assert sb["exp_avg"].data_ptr() == second.optimizer._arena.param_view(pb, second.optimizer._arena.exp_avg).data_ptr()

        # rotation keeps the newest `save_total_limit` complete folders; "epoch" strategy saves at epoch ends
        rot, _ = _toy_trainer(tmp_path / "c", 9, save_steps=2, save_total_limit=2)
        rot.train()
        assert sorted(os.listdir(tmp_path / "c")) == ["checkpoint-6", "checkpoint-8"]
        ep, _ = _toy_trainer(tmp_path / "d", 7, save_strategy="epoch", save_only_model=True)
        ep.train()
        assert sorted(os.listdir(tmp_path / "d")) == ["checkpoint-3", "checkpoint-6"]
        assert sorted(os.listdir(tmp_path / "d" / "checkpoint-6")) == ["pytorch_model.bin", "trainer_state.json"]
        with pytest.raises(ValueError):
            _toy_trainer(tmp_path / "e", 3)[0].train(resume_from_checkpoint=True)


def test_fused_embedding_layernorm_node_host_logic(golden, monkeypatch):
    """SURVEY §8 N3 (functional.EmbeddingLNFn, CT_FUSED_EMBED_LN): the gather(s) + LayerNorm as one node — Bloom's
    word_embeddings -> word_embeddings_layernorm (modeling_bloom.py:190-191, tied table: the scatter is the SECOND
    write of its gradient) and BERT's three tables -> embedding_post (modeling_bert.py:297-301, padding_idx 0 gets no
    gradient) — against the reference's golden gradients / the oracle's autograd, with the fused call counted."""
    from cleantransformer_b200 import functional as F, ops
    from cleantransformer_b200.models import modeling_bert as mbert
    from oracle import ct_oracle as O
    g = golden("bloom_tiny")
    with mock_ops.patched():
        calls = []
        inner = ops.embedding_layernorm_fwd
        monkeypatch.setattr(ops, "embedding_layernorm_fwd", lambda *a, **k: (calls.append(1), inner(*a, **k))[1])
        monkeypatch.setattr(ops, "embedding_layernorm_ok", lambda weights, gamma: True)
        monkeypatch.setattr(F, "FUSED_EMBED_LN", True)
        m = _bloom(g)
        (loss, logits, hidden), kv = m(input_ids=g["ids"], attention_mask=g["mask"], labels=g["labels"])
        loss.backward()
        assert len(calls) == 1
        _check_against_golden(m, g, loss, logits, hidden)
        with torch.no_grad():                                   # nothing saved, nothing required: same values
            (lg2, _), _ = m(input_ids=g["ids"], attention_mask=g["mask"])
        assert torch.equal(lg2, logits.detach()) and len(calls) == 2

        gb = golden("bert_tiny")
        cfg = dict(gb["cfg"])
        ids, mask, seg, pos = gb["ids"], gb["mask"], gb["seg"], gb["pos"]
        torch.manual_seed(5)
        labels = torch.randint(0, cfg["num_labels"], (ids.shape[0],))
        sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in gb["sd"].items()}
        lg_ref, _, _ = O.bert_classifier(ids, mask, seg, pos, sd, cfg["num_hidden_layers"], cfg["num_attention_heads"],
                                         cfg["layer_norm_eps"])
        torch.nn.functional.cross_entropy(lg_ref, labels).backward()
        model = mbert.BertForSequenceClassification(mbert.BertConfig(**cfg)).eval()
        model.load_state_dict(gb["sd"], strict=True)
        out = model(ids, mask, seg, pos)
        torch.nn.functional.cross_entropy(out.float(), labels).backward()
        assert len(calls) == 3 and rel_err(out, lg_ref) < 2e-4
        for name in ("bert.word_embeddings.weight", "bert.position_embeddings.weight", "bert.segment_embeddings.weight",
                     "bert.embedding_post.0.weight", "bert.embedding_post.0.bias"):
            p, ref = dict(model.named_parameters())[name], sd[name].grad
            assert float((p.grad - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-9, name
        assert float(model.bert.word_embeddings.weight.grad[0].abs().max()) == 0.0 or not bool((ids == 0).any())


def test_trainer_gradient_accumulation_clipping_and_metrics(golden, tmp_path):
    """Trainer arguments that change the arithmetic are honoured, not ignored: `gradient_accumulation_steps` (two
    half-batches per optimizer step = one full batch: same trajectory up to rounding; a resume lands on the right
    micro-batch), `max_grad_norm` (trainer.py:486-493: the clipped run differs and its update is bounded),
    `compute_metrics` + `preprocess_logits_for_metrics` in evaluate() (trainer.py:621-739)."""
    import warnings
    with mock_ops.patched():
        full, m_full = _toy_trainer(tmp_path / "full", 6, save_strategy="no")
        full.train()
        acc, m_acc = _toy_trainer(tmp_path / "acc", 6, per_device_train_batch_size=2, gradient_accumulation_steps=2,
                                  save_steps=4)
        out = acc.train()
        assert out.global_step == 6 and abs(acc.state.epoch - 2.0) < 1e-9        # 12 samples / (2 x 2) = 3 steps per epoch
        for (n, a), (_, b) in zip(m_full.named_parameters(), m_acc.named_parameters()):
            if n.endswith("query_key_value.bias"):
                continue    # its key third has an analytically zero gradient: AdamW turns the rounding noise into +-lr
            assert rel_err(b, a) < 1e-4, n   # AdamW normalises the update: fp32 rounding of the half-batch sums shows
        for ha, hb in zip(full.state.log_history, acc.state.log_history):
            assert abs(ha["loss"] - hb["loss"]) < 1e-5 * abs(ha["loss"])
        res, m_res = _toy_trainer(tmp_path / "acc", 6, per_device_train_batch_size=2, gradient_accumulation_steps=2,
                                  save_strategy="no")
        res.train(resume_from_checkpoint=True)                                    # from step 4: epoch 1, one step in
        for (n, a), (_, b) in zip(m_acc.named_parameters(), m_res.named_parameters()):
            assert torch.equal(a, b), n

        clipped, m_clip = _toy_trainer(tmp_path / "clip", 1, save_strategy="no", max_grad_norm=1e-3, learning_rate=1.0,
                                       weight_decay=0.0)
        free, m_free = _toy_trainer(tmp_path / "free", 1, save_strategy="no", learning_rate=1.0, weight_decay=0.0)
        before = [p.detach().clone() for p in m_clip.parameters()]
        clipped.train(); free.train()
        gnorm = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in m_clip.parameters()))
        assert float(gnorm) <= 1e-3 * 1.001 and any(not torch.equal(a, b) for a, b in zip(m_clip.parameters(), m_free.parameters()))
        assert all(torch.isfinite(p).all() for p in m_clip.parameters()) and len(before) > 0

        seen = {}

        def metrics(pred):
            predictions, label_ids = pred
            seen["shapes"] = (predictions.shape, label_ids.shape)
            return {"top1": float((predictions[:, :-1] == label_ids[:, 1:]).mean()), "eval_already": 1.0}

        ev, _ = _toy_trainer(tmp_path / "ev", 1, save_strategy="no")
        ev.eval_dataset, ev.compute_metrics = ev.train_dataset, metrics
        ev.preprocess_logits_for_metrics = lambda logits, labels: logits.argmax(-1)
        m = ev.evaluate()
        assert seen["shapes"] == ((12, 10), (12, 10)) and set(m) == {"eval_top1", "eval_already", "eval_loss", "eval_batches"}
        assert 0.0 <= m["eval_top1"] <= 1.0 and m["eval_batches"] == 2           # eval batch size 8 (default)
        m2 = ev.evaluate({"a": ev.train_dataset[:4], "b": ev.train_dataset[4:]})
        assert {"eval_a_loss", "eval_b_top1"} <= set(m2)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            _toy_trainer(tmp_path / "cb", 1)[0].__class__(model=m_full, callbacks=[object()])
        assert any("callbacks" in str(x.message) for x in w)


def _trainer_ddp_worker(rank, world, port, out, root):
    import os
    import types
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cleantransformer_b200.ddp import DistributedDataParallel
    from cleantransformer_b200.models import modeling_bloom as mb
    from cleantransformer_b200.trainer import Trainer
    cfg = dict(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4)
    g = torch.Generator().manual_seed(3)
    data = [dict(input_ids=torch.randint(3, 64, (10,), generator=g), attention_mask=torch.ones(10, dtype=torch.long))
            for _ in range(16)]
    for d in data:
        d["labels"] = d["input_ids"].clone()

    def collate(items):
        return {k: torch.stack([it[k] for it in items]) for k in items[0]}

    def make(max_steps):
        torch.manual_seed(4 + rank)          # different initial weights per rank: the wrapper must broadcast rank 0's
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m._tie_weight()
        ddp = DistributedDataParallel(m, bucket_cap_mb=0.005)
        a = types.SimpleNamespace(per_device_train_batch_size=2, gradient_accumulation_steps=2, learning_rate=5e-3,
                                  max_steps=max_steps, logging_steps=1, output_dir=root, save_steps=3, weight_decay=0.0)
        return Trainer(model=ddp, args=a, data_collator=collate, train_dataset=data), m

    with mock_ops.patched():
        tr, m = make(5)
        tr.train()                           # 16 samples / (2 ranks x 2 x 2) = 2 optimizer steps per epoch
        dist.barrier()
        first = [p.detach().clone() for p in m.parameters()]
        files = sorted(os.listdir(root))
        tr2, m2 = make(5)
        tr2.train(resume_from_checkpoint=True)   # checkpoint-3 (written by rank 0 only), middle of epoch 1
        second = [p.detach().clone() for p in m2.parameters()]
    out[rank] = (first, second, files, tr.state.global_step, [h["loss"] for h in tr.state.log_history])
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_drives_the_ddp_wrapper_world2_gloo(tmp_path):
    """Trainer + DistributedDataParallel + TorchAdamW, two ranks (ft_bloom_DDP.py:99-156 through the Trainer surface):
    DistributedSampler shards the data, micro-batches accumulate under no_sync(), only rank 0 writes checkpoints (with
    un-prefixed keys: `model.module.state_dict()`), both ranks resume from them and end with the bits of the
    uninterrupted run, and the replicas stay identical throughout."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_trainer_ddp_worker, args=(2, port, out, str(tmp_path)), nprocs=2, join=True)
    (a0, b0, files0, steps0, loss0), (a1, b1, files1, steps1, loss1) = out[0], out[1]
    assert steps0 == steps1 == 5 and files0 == files1 == ["checkpoint-3"]
    assert loss0 != loss1                                   # each rank saw its own shard
    for x, y in zip(a0, a1):
        assert torch.equal(x, y)                            # replicas identical after training
    for x, y in zip(a0, b0):
        assert torch.equal(x, y)                            # resumed == uninterrupted, rank 0
    for x, y in zip(a1, b1):
        assert torch.equal(x, y)                            # ... and rank 1
    sd = torch.load(str(tmp_path / "checkpoint-3" / "pytorch_model.bin"))
    assert "lm_head.weight" in sd and not any(k.startswith("module.") for k in sd)


def test_beam_search_and_no_repeat_ngram_through_the_gpt_and_bloom_mirrors(golden):
    """examples/inference_gpt2.py:64-69 (beam_size 3, no_repeat_ngram_size 2) and inference_bloom.py:88-94 through the
    model mirrors and their KV caches: the reference's REAL GPT / Bloom (same weights) produced the expected ids
    (tests/golden/generation_beam.pt). Beam search re-orders the [b, h, t, d] caches by `index_select`; the next step
    appends to the re-ordered caches."""
    from cleantransformer_b200.models import modeling_bloom as mb, modeling_gpt as mg
    g = golden("generation_beam")
    with mock_ops.patched():
        gt = golden("gpt_tiny")
        cfg = dict(gt["cfg"])
        gpt = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version="gpt2").eval()
        gpt.load_state_dict(gt["gpt2"]["sd"], strict=True)
        gpt._tie_weights()
        bloom = _bloom(golden("bloom_tiny")).eval()
        for name, model in (("gpt2", gpt), ("bloom", bloom)):
            m = g["models"][name]
            for case, ref in zip(g["real_cases"], m["outputs"]):
                out = model.generate(m["ids"].clone(), attention_mask=m["mask"].clone(), generation_configs=dict(case))
                assert out.shape == ref.shape and torch.equal(out, ref), (name, case)
