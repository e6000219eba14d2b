"""Pin oracle/ct_oracle.py against the golden vectors generated from the REAL reference
(tools/make_golden.py, run in the build container where /root/reference is importable).
CPU only; fp32; tolerance 1e-5 relative (SURVEY.md §8 d3, fp32 paths) — most cases are bit-exact
because the oracle restates the same torch op sequence."""
import torch

from conftest import rel_err
from oracle import ct_oracle as O

TOL = 1e-5


def test_layernorm_matches_reference(golden):
    g = golden("layernorm")
    ln_w = torch.ones(4, 6); ln_b = torch.zeros(4, 6)
    assert rel_err(O.layernorm(g["x"], ln_w, ln_b, 1e-5), g["y"]) < TOL
    x = g["x2"].clone().requires_grad_(True)
    w = g["w2"].clone().requires_grad_(True); b = g["b2"].clone().requires_grad_(True)
    y = O.layernorm(x, w, b, g["eps2"])
    assert rel_err(y, g["y2"]) < TOL
    y.backward(g["dy2"])
    assert rel_err(x.grad, g["dx2"]) < TOL
    assert rel_err(w.grad, g["dw2"]) < TOL
    assert rel_err(b.grad, g["db2"]) < TOL
    # and the reference's own self-check: equals torch.nn.LayerNorm (transformer.py:134-141)
    ref = torch.nn.functional.layer_norm(g["x"], (4, 6), ln_w, ln_b, 1e-5)
    assert rel_err(O.layernorm(g["x"], ln_w, ln_b, 1e-5), ref) < TOL


def test_generic_block_matches_reference(golden):
    g = golden("generic_block")
    y = O.generic_block(g["x"], g["sd"], g["n_head"], g["eps"])
    assert rel_err(y, g["y"]) < TOL
    sd = g["sd"]
    add = (1.0 - g["mask"][:, None, None, :]) * -10000.0
    a = O.attention_generic(g["x"], sd["attention.q_linear.weight"], sd["attention.q_linear.bias"],
                            sd["attention.k_linear.weight"], sd["attention.k_linear.bias"],
                            sd["attention.v_linear.weight"], sd["attention.v_linear.bias"], g["n_head"], add)
    assert rel_err(a, g["att_masked"]) < TOL


def test_gelu_matches_reference(golden):
    g = golden("gelu")
    assert rel_err(O.gelu_tanh_bloom(g["x"]), g["bloom_fwd"]) < TOL
    assert rel_err(O.gelu_tanh_bloom_back(g["g"], g["x"]), g["bloom_back"]) < TOL
    assert rel_err(O.gelu_new(g["x"]), g["gelu_new"]) < TOL


def test_bloom_matches_reference(golden):
    g = golden("bloom_tiny")
    cfg = g["cfg"]
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["sd"].items() if k != "lm_head.weight"}
    (loss, logits, hidden), kv = O.bloom_causal_lm(g["ids"], g["mask"], sd, cfg["n_layer"],
                                                   cfg["num_attention_heads"], cfg["layer_norm_epsilon"],
                                                   labels=g["labels"], training=True)
    assert rel_err(loss, g["loss"]) < TOL
    assert rel_err(logits, g["logits"]) < TOL
    assert rel_err(hidden, g["hidden"]) < TOL
    loss.backward()
    for k, gr in g["grads"].items():
        if k == "lm_head.weight":
            continue  # tied: same tensor as bloom.word_embeddings.weight
        assert rel_err(sd[k].grad, gr) < 5e-5, k
    assert rel_err(O.build_alibi_tensor(g["mask"], 8, torch.float32), g["alibi"]) < TOL
    # k_v_cache path: prefill 8 + decode 1 == full 9
    sdn = {k: v for k, v in g["sd"].items()}
    ones = torch.ones(3, 9, dtype=torch.long)
    with torch.no_grad():
        (lp, _), kv = O.bloom_causal_lm(g["ids"][:, :8], ones[:, :8], sdn, 2, 8, 1e-5)
        (ld, _), kv2 = O.bloom_causal_lm(g["ids"][:, 8:9], ones, sdn, 2, 8, 1e-5, k_v_pasts=kv)
    assert rel_err(lp, g["logits_prefill8"]) < TOL
    assert rel_err(ld, g["logits_decode"]) < TOL
    assert list(kv2[0][0].shape) == g["kv_shape"]


def test_gpt_matches_reference(golden):
    g = golden("gpt_tiny")
    cfg = g["cfg"]
    for version in ("gpt2", "gpt"):
        c = g[version]
        with torch.no_grad():
            (logits, hidden), _ = O.gpt_lm_head_model(c["ids"], c["mask"], c["sd"], cfg["n_layer"], cfg["n_head"],
                                                      cfg["n_ctx"], cfg["layer_norm_epsilon"], version=version)
        assert rel_err(logits, c["logits"]) < TOL
        assert rel_err(hidden, c["hidden"]) < TOL

        def step(ids, mask, kv, c=c, version=version):
            return O.gpt_lm_head_model(ids, mask, c["sd"], cfg["n_layer"], cfg["n_head"], cfg["n_ctx"],
                                       cfg["layer_norm_epsilon"], version=version, k_v_pasts=kv)

        with torch.no_grad():
            gen = O.greedy_generate(step, c["ids"], c["mask"], cfg["n_layer"], max_gen_len=6, pad_id=0)
        assert torch.equal(gen, c["generated"])  # token ids bit-exact
        assert gen.shape[-1] == c["ids"].shape[1] + 6 + 2  # reference emits max_gen_len + 2
        # block forward/backward
        sd = {k[len("gpt.blocks.0."):]: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(".attn.bias"))
              for k, v in c["sd"].items() if k.startswith("gpt.blocks.0.")}
        x = c["blk_x"].clone().requires_grad_(True)
        y, (k_, v_) = O.gpt_block(x, sd, "", cfg["n_head"], cfg["n_ctx"], 1e-5, version=version)
        assert rel_err(y, c["blk_y"]) < TOL
        assert rel_err(k_, c["blk_k"]) < TOL and rel_err(v_, c["blk_v"]) < TOL
        y.backward(c["blk_dy"])
        assert rel_err(x.grad, c["blk_dx"]) < TOL
        for name, gr in c["blk_grads"].items():
            assert rel_err(sd[name].grad, gr) < 5e-5, name


def test_bert_matches_reference(golden):
    g = golden("bert_tiny")
    cfg = g["cfg"]
    with torch.no_grad():
        logits, hidden, pooled = O.bert_classifier(g["ids"], g["mask"], g["seg"], g["pos"], g["sd"],
                                                   cfg["num_hidden_layers"], cfg["num_attention_heads"],
                                                   cfg["layer_norm_eps"])
    assert rel_err(logits, g["logits"]) < TOL
    assert rel_err(hidden, g["hidden"]) < TOL
    assert rel_err(pooled, g["pooled"]) < TOL


def test_optimizers_match_reference(golden):
    g = golden("optim")
    n = len(g["p0"])

    def run(stepfn, has_state=True):
        ps = [p.clone() for p in g["p0"]]
        ms = [torch.zeros_like(p) for p in ps]; vs = [torch.zeros_like(p) for p in ps]
        traj = []
        for t, step_g in enumerate(g["grads"], start=1):
            for i in range(n):
                ps[i], _, ms[i], vs[i] = stepfn(ps[i], step_g[i].clone(), ms[i], vs[i], t)
            traj.append([p.clone() for p in ps])
        return traj, ms, vs

    traj, ms, vs = run(lambda p, gr, m, v, t: O.adamw_reference_step(p, gr, m, v, t, lr=0.01, weight_decay=0.01))
    for a, b in zip(traj, g["ref_adamw"]):
        for x, y in zip(a, b):
            assert rel_err(x, y) < TOL
    for x, y in zip(ms, g["ref_adamw_m"]):
        assert rel_err(x, y) < TOL
    for x, y in zip(vs, g["ref_adamw_v"]):
        assert rel_err(x, y) < TOL
    traj, _, _ = run(lambda p, gr, m, v, t: O.adamw_reference_step(p, gr, m, v, t, lr=0.01))
    for a, b in zip(traj, g["ref_adamw_nowd"]):
        for x, y in zip(a, b):
            assert rel_err(x, y) < TOL
    traj, _, _ = run(lambda p, gr, m, v, t: O.adamw_torch_step(p, gr, m, v, t, lr=0.01, weight_decay=0.01))
    for a, b in zip(traj, g["torch_adamw"]):
        for x, y in zip(a, b):
            assert rel_err(x, y) < TOL
    # SGD (optimizer.py:28-50); reference == torch.optim.SGD on this harness
    for key, kw in (("ref_sgd", dict(lr=0.01, momentum=0.9, weight_decay=0.01)), ("ref_sgd_plain", dict(lr=0.01))):
        ps = [p.clone() for p in g["p0"]]; bufs = [None] * n
        for t, step_g in enumerate(g["grads"]):
            for i in range(n):
                ps[i], _, bufs[i] = O.sgd_reference_step(ps[i], step_g[i].clone(), bufs[i], **kw)
            for x, y in zip(ps, g[key][t]):
                assert rel_err(x, y) < TOL
    for a, b in zip(g["ref_sgd"], g["torch_sgd"]):
        for x, y in zip(a, b):
            assert rel_err(x, y) < 1e-5


def test_sampling_wrappers_match_reference(golden):
    """generation/logits_processor.py:35-79 — oracle restatement AND the product's host-side filters
    (cleantransformer_b200/generation.py) against the real reference's wrappers, clamped corners included."""
    from cleantransformer_b200 import generation as G
    from oracle import ct_oracle as O
    g = golden("sampling")
    s = g["scores"]
    for t, ref in g["temperature"].items():
        assert torch.equal(O.temperature_wrapper(s.clone(), t), ref)
        assert torch.equal(G._temperature(s.clone(), t), ref)
    for k, ref in g["top_k"].items():
        assert torch.equal(O.top_k_wrapper(s.clone(), k), ref)
        assert torch.equal(G._filter_top_k(s.clone(), k), ref)
    for p, ref in g["top_p"].items():
        assert torch.equal(O.top_p_wrapper(s.clone(), p), ref)
        assert torch.equal(G._filter_top_p(s.clone(), p), ref)
    assert int(torch.isfinite(g["top_p"][0.0]).sum()) == s.shape[0]  # top_p = 0 keeps exactly the arg-max
