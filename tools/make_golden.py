"""Generate tests/golden/*.pt by importing and running the REAL reference (read-only at
/root/reference) in the build container. The reference is pure Python and cannot travel to the GPU
box, so its live outputs on small seeded cases are committed as fixtures; tests/test_oracle_golden.py
pins oracle/ct_oracle.py against them, and the GPU parity tests then compare the CUDA path with the
oracle.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

Cases mirror the reference's own self-checks where it has them (transformer.py:134-156 seed 999,
optimizer.py:100-132) and SURVEY.md §8 d2 shapes scaled down.
"""
import os
import sys
import types

sys.dont_write_bytecode = True
REF = os.environ.get("CT_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)

# modeling_bert imports CleanTransformer.tokenizers which needs `toolz` (not installed):
# 2-function stand-in, only so the module imports (tokenizers are not on the hot path).
if "toolz" not in sys.modules:
    import itertools

    tz = types.ModuleType("toolz")
    tz.concat = lambda seqs: itertools.chain.from_iterable(seqs)

    def sliding_window(n, seq):
        seq = list(seq)
        return [tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)]

    tz.sliding_window = sliding_window
    sys.modules["toolz"] = tz

import torch  # noqa: E402

from CleanTransformer import transformer as rt  # noqa: E402
from CleanTransformer import optimizer as ropt  # noqa: E402
from CleanTransformer.models import modeling_bloom as rbloom  # noqa: E402
from CleanTransformer.models import modeling_gpt as rgpt  # noqa: E402
from CleanTransformer.models import modeling_bert as rbert  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def reinit(model, seed=999, std=0.02):
    """SURVEY.md §8 d2 weight init: matrices ~ N(0, 0.02), biases 0, LayerNorm w=1 b=0."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))  # LN gamma: not all-ones
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))        # biases / LN beta: non-zero
    return model


def sd_of(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def golden_layernorm():
    torch.manual_seed(999)  # transformer.py:134-141 (layernorm_sample)
    x = torch.rand((3, 4, 6))
    ln = rt.LayerNorm([4, 6])
    y = ln(x)
    torch.manual_seed(5)
    x2 = torch.randn(2, 5, 128)
    ln2 = rt.LayerNorm(128, eps=1e-12)
    with torch.no_grad():
        ln2.weight.copy_(torch.randn(128)); ln2.bias.copy_(torch.randn(128))
    x2r = x2.clone().requires_grad_(True)
    y2 = ln2(x2r)
    dy2 = torch.randn_like(y2)
    y2.backward(dy2)
    save("layernorm", {"x": x, "y": y.detach(), "x2": x2, "w2": ln2.weight.detach().clone(),
                       "b2": ln2.bias.detach().clone(), "eps2": 1e-12, "y2": y2.detach(), "dy2": dy2,
                       "dx2": x2r.grad.clone(), "dw2": ln2.weight.grad.clone(), "db2": ln2.bias.grad.clone()})


def golden_generic_block():
    torch.manual_seed(999)  # transformer.py:144-151 (t_TransformerBlock)
    cfg = rt.ExampleConfig()
    blk = rt.TransformerBlock(cfg).eval()
    q = torch.rand((3, 4, cfg.hidden_size))
    r = blk(q)
    # attention with an additive mask (BERT-style (1-m)*-1e4, modeling_bert.py:303-304)
    att = blk.attention
    m = torch.tensor([[1, 1, 1, 0], [1, 1, 0, 0], [1, 1, 1, 1]], dtype=torch.float32)
    add = (1.0 - m[:, None, None, :]) * -10000.0
    a = att(q, add)
    save("generic_block", {"x": q, "sd": sd_of(blk), "y": r.detach(), "mask": m, "att_masked": a.detach(),
                           "n_head": cfg.num_attention_heads, "eps": cfg.layer_norm_epsilong})


def golden_gelu():
    torch.manual_seed(3)
    x = torch.randn(4, 33) * 2
    g = torch.randn(4, 33)
    save("gelu", {"x": x, "g": g, "bloom_fwd": rbloom.bloom_gelu_forward(x),
                  "bloom_back": rbloom.bloom_gelu_back(g, (x,)),
                  "gelu_new": rgpt.NewGELUActivation()(x)})


def golden_bloom():
    cfg = rbloom.BloomConfig(vocab_size=97, hidden_size=64, n_layer=2, num_attention_heads=8)
    torch.manual_seed(999)
    model = reinit(rbloom.BloomForCausalLM(cfg))
    model._tie_weight()
    g = torch.Generator().manual_seed(1000)
    B, S = 3, 12
    ids = torch.randint(3, 97, (B, S), generator=g)
    lens = [12, 7, 9]
    mask = torch.zeros(B, S, dtype=torch.long)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
        ids[b, n:] = 3  # right padding with pad id 3 (ft_bloom.py:125)
    labels = ids.clone()  # ft_bloom.py:52: labels include pads
    model.train()  # GeLUFunction path (modeling_bloom.py:301-302); dropouts are p=0
    model.zero_grad()
    (loss, logits, hidden), kv = model(input_ids=ids, attention_mask=mask, labels=labels)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    # eval + cache: prefill on the first 8 tokens then one decode step (left-over rows all valid)
    model.eval()
    with torch.no_grad():
        full_mask = torch.ones(B, 9, dtype=torch.long)
        (lg_full, _), _ = model(input_ids=ids[:, :9], attention_mask=full_mask)
        (lg_pre, _), kv = model(input_ids=ids[:, :8], attention_mask=full_mask[:, :8])
        (lg_dec, _), kv2 = model(input_ids=ids[:, 8:9], attention_mask=full_mask, k_v_pasts=kv)
    save("bloom_tiny", {
        "cfg": {"vocab_size": 97, "hidden_size": 64, "n_layer": 2, "num_attention_heads": 8,
                "layer_norm_epsilon": cfg.layer_norm_epsilon},
        "sd": sd_of(model), "ids": ids, "mask": mask, "labels": labels,
        "loss": loss.detach(), "logits": logits.detach(), "hidden": hidden.detach(), "grads": grads,
        "alibi": rbloom.build_alibi_tensor(mask, 8, torch.float32),
        "logits_full9": lg_full, "logits_prefill8": lg_pre, "logits_decode": lg_dec,
        "kv_shape": list(kv2[0][0].shape),
    })


def golden_gpt():
    out = {}
    for version in ("gpt2", "gpt"):
        cfg = rgpt.GPTConfig(vocab_size=101, n_embd=48, n_positions=64, n_layer=2, n_head=4, n_ctx=64,
                             afn="gelu_new")
        torch.manual_seed(999)
        model = reinit(rgpt.GPTLMHeadModel(cfg, version=version)).eval()
        model._tie_weights()
        g = torch.Generator().manual_seed(999)
        B, P = 3, 8
        ids = torch.randint(1, 101, (B, P), generator=g)
        lens = [8, 5, 6]
        mask = torch.zeros(B, P, dtype=torch.long)
        for b, n in enumerate(lens):  # LEFT padding with 0 (inference_gpt2.py:55,59)
            mask[b, P - n:] = 1
            ids[b, :P - n] = 0
        with torch.no_grad():
            (logits, hidden), kv = model(ids, attention_mask=mask)
            gen = model.generate(ids, attention_mask=mask,
                                 generation_configs={"beam_size": 1, "do_sample": False, "max_gen_len": 6,
                                                     "end_ids": None, "pad_id": 0, "no_repeat_ngram_size": 0})
        # block-level fwd/bwd (config 1 shape scaled down): eval because of Dropout(0.5) gpt:136
        blk = model.gpt.blocks[0]
        torch.manual_seed(0)
        x = torch.randn(2, 16, 48, requires_grad=True)
        y, kvb = blk(x)
        torch.manual_seed(1)
        dy = torch.randn_like(y)
        model.zero_grad()
        y.backward(dy)
        bgr = {k: p.grad.detach().clone() for k, p in blk.named_parameters()}
        out[version] = {"sd": sd_of(model), "ids": ids, "mask": mask, "logits": logits, "hidden": hidden,
                        "generated": gen, "blk_x": x.detach().clone(), "blk_y": y.detach(), "blk_dy": dy,
                        "blk_dx": x.grad.clone(), "blk_grads": bgr,
                        "blk_k": kvb[0].detach(), "blk_v": kvb[1].detach()}
    out["cfg"] = {"vocab_size": 101, "n_embd": 48, "n_positions": 64, "n_layer": 2, "n_head": 4, "n_ctx": 64,
                  "layer_norm_epsilon": 1e-5, "afn": "gelu_new"}
    save("gpt_tiny", out)


def golden_bert():
    cfg = rbert.BertConfig(vocab_size=120, hidden_size=48, num_hidden_layers=2, num_attention_heads=4,
                           intermediate_size=96, max_position_embeddings=32, num_labels=5)
    torch.manual_seed(999)
    model = reinit(rbert.BertForSequenceClassification(cfg)).eval()
    g = torch.Generator().manual_seed(999)
    B, S = 3, 10
    ids = torch.randint(1, 120, (B, S), generator=g)
    lens = [10, 6, 8]
    mask = torch.zeros(B, S)
    for b, n in enumerate(lens):
        mask[b, :n] = 1.0
        ids[b, n:] = 0
    seg = torch.zeros(B, S, dtype=torch.long)
    pos = torch.arange(S)
    with torch.no_grad():
        logits = model(ids, mask, seg, pos)
        hidden, pooled = model.bert(ids, mask, seg, pos)
    save("bert_tiny", {"cfg": {"vocab_size": 120, "hidden_size": 48, "num_hidden_layers": 2,
                               "num_attention_heads": 4, "intermediate_size": 96,
                               "max_position_embeddings": 32, "num_labels": 5, "layer_norm_eps": cfg.layer_norm_eps},
                       "sd": sd_of(model), "ids": ids, "mask": mask, "seg": seg, "pos": pos,
                       "logits": logits, "hidden": hidden, "pooled": pooled})


def golden_optim():
    torch.manual_seed(999)
    p0 = [torch.rand(3, 4), torch.rand(4), torch.rand(37)]
    grads = [[torch.randn_like(p) for p in p0] for _ in range(4)]

    def run(make_opt):
        ps = [p.clone().requires_grad_(True) for p in p0]
        opt = make_opt(ps)
        traj = []
        for step_g in grads:
            for p, g in zip(ps, step_g):
                p.grad = g.clone()
            opt.step()
            traj.append([p.detach().clone() for p in ps])
        return traj, opt

    # the reference's own AdamW (optimizer.py:53-97) — must be given a LIST
    ref_adam, o = run(lambda ps: ropt.AdamW(ps, lr=0.01, weight_decay=0.01))
    ref_adam_m = [m.clone() for m in o.momentum_buffer]
    ref_adam_v = [v.clone() for v in o.rmsp_buffer]
    ref_adam_nowd, _ = run(lambda ps: ropt.AdamW(ps, lr=0.01))
    ref_sgd, _ = run(lambda ps: ropt.SGD(ps, lr=0.01, weight_decay=0.01, momentum=0.9))
    ref_sgd_plain, _ = run(lambda ps: ropt.SGD(ps, lr=0.01))
    torch_adamw, _ = run(lambda ps: torch.optim.AdamW(ps, lr=0.01, weight_decay=0.01))
    torch_sgd, _ = run(lambda ps: torch.optim.SGD(ps, lr=0.01, weight_decay=0.01, momentum=0.9))
    save("optim", {"p0": p0, "grads": grads, "ref_adamw": ref_adam, "ref_adamw_m": ref_adam_m,
                   "ref_adamw_v": ref_adam_v, "ref_adamw_nowd": ref_adam_nowd, "ref_sgd": ref_sgd,
                   "ref_sgd_plain": ref_sgd_plain, "torch_adamw": torch_adamw, "torch_sgd": torch_sgd})


def golden_sampling():
    """The reference's logits wrappers (generation/logits_processor.py:35-79) on seeded scores, including the
    clamped corners: temperature below 1e-2, top_k beyond the vocabulary, top_p at 0 and above 1."""
    from CleanTransformer.generation import logits_processor as lp
    torch.manual_seed(999)
    scores = torch.randn(4, 50) * 3
    scores[1, 7] = scores[1, 9]  # a tie
    out = {"scores": scores, "temperature": {}, "top_k": {}, "top_p": {}}
    for t in (0.7, 1.5, 0.001):
        out["temperature"][t] = lp.TemperatureLogitsWrapper(t)(None, scores.clone())
    for k in (1, 5, 1000):
        out["top_k"][k] = lp.TopKLogitsWrapper(k, min_tokens_to_keep=1)(None, scores.clone())
    for p_ in (0.8, 0.3, 0.0, 1.5):
        out["top_p"][p_] = lp.TopPLogitsWrapper(p_, min_tokens_to_keep=1)(None, scores.clone())
    save("sampling", out)


def golden_generation_loop():
    """The reference's own GenerationMixin (generation_util.py:13-119) around a deterministic toy model: greedy and
    sampled decoding (temperature / top-k / top-p wrappers, seeded torch.multinomial), EOS bookkeeping with one and
    several end ids, pad ids for finished rows, left-padded prompts."""
    from CleanTransformer.generation.generation_util import GenerationMixin

    class Cfg:
        n_layer = 2

    torch.manual_seed(0)
    emb = torch.randn(50, 50)

    class Toy(GenerationMixin):
        config = Cfg()

        def __call__(self, ids, attention_mask=None, k_v_pasts=None, **kw):
            n = attention_mask.sum(-1, keepdim=True).float()
            logits = emb[ids] + 0.01 * n[:, :, None]
            return (logits, logits), k_v_pasts

    ids = torch.tensor([[0, 0, 5, 7], [0, 3, 4, 9], [1, 2, 3, 4]])
    mask = torch.tensor([[0, 0, 1, 1], [0, 1, 1, 1], [1, 1, 1, 1]])
    cases = [dict(beam_size=1, do_sample=False, max_gen_len=6, end_ids=None, pad_id=0),
             dict(beam_size=1, do_sample=False, max_gen_len=6, end_ids=[13, 22], pad_id=0),
             dict(beam_size=1, do_sample=False, max_gen_len=3, end_ids=41, pad_id=2),
             dict(beam_size=1, do_sample=True, max_gen_len=5, end_ids=None, pad_id=0, temperature=0.7, top_k=5, top_p=0.9),
             dict(beam_size=1, do_sample=True, max_gen_len=5, end_ids=None, pad_id=0, temperature=1.0, top_k=0, top_p=0.5),
             dict(beam_size=1, do_sample=True, max_gen_len=5, end_ids=[7], pad_id=0)]
    outs = []
    for cfg in cases:
        torch.manual_seed(123)
        outs.append(Toy().generate(ids.clone(), attention_mask=mask.clone(), generation_configs=dict(cfg)))
    save("generation_loop", {"emb": emb, "ids": ids, "mask": mask, "cases": cases, "outputs": outs, "seed": 123})


def golden_generation_beam():
    """The rest of the reference's GenerationMixin.generate (generation_util.py:16-55): the no-repeat-ngram processor
    (logits_processor.py:11-32; what examples/inference_bloom.py:93 and inference_gpt2.py:68 configure) in the greedy /
    sampling loop, and beam search (generation_util.py:121-290; inference_gpt2.py:64 runs beam_size 3) — around a toy
    model whose logits depend on a per-row cache, so that the beam re-ordering of k_v_pasts matters."""
    from CleanTransformer.generation.generation_util import GenerationMixin
    from CleanTransformer.generation.logits_processor import NoRepeatNGramLogitsProcessor

    class Cfg:
        n_layer = 2

    torch.manual_seed(1)
    emb = torch.randn(40, 40)

    class Toy(GenerationMixin):
        config = Cfg()

        def __call__(self, ids, attention_mask=None, k_v_pasts=None, **kw):
            new = []
            for past in k_v_pasts:
                run = ids.sum(-1, keepdim=True).float() if past is None else past[0] + ids.sum(-1, keepdim=True).float()
                new.append((run, -run))
            n = attention_mask.sum(-1, keepdim=True).float()
            logits = emb[ids] + 0.01 * n[:, :, None] + 0.003 * torch.sin(new[0][0])[:, :, None] * emb[(ids + 1) % 40]
            return (logits, logits), new

    torch.manual_seed(5)
    proc = {"ids": torch.tensor([[3, 4, 3, 4, 3], [1, 1, 1, 1, 1], [5, 6, 7, 5, 6], [0, 0, 9, 8, 9]]),
            "scores": torch.randn(4, 12), "out": {}}
    for n in (2, 3, 6):
        proc["out"][n] = NoRepeatNGramLogitsProcessor(n)(proc["ids"], proc["scores"].clone())
    ids = torch.tensor([[0, 0, 5, 7], [0, 3, 4, 9], [1, 2, 3, 4]])
    mask = torch.tensor([[0, 0, 1, 1], [0, 1, 1, 1], [1, 1, 1, 1]])
    cases = [dict(beam_size=1, do_sample=False, max_gen_len=12, end_ids=None, pad_id=0, no_repeat_ngram_size=2),
             dict(beam_size=1, do_sample=False, max_gen_len=12, end_ids=[13], pad_id=0, no_repeat_ngram_size=3),
             dict(beam_size=1, do_sample=True, max_gen_len=8, end_ids=None, pad_id=0, no_repeat_ngram_size=2,
                  temperature=0.7, top_k=5, top_p=0.9),
             dict(beam_size=3, do_sample=False, max_gen_len=6, end_ids=[13, 22], pad_id=0),
             dict(beam_size=3, do_sample=False, max_gen_len=8, end_ids=[7, 11, 30], pad_id=2, early_stop=False),
             dict(beam_size=2, do_sample=False, max_gen_len=8, end_ids=[7, 11, 30, 31, 32, 33], pad_id=0,
                  no_repeat_ngram_size=2),
             dict(beam_size=4, do_sample=False, max_gen_len=10, end_ids=list(range(0, 40, 3)), pad_id=1),
             dict(beam_size=3, do_sample=True, max_gen_len=6, end_ids=[13, 22], pad_id=0, temperature=0.7, top_k=5,
                  top_p=0.9),
             dict(beam_size=2, do_sample=True, max_gen_len=6, end_ids=[5], pad_id=0, temperature=1.3, top_k=0, top_p=1.0,
                  no_repeat_ngram_size=2)]
    outs = []
    for cfg in cases:
        torch.manual_seed(321)
        outs.append(Toy().generate(ids.clone(), attention_mask=mask.clone(), generation_configs=dict(cfg)))
    pos = torch.tensor([[0, 0, 0, 1], [0, 0, 1, 2], [0, 1, 2, 3]])
    seg = torch.tensor([[0, 0, 1, 1], [0, 1, 1, 1], [1, 1, 1, 2]])

    class ToyPos(Toy):
        def __call__(self, ids, attention_mask=None, k_v_pasts=None, position_ids=None, segment_ids=None, **kw):
            (logits, _), new = Toy.__call__(self, ids, attention_mask=attention_mask, k_v_pasts=k_v_pasts)
            logits = logits + 0.02 * emb[(position_ids + 2 * segment_ids) % 40]
            return (logits, logits), new

    torch.manual_seed(321)
    with_pos = ToyPos().generate(ids.clone(), attention_mask=mask.clone(), position_ids=pos.clone(),
                                 segment_ids=seg.clone(), generation_configs=dict(cases[3]))
    # the same callers around the reference's REAL models (the tiny GPT / Bloom of the other fixtures): beam search
    # re-orders real [b, h, t, d] caches, the processor sees left-padded prompts
    models = {}
    gt = torch.load(os.path.join(OUT, "gpt_tiny.pt"), weights_only=False)
    gcfg = dict(gt["cfg"])
    gcfg.pop("layer_norm_epsilon", None)
    gpt = rgpt.GPTLMHeadModel(rgpt.GPTConfig(**gcfg), version="gpt2").eval()
    gpt.load_state_dict(gt["gpt2"]["sd"], strict=True)
    gpt._tie_weights()
    bt = torch.load(os.path.join(OUT, "bloom_tiny.pt"), weights_only=False)
    bloom = rbloom.BloomForCausalLM(rbloom.BloomConfig(**bt["cfg"])).eval()
    bloom.load_state_dict(bt["sd"], strict=True)
    bloom._tie_weight()
    real_cases = [dict(beam_size=3, do_sample=False, max_gen_len=8, end_ids=[5, 17, 40, 63, 88], pad_id=0,
                       no_repeat_ngram_size=2),
                  dict(beam_size=2, do_sample=False, max_gen_len=6, end_ids=list(range(0, 100, 4)), pad_id=0,
                       early_stop=False),
                  dict(beam_size=1, do_sample=False, max_gen_len=10, end_ids=None, pad_id=0, no_repeat_ngram_size=2)]
    with torch.no_grad():
        models["gpt2"] = {"ids": gt["gpt2"]["ids"], "mask": gt["gpt2"]["mask"], "outputs": [
            gpt.generate(gt["gpt2"]["ids"].clone(), attention_mask=gt["gpt2"]["mask"].clone(),
                         generation_configs=dict(c)) for c in real_cases]}
        bids, bmask = bt["ids"][:, :8].clone(), torch.ones_like(bt["ids"][:, :8])
        models["bloom"] = {"ids": bids, "mask": bmask, "outputs": [
            bloom.generate(bids.clone(), attention_mask=bmask.clone(), generation_configs=dict(c)) for c in real_cases]}
    save("generation_beam", {"emb": emb, "ids": ids, "mask": mask, "cases": cases, "outputs": outs, "seed": 321,
                             "processor": proc, "pos": pos, "seg": seg, "with_pos_case": 3, "with_pos": with_pos,
                             "real_cases": real_cases, "models": models})


if __name__ == "__main__":
    torch.set_num_threads(4)
    if len(sys.argv) > 1:  # regenerate only the named fixtures: python tools/make_golden.py sampling
        for name in sys.argv[1:]:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_layernorm()
    golden_generic_block()
    golden_gelu()
    golden_bloom()
    golden_gpt()
    golden_bert()
    golden_optim()
    golden_sampling()
    golden_generation_loop()
    golden_generation_beam()
