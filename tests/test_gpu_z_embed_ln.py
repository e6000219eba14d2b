"""GPU parity of the fused gather + LayerNorm kernel (SURVEY §8 f N3; `ct_embedding_layernorm_fwd`,
csrc/layernorm.cu) — modeling_bloom.py:190-191 (word_embeddings -> word_embeddings_layernorm) and
modeling_bert.py:297-301 (word + segment + position tables -> embedding_post LayerNorm) — against the oracle's
LayerNorm on a torch gather, against the two-kernel path it replaces, and through the Bloom / BERT mirrors with
`functional.FUSED_EMBED_LN` on and off. (File name: runs after the established GPU suite.)"""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.mark.parametrize("H,ntab,out_dtype", [(1024, 1, torch.float32), (768, 3, torch.float32), (128, 2, torch.bfloat16),
                                              (256, 3, torch.float16)])
def test_fused_gather_layernorm_vs_oracle_and_vs_the_two_kernel_path(H, ntab, out_dtype):
    from cleantransformer_b200 import ops
    from oracle import ct_oracle as O
    torch.manual_seed(H + ntab)
    B, S = 3, 257                                           # 771 rows: not a multiple of the 8 warps of a CTA
    vocabs = [5003, 2, 600][:ntab]
    tables = [torch.randn(v, H, device=DEV) * (0.02 if k == 0 else 1.0) for k, v in enumerate(vocabs)]
    ids = [torch.randint(0, v, (B, S), device=DEV) for v in vocabs]
    if ntab == 3:
        ids[2] = torch.arange(S, device=DEV)[None, :].expand(B, S).contiguous()   # position ids
    gamma = 1.0 + 0.1 * torch.randn(H, device=DEV)
    beta = 0.1 * torch.randn(H, device=DEV)
    eps = 1e-12 if ntab == 3 else 1e-5                      # BERT's eps (modeling_bert.py:30) / Bloom's
    assert ops.embedding_layernorm_ok(tables, gamma)
    emb, y, y2, mean, rstd = ops.embedding_layernorm_fwd(ids, tables, gamma, beta, eps, out_dtype, torch.bfloat16)
    # the reference arithmetic: torch gather(s), then the oracle's LayerNorm (transformer.py:79-89)
    want_emb = sum(t[i] for t, i in zip(tables, ids))
    want = O.layernorm(want_emb, gamma, beta, eps)
    assert torch.equal(emb, want_emb)                       # fp32 sums in the same order: exact
    tol = {torch.float32: 1e-5, torch.bfloat16: 4e-3, torch.float16: 1e-3}[out_dtype]
    assert y.dtype == out_dtype and y2.dtype == torch.bfloat16
    assert rel_err(y.float(), want) < tol and rel_err(y2.float(), want) < 4e-3
    # the two kernels it replaces
    e2 = None
    for i, t in zip(ids, tables):
        e2 = ops.embedding_fwd(i, t, e2, accumulate=e2 is not None)
    yb, y2b, mean_b, rstd_b = ops.layernorm_fwd(e2, gamma, beta, eps, out_dtype, torch.bfloat16)
    assert torch.equal(e2.view_as(emb), emb)
    assert rel_err(mean, mean_b) < 1e-6 and rel_err(rstd, rstd_b) < 1e-6
    assert rel_err(y.float(), yb.float()) < (1e-6 if out_dtype == torch.float32 else tol)
    # inference form: nothing saved
    emb0, y0, none2, m0, r0 = ops.embedding_layernorm_fwd(ids, tables, gamma, beta, eps, out_dtype, None, save=False)
    assert emb0 is None and none2 is None and m0 is None and r0 is None and torch.equal(y0, y)


def test_fused_gather_layernorm_edge_cases():
    """An id outside its table poisons exactly its row (ct_embedding_fwd's policy: torch would raise); empty input;
    shapes the register-resident kernel does not take are refused with the library's 'unsupported' code."""
    from cleantransformer_b200 import ops
    H = 256
    table = torch.randn(100, H, device=DEV)
    gamma, beta = torch.ones(H, device=DEV), torch.zeros(H, device=DEV)
    ids = torch.randint(0, 100, (4, 9), device=DEV)
    ids[2, 5] = 100
    ids[0, 0] = -1
    _, y, _, _, _ = ops.embedding_layernorm_fwd([ids], [table], gamma, beta, 1e-5)
    bad = torch.isnan(y).any(-1)
    assert bad[2, 5] and bad[0, 0] and int(bad.sum()) == 2
    _, y, _, _, _ = ops.embedding_layernorm_fwd([ids[:0]], [table], gamma, beta, 1e-5)
    assert y.shape == (0, 9, H)
    assert not ops.embedding_layernorm_ok([torch.randn(10, 96, device=DEV)], torch.ones(96, device=DEV))
    with pytest.raises(RuntimeError):
        ops.embedding_layernorm_fwd([ids], [torch.randn(100, 96, device=DEV)], torch.ones(96, device=DEV),
                                    torch.zeros(96, device=DEV), 1e-5)


def _grads(model):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def test_bloom_and_bert_mirrors_with_the_fused_preamble_match_the_two_kernel_preamble(monkeypatch):
    """Whole models, training direction: logits and EVERY gradient with FUSED_EMBED_LN on vs off (same kernels
    everywhere else; bounds: 1e-3 on the logits, 5e-3 of the tensor maximum on the gradients — a last-bit difference in
    the preamble may flip bf16 roundings downstream), incl. Bloom's tied table (second write of its gradient) and BERT's padding row."""
    from cleantransformer_b200 import functional as F, ops
    from cleantransformer_b200.models import modeling_bloom as mb, modeling_bert as mbert
    torch.manual_seed(11)
    with torch.device(DEV):
        bloom = mb.BloomForCausalLM(mb.BloomConfig(vocab_size=1024, hidden_size=256, n_layer=2, num_attention_heads=4))
        bert = mbert.BertForSequenceClassification(mbert.BertConfig(
            vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
            max_position_embeddings=160, num_labels=5, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    with torch.no_grad():
        for m in (bloom, bert):
            for p in m.parameters():
                if p.dim() >= 2:
                    p.normal_(0.0, 0.02)
    bloom._tie_weight()
    ids = torch.randint(3, 1024, (2, 160), device=DEV)
    mask = torch.ones(2, 160, dtype=torch.long, device=DEV)
    mask[1, 120:] = 0
    bids = torch.randint(1, 1000, (3, 160), device=DEV)
    bmask = torch.ones(3, 160, dtype=torch.long, device=DEV)
    bmask[0, 100:] = 0
    bids[0, 100:] = 0                                       # padding_idx rows
    seg = torch.zeros_like(bids)
    labels = torch.tensor([0, 3, 4], device=DEV)
    calls = []
    inner = ops.embedding_layernorm_fwd
    monkeypatch.setattr(ops, "embedding_layernorm_fwd", lambda *a, **k: (calls.append(1), inner(*a, **k))[1])

    def run(fused):
        monkeypatch.setattr(F, "FUSED_EMBED_LN", fused)
        out = {}
        bloom.train(); bert.train()
        for m in (bloom, bert):
            for p in m.parameters():
                p.grad = None
        (loss, logits, _), _ = bloom(input_ids=ids, attention_mask=mask, labels=ids)
        loss.backward()
        out["bloom"] = (logits.detach().float().clone(), _grads(bloom))
        lg = bert(bids, bmask, seg, None)
        torch.nn.functional.cross_entropy(lg.float(), labels).backward()
        out["bert"] = (lg.detach().float().clone(), _grads(bert))
        torch.cuda.synchronize()
        return out

    base = run(False)
    assert not calls
    fused = run(True)
    assert len(calls) == 2
    for name in ("bloom", "bert"):
        assert rel_err(fused[name][0], base[name][0]) < 1e-3, name
        assert set(fused[name][1]) == set(base[name][1])
        for k, g in base[name][1].items():
            # floor: analytically zero gradients (the key bias: softmax is shift invariant) are rounding noise
            err = float((fused[name][1][k] - g).abs().max() / g.abs().max().clamp_min(1e-6))
            assert err < 5e-3, (name, k, err)
    assert float(fused["bert"][1]["bert.word_embeddings.weight"][0].abs().max()) == 0.0
