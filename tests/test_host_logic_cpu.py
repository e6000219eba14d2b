"""Host-side logic that needs no GPU: arena layout, bucket planning, the decode loop's control flow,
state_dict / class-surface parity with the reference, and the N>1 DDP wrapper over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err


def test_arena_layout_and_views():
    from cleantransformer_b200.arena import ParamArena, arena_of, ALIGN
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(130))]
    before = [p.detach().clone() for p in ps]
    a = ParamArena(ps + [ps[0]])  # duplicates (tied weights) are stored once
    assert len(a.params) == 3
    assert all(o % ALIGN == 0 for o in a.offsets)
    for p, b in zip(ps, before):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == a.flat[p._ct_off:].data_ptr()
        assert p._ct_grad_view.shape == p.shape
    assert arena_of(ps) is a
    assert arena_of(ps[:2]) is None
    a.flat.zero_()
    assert float(ps[2].abs().sum()) == 0.0  # parameters are views of the flat buffer
    with pytest.raises(ValueError):
        ParamArena(iter([]))


def test_bucket_plan_covers_arena_back_to_front():
    from cleantransformer_b200.ddp import plan_buckets
    so, off = [], 0
    for n in [64, 640, 64, 6400, 128, 64]:
        so.append((off, n)); off += n
    buckets = plan_buckets(so, cap_elems=700)
    covered = sorted(i for _, _, idxs in buckets for i in idxs)
    assert covered == list(range(len(so)))
    assert buckets[0][2][0] == len(so) - 1  # last parameter first (backward order)
    for lo, hi, idxs in buckets:
        assert lo == so[min(idxs)][0] and hi == so[max(idxs)][0] + so[max(idxs)][1]
    his = [b[1] for b in buckets]
    assert his == sorted(his, reverse=True)
    # parameters on their own reduction schedule (tied tables: ddp.py) never share a bucket
    for solo in ({0}, {3}, {5}, {1, 2}):
        bs = plan_buckets(so, cap_elems=700, solo=solo)
        assert sorted(i for _, _, idxs in bs for i in idxs) == list(range(len(so)))
        for lo, hi, idxs in bs:
            assert lo == so[min(idxs)][0] and hi == so[max(idxs)][0] + so[max(idxs)][1]
            assert len(idxs) == 1 or not (set(idxs) & solo)


def test_greedy_loop_matches_reference_semantics(golden):
    """Control flow of generation.py vs the oracle restatement of generation_util.py:57-119, with a
    toy deterministic 'model' so no kernel is needed."""
    from cleantransformer_b200.generation import GenerationMixin
    from oracle import ct_oracle as O

    class Cfg:
        n_layer = 2

    class Toy(GenerationMixin):
        config = Cfg()

        def __call__(self, ids, attention_mask=None, k_v_pasts=None, **kw):
            past = 0 if k_v_pasts[0] is None else k_v_pasts[0]
            total = past + ids.shape[1]
            assert attention_mask.shape[1] == total
            logits = torch.zeros(ids.shape[0], ids.shape[1], 11)
            nxt = (ids[:, -1] * 3 + total) % 11
            logits[torch.arange(ids.shape[0]), -1, nxt] = 1.0
            return (logits, None), [total, total]

    ids = torch.tensor([[0, 4, 5], [0, 0, 7]])
    mask = torch.tensor([[1, 1, 1], [0, 0, 1]])
    toy = Toy()
    out = toy.generate(ids, attention_mask=mask, generation_configs={"beam_size": 1, "do_sample": False, "max_gen_len": 5,
                                                                    "end_ids": None, "pad_id": 0})
    ref = O.greedy_generate(lambda i, m, kv: toy(i, attention_mask=m, k_v_pasts=kv), ids, mask, 2, 5, 0)
    assert torch.equal(out, ref)
    assert out.shape == (2, 1, 3 + 5 + 2)  # the reference emits max_gen_len + 2 tokens
    # EOS handling: finished rows emit pad_id
    out2 = toy.generate(ids, attention_mask=mask, generation_configs={"beam_size": 1, "do_sample": False, "max_gen_len": 5,
                                                                     "end_ids": int(out[0, 0, 3]), "pad_id": 0})
    assert int(out2[0, 0, 3]) == int(out[0, 0, 3]) and int(out2[0, 0, 4]) == 0
    with pytest.raises(TypeError):  # like the reference, beam search cannot run without end ids (generation_util.py:140)
        toy.generate(ids, attention_mask=mask, generation_configs={"beam_size": 4})


def test_class_surface_and_state_dict_keys_match_reference(golden):
    from cleantransformer_b200 import transformer as T, optimizer as Opt
    from cleantransformer_b200.models import modeling_bloom as mb, modeling_gpt as mg, modeling_bert as mbert
    g = golden("bloom_tiny")
    m = mb.BloomForCausalLM(mb.BloomConfig(**g["cfg"]))
    assert list(m.state_dict().keys()) == list(g["sd"].keys())
    m.load_state_dict(g["sd"], strict=True)
    m._tie_weight()
    assert m.lm_head.weight is m.bloom.word_embeddings.weight
    gg = golden("gpt_tiny")
    for v in ("gpt2", "gpt"):
        gm = mg.GPTLMHeadModel(mg.GPTConfig(**gg["cfg"]), version=v)
        assert list(gm.state_dict().keys()) == list(gg[v]["sd"].keys())
        assert gm.gpt.blocks[0].attn.c_attn.weight.shape == (48, 144)  # Conv1D stores [in, out]
    gb = golden("bert_tiny")
    bm = mbert.BertForSequenceClassification(mbert.BertConfig(**gb["cfg"]))
    assert list(bm.state_dict().keys()) == list(gb["sd"].keys())
    blk = T.TransformerBlock(T.ExampleConfig())
    assert list(blk.state_dict().keys()) == list(golden("generic_block")["sd"].keys())
    assert T.MultiHeadAttention is T.AttentionLayer
    o = Opt.AdamW(m.parameters())  # a generator is fine (the reference silently breaks on it)
    assert len(o.params) == len(list(m.parameters())) and o.steps[0] == 1
    assert hasattr(o, "momentum_buffer") and hasattr(o, "rmsp_buffer")
    assert torch.allclose(mb.build_alibi_tensor(g["mask"], 8, torch.float32), g["alibi"])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cleantransformer_b200.ddp import DistributedDataParallel
    torch.manual_seed(100 + rank)  # different init per rank: the wrapper must sync from rank 0
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    ddp = DistributedDataParallel(model, bucket_cap_mb=0.001)  # tiny cap -> several buckets
    assert len(ddp.buckets) >= 2
    sd0 = [p.detach().clone() for p in model.parameters()]
    torch.manual_seed(7 + rank)
    x = torch.randn(8, 16)
    for step in range(2):
        for p in model.parameters():
            p.grad = None
        y = ddp(x)
        (y ** 2).mean().backward()
    grads = [p.grad.detach().clone() for p in model.parameters()]
    assert all(k.startswith("module.") for k in ddp.state_dict().keys())
    # local (unsynchronised) gradient for the cross-check
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    with torch.no_grad():
        for pr, p in zip(ref.parameters(), sd0):
            pr.copy_(p)
    (ref(x) ** 2).mean().backward()
    out[rank] = (sd0, grads, [p.grad.clone() for p in ref.parameters()])
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_wrapper_world2_gloo():
    """N>1 path on CPU: parameter sync from rank 0, bucketed reduction, averaging."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    p0, g0, l0 = out[0]
    p1, g1, l1 = out[1]
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)  # parameters were broadcast from rank 0
    for a, b, x, y in zip(g0, g1, l0, l1):
        assert torch.allclose(a, b)  # every rank holds the same reduced gradient
        assert rel_err(a, (x + y) / 2) < 1e-6  # and it is the mean of the per-rank gradients


def test_adamw_load_state_dict_lands_in_the_arena(monkeypatch):
    """Resume (trainer / examples/ft_bloom_DDP.py:155-156 checkpoints): torch's load_state_dict hands the
    optimizer cloned moment tensors; the flat AdamW kernel reads the arena buffers, so the loaded moments must be
    copied into the arena and the state re-pointed at the arena views — both when the state is loaded before the
    first step and after it. (Host logic only: no kernel runs.)"""
    from cleantransformer_b200 import optimizer as opt_mod
    monkeypatch.setattr(opt_mod, "_require_cuda", lambda ps: None)

    def make():
        torch.manual_seed(0)
        return [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(9))]

    src = opt_mod.TorchAdamW(make(), lr=1e-3)
    src._setup()
    for i, p in enumerate(src._all_params()):
        src.state[p]["exp_avg"].fill_(0.25 * (i + 1))
        src.state[p]["exp_avg_sq"].fill_(0.5 * (i + 1))
        src.state[p]["step"] += 3
    sd = src.state_dict()
    assert float(src._arena.exp_avg.abs().sum()) > 0  # state tensors ARE arena views

    for setup_first in (False, True):
        dst = opt_mod.TorchAdamW(make(), lr=1e-3)
        if setup_first:
            dst._setup()
        dst.load_state_dict(sd)
        dst._setup()
        a = dst._arena
        for i, p in enumerate(dst._all_params()):
            st = dst.state[p]
            assert st["exp_avg"].data_ptr() == a.param_view(p, a.exp_avg).data_ptr()
            assert st["exp_avg_sq"].data_ptr() == a.param_view(p, a.exp_avg_sq).data_ptr()
            assert torch.all(a.param_view(p, a.exp_avg) == 0.25 * (i + 1))
            assert torch.all(a.param_view(p, a.exp_avg_sq) == 0.5 * (i + 1))
            assert int(st["step"]) == 3


def _fd_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cleantransformer_b200.ddp import DistributedDataParallel as D
    d = object.__new__(D)
    torch.nn.Module.__init__(d)
    d.rank, d.world, d.group = rank, world, None
    r = os.memfd_create("ct_test_%d" % rank)
    os.write(r, b"from rank %d" % rank)
    got = d._exchange_fds("t1", r, list(range(world)), list(range(world)))       # everyone -> everyone
    msgs = {q: os.pread(fd, 64, 0) for q, fd in got.items()}
    r2 = os.memfd_create("ct_test_root")
    os.write(r2, b"root")
    got2 = d._exchange_fds("t2", r2, [0], list(range(world)))                    # rank 0 -> the others
    msgs2 = {q: os.pread(fd, 64, 0) for q, fd in got2.items()}
    out[rank] = (msgs, msgs2)
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_file_descriptor_exchange_world3_gloo():
    """ddp._exchange_fds (how the cuMemCreate allocations and the multicast object reach the peer processes): SCM_RIGHTS
    over abstract Unix sockets, all-to-all and one-to-all, checked with memfds standing in for the CUDA handles."""
    world = 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fd_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        msgs, msgs2 = out[r]
        assert msgs == {q: b"from rank %d" % q for q in range(world) if q != r}
        assert msgs2 == ({} if r == 0 else {0: b"root"})


def test_ddp_tied_table_state_machine():
    """ddp._on_grad_written for a tied token table on the peer-memory path: a dense FIRST gradient (LM-head wgrad)
    launches the table's own bucket at once; later token scatters exchanged by _sparse_embedding_bwd need no launch,
    however many there are (GPT looks tokens_embed up twice with segment_ids, modeling_gpt.py:186-188: three writes);
    a scatter that arrives first stays local and the table is reduced densely once the forward's use count is reached;
    a dense gradient after the early launch, or any write after its bucket was reduced, is an error; ordinary
    parameters count down their bucket against the use count announced by functional.note_use."""
    from cleantransformer_b200.ddp import DistributedDataParallel as D
    d = object.__new__(D)
    torch.nn.Module.__init__(d)
    tied, plain_a, plain_b = [torch.nn.Parameter(torch.zeros(4, 4)) for _ in range(3)]
    launched = []

    def launch(bi, final=False):
        launched.append(bi)
        d._launched[bi] = True

    d._launch = launch
    d._bucket_of = {id(tied): 0, id(plain_a): 1, id(plain_b): 1}
    d._grad_off = {id(tied): 0}
    d.require_backward_grad_sync = True

    def begin(tied_uses=3, a_uses=1):
        d._pending, d._launched, d._writes, d._cb_queued, d._early = [1, 2], [False, False], {}, True, {}
        d._scatter_how = {}
        tied._ct_uses, plain_a._ct_uses, plain_b._ct_uses = tied_uses, a_uses, 0   # plain_b: torch-autograd gradient
        launched.clear()

    begin()
    d._on_grad_written(tied)                       # LM-head wgrad
    assert launched == [0] and d._early[id(tied)] == "dense" and d._pending[0] == 0
    d._on_grad_written(plain_a)
    assert launched == [0] and d._pending[1] == 1
    d._on_grad_written(plain_b)
    assert launched == [0, 1]
    for _ in range(2):                             # token scatter + segment scatter, both exchanged across ranks
        d._scatter_how[id(tied)] = "exchanged"     # what _sparse_embedding_bwd records right before the notification
        d._on_grad_written(tied)
    assert launched == [0, 1]
    with pytest.raises(RuntimeError):              # a write after its bucket was all-reduced would be lost
        d._on_grad_written(plain_a)
    begin()
    d._on_grad_written(tied)
    with pytest.raises(RuntimeError):              # second DENSE contribution after the early all-reduce
        d._on_grad_written(tied)
    begin(tied_uses=2)
    d._scatter_how[id(tied)] = "local"             # scatter first (no dense gradient yet): stays local ...
    d._on_grad_written(tied)
    assert launched == [] and d._early[id(tied)] == "local"
    d._on_grad_written(tied)                       # ... and the table is reduced densely once both writes are in
    assert launched == [0]
    begin(a_uses=2)                                # a module used twice: reduce after the second write only
    d._on_grad_written(plain_a)
    assert d._pending[1] == 2
    d._on_grad_written(plain_a); d._on_grad_written(plain_b)
    assert launched == [1]
    begin()
    d.require_backward_grad_sync = False           # no_sync(): gradients stay local, nothing is launched
    d._on_grad_written(tied); d._on_grad_written(plain_a)
    assert launched == []


def test_generation_loop_matches_the_real_reference(golden):
    """cleantransformer_b200.generation.GenerationMixin against outputs of the REFERENCE's GenerationMixin
    (tests/golden/generation_loop.pt, tools/make_golden.py): greedy and seeded sampling through the temperature /
    top-k / top-p wrappers, one and several end ids, pad ids for finished rows, left-padded prompts — token ids
    bit-exact, shapes included (the reference emits max_gen_len + 2 tokens)."""
    from cleantransformer_b200.generation import GenerationMixin
    g = golden("generation_loop")
    emb = g["emb"]

    class Cfg:
        n_layer = 2

    class Toy(GenerationMixin):
        config = Cfg()

        def __call__(self, ids, attention_mask=None, k_v_pasts=None, **kw):
            n = attention_mask.sum(-1, keepdim=True).float()
            logits = emb[ids] + 0.01 * n[:, :, None]
            return (logits, logits), k_v_pasts

    for cfg, ref in zip(g["cases"], g["outputs"]):
        torch.manual_seed(g["seed"])
        out = Toy().generate(g["ids"].clone(), attention_mask=g["mask"].clone(), generation_configs=dict(cfg))
        assert out.shape == ref.shape and torch.equal(out, ref), cfg


def test_alibi_tensor_is_bit_identical_to_the_reference_construction():
    """modeling_bloom.build_alibi_tensor / alibi_slopes follow the reference's fp32 tensor arithmetic
    (modeling_bloom.py:309-331) — also for head counts that are not powers of two and for the 12 / 16 / 20-head
    cases where a higher-precision formula differs in the last bit. The oracle's restatement is pinned to the
    reference by tests/test_oracle_golden.py."""
    from cleantransformer_b200.models import modeling_bloom as mb
    from oracle import ct_oracle as O
    torch.manual_seed(0)
    for heads in (1, 2, 4, 6, 8, 12, 16, 20, 25, 32, 40):
        mask = (torch.rand(3, 11) > 0.3).long()
        mask[0] = 1
        mask[1, :4] = 0
        for dtype in (torch.float32, torch.bfloat16):
            assert torch.equal(mb.build_alibi_tensor(mask, heads, dtype), O.build_alibi_tensor(mask, heads, dtype)), heads
        assert torch.equal(mb.alibi_slopes(heads), O.alibi_slopes(heads))


def test_head_mask_is_refused_not_ignored():
    """The reference scales the softmax weights by head_mask (transformer.py:48-50, modeling_bloom.py:112-113,
    modeling_gpt.py:95-96); the fused kernels cannot, so a non-None head_mask must raise instead of being dropped."""
    from cleantransformer_b200 import transformer as T
    from cleantransformer_b200.models import modeling_bloom as mb, modeling_gpt as mg
    hm = torch.ones(1)
    x = torch.zeros(1, 4, 32)
    with pytest.raises(NotImplementedError):
        T.AttentionLayer(T.ExampleConfig()).eval()(torch.zeros(1, 4, T.ExampleConfig().hidden_size), None, hm)
    bloom = mb.BloomForCausalLM(mb.BloomConfig(vocab_size=32, hidden_size=32, n_layer=1, num_attention_heads=2)).eval()
    with pytest.raises(NotImplementedError):
        bloom(input_ids=torch.zeros(1, 4, dtype=torch.long), attention_mask=torch.ones(1, 4, dtype=torch.long), head_mask=hm)
    with pytest.raises(NotImplementedError):
        bloom.bloom.blocks[0].self_attention(x, x, alibi=torch.zeros(2, 1, 4), head_mask=hm)
    gpt = mg.AttentionLayer(mg.GPTConfig(vocab_size=32, n_embd=32, n_positions=8, n_layer=1, n_head=2, n_ctx=8)).eval()
    with pytest.raises(NotImplementedError):
        gpt(x, head_mask=hm)


def test_kv_cache_base_is_found_again_after_the_caller_reslices_the_cache():
    """ops._kv_base_of: kv_cache_append leaves the preallocated buffer as an attribute on the view
    it returns; a caller that re-slices / re-wraps that view loses Python attributes, so the buffer is also registered
    by storage address and recognised when the tensor still is a prefix view of it."""
    import weakref
    from cleantransformer_b200 import ops
    base = torch.zeros(2, 3, 16, 8)
    view = base[:, :, :5]
    view._ct_cache_base = base
    ops._KV_BASES[base.untyped_storage().data_ptr()] = weakref.ref(base)
    assert ops._kv_base_of(view) is base
    assert ops._kv_base_of(view[:, :, :5]) is base            # re-sliced: attribute gone, registry finds it
    assert ops._kv_base_of(base[:, :, :7].detach()) is base
    assert ops._kv_base_of(base[:, :, 2:7]) is None           # not a prefix
    assert ops._kv_base_of(base[:, :2, :5]) is None           # other head count
    assert ops._kv_base_of(torch.zeros(2, 3, 5, 8)) is None   # unrelated storage
    del ops._KV_BASES[base.untyped_storage().data_ptr()]


def test_dropout_stream_numbers_follow_the_call_order_and_the_seed_is_torchs():
    """functional.next_dropout: every active dropout site of a forward takes the next stream number; the seed is the one
    torch.manual_seed set unless manual_dropout_seed overrides it — what oracle.DropoutFeeder mirrors."""
    from cleantransformer_b200 import functional as F
    from oracle import ct_oracle as O
    F.manual_dropout_seed(1234)
    a, b = F.next_dropout(0.1), F.next_dropout(0.5)
    assert a == (0.1, 1234, 1) and b == (0.5, 1234, 2)
    F.manual_dropout_seed(99)
    assert F.next_dropout(0.25) == (0.25, 99, 1)
    F._DropoutState.seed = None
    torch.manual_seed(4242)
    assert F.next_dropout(0.1)[1] == 4242
    # the restated generator: deterministic, seed- and stream-dependent, the right rate, p = 0 keeps everything
    m1 = O.dropout_mask_elementwise((64, 1024), 0.3, 7, 1)
    assert torch.equal(m1, O.dropout_mask_elementwise((64, 1024), 0.3, 7, 1))
    assert not torch.equal(m1, O.dropout_mask_elementwise((64, 1024), 0.3, 7, 2))
    assert not torch.equal(m1, O.dropout_mask_elementwise((64, 1024), 0.3, 8, 1))
    assert abs(1 - float(m1.float().mean()) - 0.3) < 0.01
    assert bool(O.dropout_mask_elementwise((8, 8), 0.0, 7, 1).all())
    ma = O.dropout_mask_attention(2, 3, 5, 7, 0.5, 11, 4)
    assert ma.shape == (2, 3, 5, 7) and not torch.equal(ma[0, 0], ma[0, 1]) and not torch.equal(ma[0, 0], ma[1, 0])


def test_async_checkpointer_snapshots_at_call_time_and_groups_by_storage(tmp_path):
    """checkpoint.AsyncCheckpointer (examples/ft_bloom_DDP.py:155-156 is the blocking save it
    replaces): the file holds the values AT THE CALL even when the live tensors change right after it, views of one
    flat buffer are saved as ONE storage holding only the covering range, shared memory stays shared, strides and
    nested containers survive, a failing write surfaces at the next synchronisation point, files appear atomically."""
    import os
    from cleantransformer_b200.checkpoint import AsyncCheckpointer
    flat = torch.arange(1000, dtype=torch.float32)
    w, b = flat[64:192].view(8, 16), flat[256:272]
    tied = w
    col = torch.arange(12.).view(3, 4).t()          # non-contiguous, its own storage
    step = torch.tensor(7.0)
    obj = {"model": {"w": w, "b": b, "tied": tied, "col": col, "empty": torch.empty(0, 3)},
           "opt": {"state": {0: {"step": step, "m": flat[512:640].view(8, 16)}}, "groups": [{"lr": 1e-3, "params": [0]}]},
           "tuple": (1, "x", torch.ones(2, dtype=torch.int64))}
    expect_w, expect_m = w.clone(), flat[512:640].clone()
    with AsyncCheckpointer() as ck:
        path = str(tmp_path / "deep" / "x.pt")
        ck.save({path: obj}, json_files={str(tmp_path / "s.json"): {"global_step": 7}},
                on_done=lambda: open(tmp_path / "done", "w").close())
        flat.zero_()            # the "next optimizer step"
        step += 1
        obj["opt"]["groups"][0]["lr"] = 5.0
        ck.wait()
        assert os.path.exists(tmp_path / "done") and not os.path.exists(path + ".tmp")
        r = torch.load(path, weights_only=False)
        assert torch.equal(r["model"]["w"], expect_w) and torch.equal(r["opt"]["state"][0]["m"].flatten(), expect_m)
        assert float(r["opt"]["state"][0]["step"]) == 7.0 and r["opt"]["groups"][0]["lr"] == 1e-3
        assert r["model"]["w"].data_ptr() == r["model"]["tied"].data_ptr()
        assert r["model"]["w"].untyped_storage().data_ptr() == r["opt"]["state"][0]["m"].untyped_storage().data_ptr()
        assert r["model"]["w"].untyped_storage().nbytes() == (640 - 64) * 4        # the covering range, not the arena
        assert r["model"]["col"].stride() == (1, 4) and torch.equal(r["model"]["col"], col)
        assert r["model"]["empty"].shape == (0, 3) and r["tuple"][:2] == (1, "x")
        assert open(tmp_path / "s.json").read().strip().startswith("{")
        # a module's state_dict keeps its class and `_metadata` (load_state_dict reads the versions from it)
        lin = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2))
        live = lin.state_dict()
        ck.save({str(tmp_path / "sd.pt"): live})
        ck.wait()
        back = torch.load(str(tmp_path / "sd.pt"), weights_only=False)
        assert type(back) is type(live) and back._metadata == live._metadata and list(back) == list(live)
        torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2)).load_state_dict(back, strict=True)
        # a failure in the writer thread is not lost
        os.makedirs(tmp_path / "isdir.pt")
        ck.save({str(tmp_path / "isdir.pt"): {"a": torch.ones(1)}})
        with pytest.raises(RuntimeError, match="asynchronous checkpoint failed"):
            ck.wait()
        ck.save({str(tmp_path / "ok.pt"): {"a": torch.ones(1)}})
        ck.wait()
        assert torch.load(str(tmp_path / "ok.pt"))["a"].item() == 1.0


def test_no_repeat_ngram_processor_and_beam_search_match_the_real_reference(golden):
    """The rest of GenerationMixin.generate against outputs of the REFERENCE (tests/golden/generation_beam.pt,
    tools/make_golden.py): NoRepeatNGramLogitsProcessor (logits_processor.py:11-32; examples/inference_bloom.py:93) on
    its own and inside the greedy / sampling loop, and beam search (generation_util.py:121-290;
    examples/inference_gpt2.py:64) — arg-top-k and seeded sampling over the joint beam x vocabulary scores, candidate
    bookkeeping, early_stop on and off, pad ids for finished rows, beam re-ordering of ids / mask / positions /
    segments / caches — token ids bit-exact, shapes included."""
    from cleantransformer_b200.generation import GenerationMixin, _ban_repeated_ngrams
    g = golden("generation_beam")
    emb = g["emb"]
    pr = g["processor"]
    for n, want in pr["out"].items():
        assert torch.equal(_ban_repeated_ngrams(pr["ids"], pr["scores"].clone(), n), want), n

    class Cfg:
        n_layer = 2

    class Toy(GenerationMixin):
        config = Cfg()
        training = False

        def __call__(self, ids, attention_mask=None, k_v_pasts=None, **kw):
            new = []
            for past in k_v_pasts:
                run = ids.sum(-1, keepdim=True).float() if past is None else past[0] + ids.sum(-1, keepdim=True).float()
                new.append((run, -run))
            n = attention_mask.sum(-1, keepdim=True).float()
            logits = emb[ids] + 0.01 * n[:, :, None] + 0.003 * torch.sin(new[0][0])[:, :, None] * emb[(ids + 1) % 40]
            return (logits, logits), new

    class ToyPos(Toy):
        def __call__(self, ids, attention_mask=None, k_v_pasts=None, position_ids=None, segment_ids=None, **kw):
            (logits, _), new = Toy.__call__(self, ids, attention_mask=attention_mask, k_v_pasts=k_v_pasts)
            logits = logits + 0.02 * emb[(position_ids + 2 * segment_ids) % 40]
            return (logits, logits), new

    assert len(g["cases"]) == 9
    for cfg, ref in zip(g["cases"], g["outputs"]):
        torch.manual_seed(g["seed"])
        out = Toy().generate(g["ids"].clone(), attention_mask=g["mask"].clone(), generation_configs=dict(cfg))
        assert out.shape == ref.shape and torch.equal(out, ref), cfg
    torch.manual_seed(g["seed"])
    out = ToyPos().generate(g["ids"].clone(), attention_mask=g["mask"].clone(), position_ids=g["pos"].clone(),
                            segment_ids=g["seg"].clone(), generation_configs=dict(g["cases"][g["with_pos_case"]]))
    assert torch.equal(out, g["with_pos"])
    seen = []
    Toy().generate(g["ids"].clone(), attention_mask=g["mask"].clone(), generation_configs=dict(g["cases"][3]),
                   steamers=lambda ids: seen.append(tuple(ids.shape)) or len(seen) == 2)
    assert seen == [(3, 3, 5), (3, 3, 6)]            # streamers see [bsz, beam, len] and may stop the loop
