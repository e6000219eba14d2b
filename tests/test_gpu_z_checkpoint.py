"""GPU side of SURVEY §8 N4: the asynchronous checkpoint (cleantransformer_b200/checkpoint.py) around the real flat
arenas, and Trainer save / resume through the CUDA kernels. The blocking save it replaces is
examples/ft_bloom_DDP.py:155-156 / trainer/trainer.py:1303-1342. (File name: last of the GPU suite.)"""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.mark.parametrize("stage", ["device", "host"])
def test_snapshot_holds_the_values_of_the_call_while_the_arena_moves_on(tmp_path, stage):
    """256 MB of arena views: save() only enqueues; the live buffer is overwritten by the very next kernel on the
    training stream (after guard() in stage "host", which is what makes that legal) and the file still holds the
    values of the call. One storage per flat buffer in the file, tied views still tied."""
    from cleantransformer_b200.arena import ParamArena
    from cleantransformer_b200.checkpoint import AsyncCheckpointer
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(4096, 4096, device=DEV)) for _ in range(3)] + \
             [torch.nn.Parameter(torch.randn(4096, device=DEV))]
    arena = ParamArena(params)
    arena.ensure_state()
    arena.exp_avg.normal_()
    want_p = [p.detach().clone() for p in params]
    want_m = arena.exp_avg.clone()
    sd = {"p%d" % i: p.data for i, p in enumerate(params)}
    sd["tied"] = params[0].data
    opt = {"state": {i: {"step": torch.tensor(3.0), "exp_avg": arena.param_view(p, arena.exp_avg)}
                     for i, p in enumerate(params)}}
    torch.cuda.synchronize()
    with AsyncCheckpointer(stage=stage) as ck:
        ck.save({str(tmp_path / "model.bin"): sd, str(tmp_path / "opt.pt"): opt})
        ck.guard()
        arena.flat.add_(1.0)          # "the next optimizer step"
        arena.exp_avg.zero_()
        ck.wait()
    torch.cuda.synchronize()
    r = torch.load(str(tmp_path / "model.bin"))
    ro = torch.load(str(tmp_path / "opt.pt"), weights_only=False)
    for i, w in enumerate(want_p):
        assert torch.equal(r["p%d" % i], w.cpu()), i
        assert torch.equal(ro["state"][i]["exp_avg"], want_m[params[i]._ct_off:params[i]._ct_off + w.numel()].view(w.shape).cpu())
        assert float(ro["state"][i]["step"]) == 3.0
    assert r["tied"].data_ptr() == r["p0"].data_ptr()
    assert len({t.untyped_storage().data_ptr() for t in r.values()}) == 1
    assert torch.equal(params[0].detach(), want_p[0] + 1.0)        # and the live arena did move on


def _toy(tmp_path, max_steps, **extra):
    from cleantransformer_b200.models import modeling_bloom as mb
    from cleantransformer_b200.trainer import Trainer
    cfg = dict(vocab_size=512, hidden_size=128, n_layer=2, num_attention_heads=2)     # head_dim 64: the tcgen05 kernels
    g = torch.Generator().manual_seed(3)
    data = [dict(input_ids=torch.randint(3, 512, (128,), generator=g), attention_mask=torch.ones(128, dtype=torch.long))
            for _ in range(12)]
    for d in data:
        d["labels"] = d["input_ids"].clone()

    def collate(items):
        return {k: torch.stack([it[k] for it in items]) for k in items[0]}

    torch.manual_seed(4)
    with torch.device(DEV):
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02)
    m._tie_weight()
    a = dict(per_device_train_batch_size=4, learning_rate=1e-3, max_steps=max_steps, logging_steps=1,
             output_dir=str(tmp_path), weight_decay=0.01, save_steps=4)
    a.update(extra)
    return Trainer(model=m, args=types.SimpleNamespace(**a), data_collator=collate, train_dataset=data), m


def test_trainer_saves_behind_the_step_loop_and_resumes_on_the_gpu(tmp_path):
    """8 steps with a checkpoint every 4: the folders are reference-shaped and loadable by torch.optim.AdamW, the last
    snapshot IS the final model; a fresh Trainer that loads checkpoint-8 holds the same bits (parameters, and the
    moments adopted INTO its flat arena); a run resumed from checkpoint-4 (middle of epoch 1) sees the same batches and
    follows the same loss curve. (Bit-exactness of a whole resumed trajectory is asserted by the CPU twin of this test,
    where the stand-in kernels are deterministic; here gradients are accumulated with atomics and AdamW turns the
    rounding noise of a mathematically zero gradient, e.g. the key bias, into +-lr.)"""
    import shutil
    straight, m_a = _toy(tmp_path / "a", 8)
    out = straight.train()
    assert out.global_step == 8 and sorted(os.listdir(tmp_path / "a")) == ["checkpoint-4", "checkpoint-8"]
    for f in ("checkpoint-4", "checkpoint-8"):
        assert sorted(os.listdir(tmp_path / "a" / f)) == ["optimizer.pt", "pytorch_model.bin", "rng_state.pth",
                                                            "trainer_state.json"]
    losses = [h["loss"] for h in straight.state.log_history]
    assert losses[-1] < losses[0]
    osd = torch.load(str(tmp_path / "a" / "checkpoint-8" / "optimizer.pt"))
    ref_opt = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in m_a.parameters()], lr=1.0)
    ref_opt.load_state_dict(osd)
    assert all(float(s["step"]) == 8.0 for s in ref_opt.state.values())
    sd8 = torch.load(str(tmp_path / "a" / "checkpoint-8" / "pytorch_model.bin"), map_location=DEV)
    for n, p in m_a.state_dict().items():
        assert torch.equal(sd8[n], p), n

    loaded, m_c = _toy(tmp_path / "c", 8)
    loaded._load_checkpoint(str(tmp_path / "a" / "checkpoint-8"))
    loaded.optimizer._setup()
    assert loaded.state.global_step == 8
    arena = loaded.optimizer._arena
    for (n, pa), (_, pc) in zip(m_a.named_parameters(), m_c.named_parameters()):
        assert torch.equal(pa, pc), n
        sa, sc = straight.optimizer.state[pa], loaded.optimizer.state[pc]
        assert torch.equal(sa["exp_avg"], sc["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sc["exp_avg_sq"]), n
        assert float(sc["step"]) == 8.0 and not sc["step"].is_cuda
        assert sc["exp_avg"].data_ptr() == arena.param_view(pc, arena.exp_avg).data_ptr()

    os.makedirs(tmp_path / "b")
    shutil.copytree(tmp_path / "a" / "checkpoint-4", tmp_path / "b" / "checkpoint-4")
    second, m_b = _toy(tmp_path / "b", 8)
    seen = []
    inner = second.training_step
    second.training_step = lambda model, inputs: (seen.append(inputs["input_ids"].clone()), inner(model, inputs))[1]
    second.train(resume_from_checkpoint=True)
    assert second.state.global_step == 8 and len(seen) == 4
    assert [h["step"] for h in second.state.log_history] == list(range(1, 9))
    assert second.state.log_history[:4] == straight.state.log_history[:4]
    for ha, hb in zip(straight.state.log_history[4:], second.state.log_history[4:]):
        assert abs(ha["loss"] - hb["loss"]) < 2e-3 * abs(ha["loss"]), (ha, hb)
    assert sorted(os.listdir(tmp_path / "b")) == ["checkpoint-4", "checkpoint-8"]
