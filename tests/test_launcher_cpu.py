"""The launcher (cleantransformer_b200/run.py) that lets the reference's example scripts run unmodified:
alias installation, torch.optim / DDP swaps, and — when the reference checkout is present (build
container only, never on the GPU box) — the reference's own examples/inference_bloom.py loader executed
UNCHANGED against this package's classes."""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _run(args, cwd=None, env_extra=None):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1")
    env.update(env_extra or {})
    r = subprocess.run([sys.executable] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
    return r.stdout


def test_install_aliases_and_swaps(tmp_path):
    script = tmp_path / "probe.py"
    script.write_text(textwrap.dedent('''
        import sys
        from CleanTransformer.models.modeling_bloom import BloomForCausalLM, BloomConfig, BloomAttentionLayer
        from CleanTransformer.models.modeling_gpt import GPTLMHeadModel, GPTConfig, Conv1D
        from CleanTransformer.models.modeling_bert import BertForSequenceClassification, BertTokenizer, BertConfig
        from CleanTransformer.transformer import MultiHeadAttention, AttentionLayer, LayerNorm, TransformerBlock
        from CleanTransformer.generation.generation_util import GenerationMixin
        from CleanTransformer.trainer.trainer import Trainer
        from CleanTransformer.optimizer import AdamW as RefAdamW, SGD
        from torch.optim import AdamW
        from torch.nn.parallel import DistributedDataParallel as DDP
        from transformers import BloomTokenizerFast
        assert MultiHeadAttention is AttentionLayer
        for c in (BloomForCausalLM, GPTLMHeadModel, BertForSequenceClassification, LayerNorm, GenerationMixin, Trainer,
                  RefAdamW, SGD, AdamW, DDP):
            assert c.__module__.startswith("cleantransformer_b200"), (c, c.__module__)
        assert sys.argv[1:] == ["--flag", "7"], sys.argv
        print("PROBE-OK")
    '''))
    out = _run(["-m", "cleantransformer_b200.run", "--ct-keep-default-device", str(script), "--flag", "7"])
    assert "PROBE-OK" in out
    if os.path.isdir(os.path.join(REF, "CleanTransformer")):
        # started from the reference root (as documented): what is NOT mirrored still imports — from the reference
        script2 = tmp_path / "probe2.py"
        script2.write_text(textwrap.dedent('''
            from CleanTransformer.loss import CrossEntropyLoss                  # reference file (outside the hot path)
            from CleanTransformer.models.modeling_bloom import BloomForCausalLM  # mirror
            assert CrossEntropyLoss.__module__ == "CleanTransformer.loss"
            assert BloomForCausalLM.__module__ == "cleantransformer_b200.models.modeling_bloom"
            print("PROBE2-OK")
        '''))
        assert "PROBE2-OK" in _run(["-m", "cleantransformer_b200.run", "--ct-keep-default-device", str(script2)], cwd=REF)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")
def test_reference_inference_bloom_loader_runs_unchanged(tmp_path):
    """examples/inference_bloom.py (unmodified, imported from the read-only reference tree) builds OUR
    BloomForCausalLM through its own load_config / load_model: HF-style config synonyms, HF key remap,
    strict state_dict load, eval(), _tie_weight()."""
    cfg = dict(vocab_size=97, n_embed=32, n_layer=2, num_attention_heads=4, layer_norm_epsilon=1e-5,
               hidden_dropout=0.0, attention_dropout=0.0)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    torch.manual_seed(0)
    H, L, V = 32, 2, 97
    sd = {"transformer.word_embeddings.weight": torch.randn(V, H),
          "transformer.word_embeddings_layernorm.weight": torch.ones(H), "transformer.word_embeddings_layernorm.bias": torch.zeros(H),
          "transformer.ln_f.weight": torch.ones(H), "transformer.ln_f.bias": torch.zeros(H)}
    shapes = {"input_layernorm": (H,), "self_attention.query_key_value": (3 * H, H), "self_attention.dense": (H, H),
              "post_attention_layernorm": (H,), "mlp.dense_h_to_4h": (4 * H, H), "mlp.dense_4h_to_h": (H, 4 * H)}
    for i in range(L):
        for name, shp in shapes.items():
            sd["transformer.h.%d.%s.weight" % (i, name)] = torch.randn(*shp)
            sd["transformer.h.%d.%s.bias" % (i, name)] = torch.randn(shp[0])
    torch.save(sd, tmp_path / "pytorch_model.bin")
    script = tmp_path / "use_ref_loader.py"
    script.write_text(textwrap.dedent('''
        import sys, torch
        sys.path.insert(0, %r)
        from examples.inference_bloom import load_model, load_config      # the reference's own file
        config = load_config(%r)
        model = load_model(config, %r)
        assert type(model).__module__ == "cleantransformer_b200.models.modeling_bloom", type(model).__module__
        assert model.lm_head.weight is model.bloom.word_embeddings.weight and not model.training
        assert len(model.bloom.blocks) == 2 and model.bloom.blocks[0].self_attention.num_heads == 4
        sd = torch.load(%r)
        assert torch.equal(model.bloom.blocks[1].mlp.dense_4h_to_h.weight, sd["transformer.h.1.mlp.dense_4h_to_h.weight"])
        # ... and generates with the example's OWN generation_configs (inference_bloom.py:87-98: sampling with
        # temperature / top-k / top-p AND no_repeat_ngram_size 2) and gpt2's beam_size 3 (inference_gpt2.py:63-73);
        # plain-torch stand-ins replace the CUDA kernels here (no GPU in this container)
        sys.path.insert(0, %r)
        import mock_ops
        ids = torch.tensor([[3, 3, 11, 12, 13, 14], [21, 22, 23, 24, 25, 26]])
        mask = torch.tensor([[0, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]])       # padding_side='left'
        cfgs = {"beam_size": 1, "max_gen_len": 12, "end_ids": 2, "pad_id": 3, "early_stop": True,
                "no_repeat_ngram_size": 2, "do_sample": True, "temperature": 0.8, "top_k": 10, "top_p": 0.8}
        with mock_ops.patched():
            torch.manual_seed(0)
            out = model.generate(input_ids=ids, attention_mask=mask, generation_configs=cfgs)
            rows = out.numpy().tolist()
            assert out.shape[:2] == (2, 1) and 7 <= out.shape[2] <= 6 + 12 + 2
            for row in rows:
                seq = row[0][6:]
                grams = [tuple(seq[i:i + 2]) for i in range(len(seq) - 1) if 3 not in seq[i:i + 2]]
                assert len(grams) == len(set(grams)), seq          # no bigram repeats among the generated tokens
            beams = model.generate(input_ids=ids, attention_mask=mask, generation_configs=dict(cfgs, beam_size=3))
            assert beams.shape == (2, 3, 6 + 12 + 2)
        print("REF-LOADER-OK")
    ''' % (REF, str(tmp_path / "config.json"), str(tmp_path / "pytorch_model.bin"), str(tmp_path / "pytorch_model.bin"),
           os.path.join(ROOT, "tests"))))
    out = _run(["-m", "cleantransformer_b200.run", "--ct-keep-default-device", str(script)], cwd=str(tmp_path))
    assert "REF-LOADER-OK" in out


def test_trainer_surface_cpu_side():
    """Constructor signature / helper behaviour that does not need a GPU."""
    from cleantransformer_b200.trainer import Trainer
    with pytest.raises(RuntimeError):
        Trainer()
    t = Trainer(model=torch.nn.Linear(2, 2), args=None, train_dataset=[{"x": torch.zeros(2)}] * 5)
    assert t._arg("learning_rate") == 5e-5 and len(t.get_train_dataloader()) == 1
    assert float(Trainer._loss_of(((torch.tensor(3.0), None, None), None))) == 3.0
    assert float(Trainer._loss_of({"loss": torch.tensor(2.0)})) == 2.0
    for name in ("train", "evaluate", "save_model", "compute_loss", "training_step"):
        assert callable(getattr(t, name))
    import inspect
    assert list(inspect.signature(Trainer.__init__).parameters)[1:] == [       # trainer/trainer.py:140-153
        "model", "args", "data_collator", "train_dataset", "eval_dataset", "tokenizer", "model_init", "compute_metrics",
        "optimizers", "callbacks", "preprocess_logits_for_metrics"]
    assert list(inspect.signature(Trainer.train).parameters)[1:3] == ["resume_from_checkpoint", "kwargs"] or \
        "resume_from_checkpoint" in inspect.signature(Trainer.train).parameters


def test_launcher_async_save_flag_routes_torch_save_and_flushes_at_exit(tmp_path):
    """`--ct-async-save`: the reference scripts' `torch.save(model.state_dict(), path)`
    (examples/ft_bloom_DDP.py:155-156) goes through checkpoint.save_async when the object holds device tensors and is
    untouched otherwise; whatever is still queued is on disk when the interpreter exits."""
    script = tmp_path / "saver.py"
    script.write_text(textwrap.dedent('''
        import sys, torch
        from cleantransformer_b200 import checkpoint
        sd = {"w": torch.arange(6.).view(2, 3), "nested": {"step": torch.tensor(2.0)}}
        torch.save(sd, sys.argv[1] + "/plain.pt")                       # CPU tensors: the real torch.save, on disk now
        assert torch.load(sys.argv[1] + "/plain.pt")["w"].sum() == 15
        assert checkpoint._default is None
        checkpoint.has_device_tensor = lambda obj: True                 # stand-in for "holds CUDA tensors"
        torch.save(sd, sys.argv[1] + "/async.pt")
        sd["w"].zero_()
        assert checkpoint._default is not None
        with open(sys.argv[1] + "/buffer.bin", "wb") as f:              # file objects keep the blocking path
            torch.save(sd, f)
        print("SAVER-OK")
    '''))
    out = _run(["-m", "cleantransformer_b200.run", "--ct-keep-default-device", "--ct-async-save", str(script), str(tmp_path)])
    assert "SAVER-OK" in out
    import torch
    assert torch.load(str(tmp_path / "async.pt"))["w"].sum() == 15 and os.path.exists(tmp_path / "buffer.bin")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "CleanTransformer")), reason="reference checkout not present")
def test_checkpoints_load_into_the_reference_classes(tmp_path):
    """what Trainer / AsyncCheckpointer write (plain state_dicts, examples/ft_bloom_DDP.py:155-156) is
    read back by the REFERENCE's own classes: strict `load_state_dict` into CleanTransformer's BloomForCausalLM and
    the same logits from its own forward; `optimizer.pt` continues a `torch.optim.AdamW` (what ft_bloom.py:70 builds)
    with the same next step as this package's AdamW takes."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mock_ops
    from cleantransformer_b200.models import modeling_bloom as mb
    from cleantransformer_b200.trainer import Trainer
    cfg = dict(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4, layer_norm_epsilon=1e-5,
               hidden_dropout=0.0, attention_dropout=0.0)
    g = torch.Generator().manual_seed(3)
    data = [dict(input_ids=torch.randint(3, 64, (10,), generator=g), attention_mask=torch.ones(10, dtype=torch.long))
            for _ in range(8)]
    for d in data:
        d["labels"] = d["input_ids"].clone()
    collate = lambda items: {k: torch.stack([it[k] for it in items]) for k in items[0]}   # noqa: E731
    with mock_ops.patched():
        torch.manual_seed(4)
        m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        m._tie_weight()
        args = types.SimpleNamespace(per_device_train_batch_size=4, learning_rate=5e-3, max_steps=3, save_steps=3,
                                     output_dir=str(tmp_path), weight_decay=0.01)
        tr = Trainer(model=m, args=args, data_collator=collate, train_dataset=data)
        tr.train()
        ids, mask = data[0]["input_ids"][None], data[0]["attention_mask"][None]
        m.eval()
        with torch.no_grad():
            (logits, _), _ = m(input_ids=ids, attention_mask=mask)
        batch = collate(data[:4])
        m.train()
        tr.optimizer.zero_grad()
        (loss, _, _), _ = m(**batch)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        tr.optimizer.step()
        after = {n: p.detach().clone() for n, p in m.named_parameters()}
    torch.save({"cfg": cfg, "ids": ids, "mask": mask, "logits": logits, "grads": grads, "after": after},
               tmp_path / "expect.pt")
    script = tmp_path / "load_in_reference.py"
    script.write_text(textwrap.dedent('''
        import sys, torch
        sys.path.insert(0, %r)
        from CleanTransformer.models.modeling_bloom import BloomForCausalLM, BloomConfig     # the reference itself
        assert BloomForCausalLM.__module__ == "CleanTransformer.models.modeling_bloom"
        e = torch.load(%r, weights_only=False)
        ck = %r
        model = BloomForCausalLM(BloomConfig(**e["cfg"]))
        model.load_state_dict(torch.load(ck + "/pytorch_model.bin"), strict=True)
        model._tie_weight()
        model.eval()
        with torch.no_grad():
            (logits, _), _ = model(input_ids=e["ids"], attention_mask=e["mask"])
        err = float((logits - e["logits"]).abs().max() / e["logits"].abs().max())
        assert err < 1e-5, err
        # continue the run with the reference's optimizer from optimizer.pt, fed this package's gradients
        params = dict(model.named_parameters())
        opt = torch.optim.AdamW(list(model.parameters()), lr=5e-3, weight_decay=0.01)
        opt.load_state_dict(torch.load(ck + "/optimizer.pt"))
        for n, p in params.items():
            p.grad = e["grads"][n].clone()
        opt.step()
        worst = max(float((p - e["after"][n]).abs().max() / e["after"][n].abs().max()) for n, p in params.items())
        assert worst < 1e-5, worst
        print("REF-CHECKPOINT-OK %%.1e %%.1e" %% (err, worst))
    ''' % (REF, str(tmp_path / "expect.pt"), str(tmp_path / "checkpoint-3"))))
    out = _run([str(script)], cwd=str(tmp_path), env_extra={"PYTHONPATH": ""})
    assert "REF-CHECKPOINT-OK" in out, out


def test_launcher_default_device_patches_keep_cpu_scripts_working(tmp_path):
    """With the GPU as default device (what install() does for the examples' bare `torch.tensor(...)` inputs) a script
    written for the CPU still shuffles its DataLoader (examples/ft_bloom.py:58, ft_bloom_DDP.py:71: the samplers hand a
    CPU generator to torch.randperm) and still calls `.numpy()` on its results (inference_bloom.py:100). Simulated
    here with the `meta` device standing in for the GPU."""
    script = tmp_path / "probe.py"
    script.write_text(textwrap.dedent('''
        import torch
        from torch.utils.data import RandomSampler
        from torch.utils.data.distributed import DistributedSampler
        from cleantransformer_b200 import run
        data = list(range(10))
        torch.set_default_device("meta")
        # (the samplers are driven directly with an explicit generator: DataLoader itself also draws a base seed with
        # `.item()`, which the meta stand-in cannot do and a real GPU can)
        try:
            list(RandomSampler(data, generator=torch.Generator().manual_seed(0)))
            broken = False
        except Exception:
            broken = True
        assert broken, "torch handles CPU generators under a device default now: the patch can go"
        run._patch_for_default_device(torch)
        assert sorted(RandomSampler(data, generator=torch.Generator().manual_seed(0))) == data
        ds = DistributedSampler(data, num_replicas=2, rank=1, shuffle=True, seed=3)
        assert len(list(ds)) == 5
        g = torch.Generator().manual_seed(1)
        assert torch.randint(0, 5, (3,), generator=g).device.type == "cpu"
        assert torch.randn(3).device.type == "meta"                  # everything else follows the default device
        torch.set_default_device("cpu")
        assert torch.arange(3.).numpy().tolist() == [0.0, 1.0, 2.0]
        print("PATCH-OK")
    '''))
    out = _run([str(script)], cwd=str(tmp_path))
    assert "PATCH-OK" in out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")
def test_reference_ft_bloom_train_function_runs_unchanged(tmp_path):
    """examples/ft_bloom.py `train()` (:63-97) — imported from the read-only reference tree, not a line changed — drives
    THIS package: its `from torch.optim import AdamW` resolves to TorchAdamW, its model is the Bloom mirror, its
    `torch.save(model.state_dict(), ...)` goes through the asynchronous checkpointer (`--ct-async-save`). Kernels are
    replaced by the plain-torch stand-ins (no GPU here); the loss falls step after step and the saved checkpoint reloads."""
    script = tmp_path / "use_ref_train.py"
    script.write_text(textwrap.dedent('''
        import io, contextlib, os, re, sys, torch
        sys.path.insert(0, %r)
        sys.path.insert(0, %r)
        import mock_ops
        import examples.ft_bloom as ref                                     # the reference's own file
        from CleanTransformer.models.modeling_bloom import BloomForCausalLM, BloomConfig
        assert ref.AdamW.__module__ == "cleantransformer_b200.optimizer", ref.AdamW
        assert BloomForCausalLM.__module__ == "cleantransformer_b200.models.modeling_bloom"
        from cleantransformer_b200 import checkpoint
        checkpoint.has_device_tensor = lambda obj: True                     # stand-in for "holds CUDA tensors"
        torch.manual_seed(999)
        model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4))
        model._tie_weight()
        g = torch.Generator().manual_seed(3)
        batches = []
        for _ in range(2):
            ids = torch.randint(3, 64, (4, 10), generator=g)
            batches.append({"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone(),
                            "prompts": ["x"] * 4})
        out = io.StringIO()
        with mock_ops.patched(), contextlib.redirect_stdout(out):
            # the reference's loop moves every tensor of a batch in place: hand it fresh dicts per epoch
            class Loader:
                def __iter__(self):
                    return iter([dict(b) for b in batches])
            ref.train(model, Loader(), epoches=10, save_interval=10, print_interval=1, save_dir=%r)
        losses = [float(x) for x in re.findall(r"loss: ([0-9.eE+-]+)", out.getvalue())]
        # (the reference hard-codes lr = 1e-5, ft_bloom.py:70: a slow but strictly monotonic descent on both batches)
        assert len(losses) == 20 and all(b < a for a, b in zip(losses[:-2], losses[2:])), losses
        checkpoint.default_checkpointer().wait()
        sd = torch.load(os.path.join(%r, "model_step_20.pt"))
        fresh = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4))
        fresh.load_state_dict(sd, strict=True)
        assert all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), model.state_dict().values()))
        print("REF-TRAIN-OK %%.3f -> %%.3f" %% (losses[0], losses[-1]))
    ''' % (REF, os.path.join(ROOT, "tests"), str(tmp_path / "ckpt"), str(tmp_path / "ckpt"))))
    out = _run(["-m", "cleantransformer_b200.run", "--ct-keep-default-device", "--ct-async-save", str(script)], cwd=REF)
    assert "REF-TRAIN-OK" in out, out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")
@pytest.mark.parametrize("amp", [False, True], ids=["plain", "torch_amp_branch"])
def test_reference_ft_bloom_ddp_train_function_runs_unchanged_on_two_ranks(tmp_path, amp):
    """examples/ft_bloom_DDP.py `train()` (:77-156), unmodified, under torchrun with two gloo ranks: its
    `DDP(model, device_ids=[local_rank])` is this package's wrapper (rank 0's parameters are broadcast, gradients are
    averaged bucket by bucket), its AdamW the flat-arena one, its rank-0 `torch.save(model.state_dict())` keeps the
    `module.` prefix. Replicas end identical although every rank saw its own batches. `torch_amp_branch`: the
    `--use_torch_amp` path (:108-128; GradScaler and autocast disable themselves without CUDA) — the branch that never
    calls `optimizer.zero_grad()`, so every backward ACCUMULATES into the already averaged gradients and the wrapper
    reduces the sum again, as torch's reducer does."""
    import socket
    script = tmp_path / "use_ref_ddp_train.py"
    script.write_text(textwrap.dedent('''
        import io, contextlib, os, sys, torch
        import torch.distributed as dist
        sys.path.insert(0, %r)
        sys.path.insert(0, %r)
        import mock_ops
        import examples.ft_bloom_DDP as ref                                 # the reference's own file
        from CleanTransformer.models.modeling_bloom import BloomForCausalLM, BloomConfig
        assert ref.DDP.__module__ == "cleantransformer_b200.ddp" and ref.AdamW.__module__ == "cleantransformer_b200.optimizer"
        dist.init_process_group("gloo")
        rank = dist.get_rank()
        torch.manual_seed(100 + rank)                                       # replicas start DIFFERENT
        model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=2, num_attention_heads=4))
        model._tie_weight()
        g = torch.Generator().manual_seed(7 + rank)                         # ... and see different data
        batches = []
        for _ in range(2):
            ids = torch.randint(3, 64, (2, 10), generator=g)
            batches.append({"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone()})

        class Sampler:
            epochs = []
            def set_epoch(self, e):
                self.epochs.append(e)

        class Loader:
            sampler = Sampler()
            def __iter__(self):
                return iter([dict(b) for b in batches])

        with mock_ops.patched(), contextlib.redirect_stdout(io.StringIO()):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref.train(model, Loader(), epoches=2, save_interval=4, print_interval=1, save_dir=%r, use_torch_amp=%r)
        assert Loader.sampler.epochs == [0, 1]
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        both = [torch.empty_like(flat) for _ in range(2)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                                # identical replicas after 4 steps
        dist.barrier()
        if rank == 0:
            sd = torch.load(os.path.join(%r, "model_step_4.pt"))
            assert all(k.startswith("module.") for k in sd) and "module.lm_head.weight" in sd
            assert torch.equal(sd["module.bloom.ln_f.weight"], model.bloom.ln_f.weight.detach())
            print("REF-DDP-TRAIN-OK")
        dist.destroy_process_group()
    ''' % (REF, os.path.join(ROOT, "tests"), str(tmp_path / "ckpt"), amp, str(tmp_path / "ckpt"))))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = _run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", str(port), "-m", "cleantransformer_b200.run", "--ct-keep-default-device", str(script)],
               cwd=REF)
    assert "REF-DDP-TRAIN-OK" in out, out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")
def test_reference_gpt2_and_bert_loaders_run_unchanged_and_agree_with_huggingface(tmp_path):
    """examples/inference_gpt2.py and inference_bert.py `load_config` / `load_model` (their HF -> reference key maps,
    :14-44 / :14-52), unmodified, build THIS package's GPTLMHeadModel / BertForSequenceClassification from random-init
    HuggingFace checkpoints (strict load); the logits then agree with the HuggingFace models' own (the independent
    cross-oracle of SURVEY §8 c5; plain-torch stand-ins for the kernels, fp32)."""
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    gcfg = transformers.GPT2Config(vocab_size=101, n_embd=48, n_layer=2, n_head=4, n_positions=64, n_ctx=64,
                                   resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    hf_gpt = transformers.GPT2LMHeadModel(gcfg).eval()
    gsd = {k: v.clone() for k, v in hf_gpt.transformer.state_dict().items()}
    for i in range(2):   # the causal buffer is part of the reference's state_dict (modeling_gpt.py:57-58)
        gsd["h.%d.attn.bias" % i] = torch.tril(torch.ones(64, 64)).view(1, 1, 64, 64)
    os.makedirs(tmp_path / "gpt2")
    torch.save(gsd, tmp_path / "gpt2" / "pytorch_model.bin")
    (tmp_path / "gpt2" / "config.json").write_text(json.dumps(
        dict(vocab_size=101, n_embd=48, n_layer=2, n_head=4, n_positions=64, n_ctx=64, afn="gelu_new",
             layer_norm_epsilon=1e-5)))
    ids = torch.randint(1, 101, (2, 9))
    with torch.no_grad():
        g_logits = hf_gpt(input_ids=ids).logits
    bcfg = transformers.BertConfig(vocab_size=120, hidden_size=32, num_hidden_layers=12, num_attention_heads=4,
                                   intermediate_size=64, max_position_embeddings=40, num_labels=5,
                                   hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf_bert = transformers.BertForSequenceClassification(bcfg).eval()
    os.makedirs(tmp_path / "bert")
    torch.save(hf_bert.state_dict(), tmp_path / "bert" / "pytorch_model.bin")
    (tmp_path / "bert" / "config.json").write_text(json.dumps(
        dict(vocab_size=120, hidden_size=32, num_hidden_layers=12, num_attention_heads=4, intermediate_size=64,
             max_position_embeddings=40, num_labels=5, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)))
    bids = torch.randint(1, 120, (3, 11))
    bmask = torch.ones_like(bids)
    bmask[1, 7:] = 0
    with torch.no_grad():
        b_logits = hf_bert(input_ids=bids, attention_mask=bmask, token_type_ids=torch.zeros_like(bids)).logits
    torch.save({"ids": ids, "g_logits": g_logits, "bids": bids, "bmask": bmask, "b_logits": b_logits},
               tmp_path / "expect.pt")
    script = tmp_path / "use_ref_loaders.py"
    script.write_text(textwrap.dedent('''
        import sys, torch
        sys.path.insert(0, %r)
        sys.path.insert(0, %r)
        import mock_ops
        from examples import inference_gpt2, inference_bert                 # the reference's own files
        e = torch.load(%r)
        d = %r
        gpt = inference_gpt2.load_model(inference_gpt2.load_config(d + "/gpt2/config.json"), d + "/gpt2/pytorch_model.bin")
        bert = inference_bert.load_model(inference_bert.load_config(d + "/bert/config.json"), d + "/bert/pytorch_model.bin").eval()
        assert type(gpt).__module__ == "cleantransformer_b200.models.modeling_gpt" and not gpt.training
        assert type(bert).__module__ == "cleantransformer_b200.models.modeling_bert"
        with mock_ops.patched(), torch.no_grad():
            (logits, _), _ = gpt(e["ids"], attention_mask=torch.ones_like(e["ids"]))
            b_logits = bert(e["bids"], e["bmask"], torch.zeros_like(e["bids"]), torch.arange(e["bids"].shape[1]))
        ge = float((logits - e["g_logits"]).abs().max() / e["g_logits"].abs().max())
        be = float((b_logits - e["b_logits"]).abs().max() / e["b_logits"].abs().max())
        assert ge < 1e-5 and be < 1e-5, (ge, be)
        print("REF-LOADERS-OK %%.1e %%.1e" %% (ge, be))
    ''' % (REF, os.path.join(ROOT, "tests"), str(tmp_path / "expect.pt"), str(tmp_path))))
    out = _run(["-m", "cleantransformer_b200.run", "--ct-keep-default-device", str(script)], cwd=REF)
    assert "REF-LOADERS-OK" in out, out
