"""Run an UNMODIFIED reference script (examples/ft_bloom.py, ft_bloom_DDP.py, inference_*.py) on the
sm_100a kernels:

    python -m cleantransformer_b200.run examples/ft_bloom.py --config_fn ... --ckpt ...
    torchrun --nproc-per-node 8 -m cleantransformer_b200.run examples/ft_bloom_DDP.py ...

What `install()` does before the script is executed with runpy:
  * `CleanTransformer` and its sub-modules (transformer, optimizer, models.modeling_{bloom,gpt,bert},
    generation.generation_util, trainer.trainer) resolve to this package's mirrors, so
    `from CleanTransformer.models.modeling_bloom import BloomForCausalLM` (examples/inference_bloom.py:12)
    builds the kernel-backed classes with the reference's names and state_dict keys;
  * `torch.optim.AdamW` (examples/ft_bloom.py:19) -> optimizer.TorchAdamW (one fused kernel over a flat
    arena) and `torch.nn.parallel.DistributedDataParallel` (examples/ft_bloom_DDP.py:17) -> ddp.
    DistributedDataParallel (bucketed P2P all-reduce over NVSwitch);
  * `transformers.BloomTokenizerFast` (removed from recent transformers releases) is aliased to
    PreTrainedTokenizerFast;
  * with `--ct-async-save`, `torch.save(obj, path)` of anything that holds CUDA tensors (the periodic
    `torch.save(model.state_dict(), ...)` of examples/ft_bloom_DDP.py:155-156) is snapshotted on the device and
    written behind the step loop by `checkpoint.AsyncCheckpointer`; files are complete at interpreter exit at the
    latest (opt-in: a script that reads its own checkpoint back right away must keep the blocking save);
  * the default device becomes the local GPU, because the inference examples build their inputs with
    bare `torch.tensor(...)` (examples/inference_bert.py:71-73) and there is no CPU fallback here; `Tensor.numpy()`
    of a CUDA tensor copies to the host first, because the same scripts print `generated.numpy().tolist()`.
Nothing is copied from or written to the reference tree; the script's directory layout
(`sys.path.append('.')`, `from examples.inference_bloom import ...`) works as it does upstream when the
launcher is started from the reference root.
"""
import importlib
import os
import runpy
import sys
import types

_ALIASES = {
    "CleanTransformer.transformer": "cleantransformer_b200.transformer",
    "CleanTransformer.optimizer": "cleantransformer_b200.optimizer",
    "CleanTransformer.models.modeling_bloom": "cleantransformer_b200.models.modeling_bloom",
    "CleanTransformer.models.modeling_gpt": "cleantransformer_b200.models.modeling_gpt",
    "CleanTransformer.models.modeling_bert": "cleantransformer_b200.models.modeling_bert",
    "CleanTransformer.generation.generation_util": "cleantransformer_b200.generation",
    "CleanTransformer.trainer.trainer": "cleantransformer_b200.trainer",
}
_PACKAGES = ["CleanTransformer", "CleanTransformer.models", "CleanTransformer.generation", "CleanTransformer.trainer"]
_installed = {}


def install(swap_optimizer=True, swap_ddp=True, default_device=True, async_save=False):
    """Idempotent. Returns a dict describing what was patched (used by the tests)."""
    if _installed:
        return _installed
    import torch

    for name in _PACKAGES:
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            # a package: the mirrored sub-modules come from sys.modules (below); anything this package does not mirror
            # (CleanTransformer/loss.py, tokenizers.py: outside the hot path) is imported from the reference checkout the
            # launcher was started in, if there is one
            here = os.path.join(os.getcwd(), *name.split("."))
            pkg.__path__ = [here] if os.path.isdir(here) else []
            pkg.__doc__ = "alias package installed by cleantransformer_b200.run"
            sys.modules[name] = pkg
    for alias, target in _ALIASES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    for name in _PACKAGES[1:]:
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules[parent], leaf, sys.modules[name])
    _installed["aliases"] = sorted(_ALIASES)

    try:
        import transformers
        if not hasattr(transformers, "BloomTokenizerFast"):
            transformers.BloomTokenizerFast = transformers.PreTrainedTokenizerFast
            _installed["BloomTokenizerFast"] = "PreTrainedTokenizerFast"
    except Exception as ex:  # transformers is only needed by the scripts' tokenizers
        _installed["transformers_error"] = repr(ex)

    if swap_optimizer:
        from .optimizer import TorchAdamW
        _installed["torch.optim.AdamW"] = torch.optim.AdamW
        torch.optim.AdamW = TorchAdamW
    if swap_ddp:
        from .ddp import DistributedDataParallel
        _installed["DistributedDataParallel"] = torch.nn.parallel.DistributedDataParallel
        torch.nn.parallel.DistributedDataParallel = DistributedDataParallel
    if async_save:
        from . import checkpoint
        real_save = torch.save
        _installed["torch.save"] = real_save

        def save(obj, f, *args, **kwargs):
            if isinstance(f, (str, os.PathLike)) and not args and not kwargs and checkpoint.has_device_tensor(obj):
                return checkpoint.save_async(obj, f)
            return real_save(obj, f, *args, **kwargs)

        torch.save = save
    if default_device and torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        torch.set_default_device("cuda:%d" % local)
        _installed["default_device"] = "cuda:%d" % local
        _patch_for_default_device(torch)
    return _installed


def _patch_for_default_device(torch):
    """What else has to give once a non-CPU default device is set behind a script written for the CPU:
      * `generated_sequence.numpy().tolist()` (examples/inference_bloom.py:100, inference_gpt2.py:76): tensors that
        became device tensors only because of the default device are brought back first;
      * DataLoader shuffling (examples/ft_bloom.py:58 `shuffle=True`, ft_bloom_DDP.py:71 DistributedSampler): the
        samplers call `torch.randperm(n, generator=<CPU generator>)`, which the default device would turn into a
        device-side call with a CPU generator — an error. A factory call that is handed a CPU generator stays on
        the CPU."""
    _numpy = torch.Tensor.numpy
    _installed["Tensor.numpy"] = _numpy

    def numpy(self, *args, **kwargs):
        return _numpy(self if self.device.type == "cpu" else self.cpu(), *args, **kwargs)

    torch.Tensor.numpy = numpy

    def keep_cpu_generators_on_the_cpu(fn):
        def wrapper(*args, **kwargs):
            g = kwargs.get("generator")
            if g is not None and g.device.type == "cpu" and kwargs.get("device") is None:
                kwargs["device"] = "cpu"
            return fn(*args, **kwargs)
        wrapper.__wrapped__ = fn
        return wrapper

    for name in ("randperm", "randint", "rand", "randn"):
        _installed["torch." + name] = getattr(torch, name)
        setattr(torch, name, keep_cpu_generators_on_the_cpu(getattr(torch, name)))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    flags = {"swap_optimizer": True, "swap_ddp": True, "default_device": True, "async_save": False}
    while argv and argv[0].startswith("--ct-"):
        flag = argv.pop(0)
        key = {"--ct-keep-torch-adamw": "swap_optimizer", "--ct-keep-torch-ddp": "swap_ddp",
               "--ct-keep-default-device": "default_device"}.get(flag)
        if flag == "--ct-async-save":
            flags["async_save"] = True
            continue
        if key is None:
            raise SystemExit("unknown launcher flag %s" % flag)
        flags[key] = False
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    install(**flags)
    sys.argv = argv
    here = os.getcwd()
    if here not in sys.path:
        sys.path.insert(0, here)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
