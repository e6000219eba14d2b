"""Trainer — the class surface of CleanTransformer/trainer/trainer.py (an annotated restatement of the
HuggingFace Trainer that needs accelerate / peft / datasets and has no arithmetic of its own; SURVEY.md
§0 D5). Only its hot-path lines are reproduced: `model(**inputs)` (trainer.py:560), backward (:555) and
`optimizer.step()` (:501), driven by a plain loop over a DataLoader; everything they call runs on the
sm_100a kernels (fused AdamW arena, bucketed P2P DDP). Checkpoints follow the reference's layout
(`checkpoint-<step>/{pytorch_model.bin, optimizer.pt, scheduler.pt, trainer_state.json, rng_state.pth}`,
trainer.py:1303-1342, rotation :1465-1486, resume :351-379, 448-453) but are written by `checkpoint.AsyncCheckpointer`:
the step loop only enqueues three flat copies, the files are written behind it. Callbacks, hub upload and the
accelerate plumbing are out of scope.

Accepted `args`: any object with the TrainingArguments attribute names used below (missing ones take the
HF defaults): per_device_train_batch_size, per_device_eval_batch_size, learning_rate, weight_decay,
adam_beta1, adam_beta2, adam_epsilon, num_train_epochs, max_steps, logging_steps, output_dir, seed,
save_strategy ("steps" | "epoch" | "no"), save_steps, save_total_limit, save_only_model, gradient_accumulation_steps,
max_grad_norm (applied only when the attribute is present and > 0: an HF `TrainingArguments` carries 1.0).
`compute_metrics` / `preprocess_logits_for_metrics` are honoured by evaluate() (trainer.py:621-739); `callbacks` are
accepted for signature compatibility and NOT dispatched (a warning says so).
"""
import json
import math
import os
import random
import re
import shutil
import time
import types
import warnings

import torch

from . import functional as F_
from .checkpoint import AsyncCheckpointer, unwrap
from .optimizer import TorchAdamW

PREFIX_CHECKPOINT_DIR = "checkpoint"
WEIGHTS_NAME, OPTIMIZER_NAME, SCHEDULER_NAME = "pytorch_model.bin", "optimizer.pt", "scheduler.pt"
TRAINER_STATE_NAME, RNG_STATE_NAME = "trainer_state.json", "rng_state.pth"

_DEFAULTS = dict(per_device_train_batch_size=8, per_device_eval_batch_size=8, learning_rate=5e-5,
                 weight_decay=0.0, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, num_train_epochs=3.0,
                 max_steps=-1, logging_steps=500, output_dir="./", seed=42, dataloader_drop_last=False,
                 save_strategy="steps", save_steps=500, save_total_limit=None, save_only_model=False,
                 gradient_accumulation_steps=1, max_grad_norm=None)


def get_last_checkpoint(folder):
    """Highest-numbered `checkpoint-<step>` directory of `folder` that holds a COMPLETE checkpoint (the trainer state
    is the last file of a checkpoint to be written), or None."""
    best, best_step = None, -1
    if not os.path.isdir(folder):
        return None
    for name in os.listdir(folder):
        m = re.fullmatch(PREFIX_CHECKPOINT_DIR + r"-(\d+)", name)
        path = os.path.join(folder, name)
        if m and os.path.isfile(os.path.join(path, TRAINER_STATE_NAME)) and int(m.group(1)) > best_step:
            best, best_step = path, int(m.group(1))
    return best


class TrainOutput(types.SimpleNamespace):
    pass


class EvalPrediction(types.SimpleNamespace):
    """transformers.EvalPrediction: `.predictions`, `.label_ids` (numpy arrays)."""

    def __iter__(self):
        return iter((self.predictions, self.label_ids))


class Trainer:
    def __init__(self, model=None, args=None, data_collator=None, train_dataset=None, eval_dataset=None,
                 tokenizer=None, model_init=None, compute_metrics=None, optimizers=(None, None), callbacks=None,
                 preprocess_logits_for_metrics=None):   # trainer.py:140-153, same positional order
        if model is None:
            if model_init is None:
                raise RuntimeError("`Trainer` requires either a `model` or `model_init` argument")
            model = model_init()
        self.model = model
        self.args = args if args is not None else types.SimpleNamespace()
        self.data_collator = data_collator
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.tokenizer = tokenizer
        self.compute_metrics = compute_metrics
        self.preprocess_logits_for_metrics = preprocess_logits_for_metrics
        self.callbacks = list(callbacks or [])
        if self.callbacks:
            warnings.warn("cleantransformer_b200.Trainer accepts `callbacks` for signature compatibility but does not "
                          "dispatch them")
        self.optimizer, self.lr_scheduler = optimizers
        self.state = types.SimpleNamespace(global_step=0, epoch=0.0, log_history=[])
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self._checkpointer = None
        self._train_generator = None
        self._loss_scale = 1.0

    # ---- helpers -------------------------------------------------------------------------------
    def _arg(self, name):
        return getattr(self.args, name, _DEFAULTS[name])

    def _loader(self, dataset, batch_size, shuffle):
        sampler = None
        if torch.distributed.is_available() and torch.distributed.is_initialized() and shuffle:
            sampler = torch.utils.data.distributed.DistributedSampler(dataset, seed=self._arg("seed"))
        g = torch.Generator()
        g.manual_seed(self._arg("seed"))
        if shuffle:
            self._train_generator = g if sampler is None else None
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle and sampler is None,
                                           sampler=sampler, collate_fn=self.data_collator,
                                           drop_last=self._arg("dataloader_drop_last"),
                                           generator=g if sampler is None and shuffle else None)

    @staticmethod
    def _host_batches(loader):
        """Batches are drawn with the CPU as the default device: the launcher (run.py) makes the GPU the default
        device, under which the sampler's `randperm` would meet its CPU generator on the wrong device."""
        with torch.device("cpu"):
            it = iter(loader)
        while True:
            with torch.device("cpu"):
                try:
                    batch = next(it)
                except StopIteration:
                    return
            yield batch

    def get_train_dataloader(self):
        if self.train_dataset is None:
            raise ValueError("Trainer: training requires a train_dataset.")
        return self._loader(self.train_dataset, self._arg("per_device_train_batch_size"), True)

    def get_eval_dataloader(self, eval_dataset=None):
        ds = eval_dataset if eval_dataset is not None else self.eval_dataset
        if ds is None:
            raise ValueError("Trainer: evaluation requires an eval_dataset.")
        return self._loader(ds, self._arg("per_device_eval_batch_size"), False)

    def create_optimizer(self):
        if self.optimizer is None:
            self.optimizer = TorchAdamW(self.model.parameters(), lr=self._arg("learning_rate"),
                                        betas=(self._arg("adam_beta1"), self._arg("adam_beta2")),
                                        eps=self._arg("adam_epsilon"), weight_decay=self._arg("weight_decay"))
        return self.optimizer

    def _prepare_inputs(self, inputs):
        if self.device is None:
            return inputs
        return {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in inputs.items()}

    @staticmethod
    def _loss_of(outputs):
        """trainer.py:558-588 `compute_loss`: dict['loss'] or first element; the reference models return
        ((loss, logits, hidden), k_v_pasts)."""
        if isinstance(outputs, dict):
            return outputs["loss"]
        first = outputs[0]
        if isinstance(first, (tuple, list)):
            first = first[0]
        return first

    def compute_loss(self, model, inputs, return_outputs=False):
        outputs = model(**inputs)
        loss = self._loss_of(outputs)
        return (loss, outputs) if return_outputs else loss

    def training_step(self, model, inputs):
        """trainer.py:543-556; the loss is scaled by 1 / gradient_accumulation_steps for the backward (the accumulated
        gradient is the mean over the micro-batches, as accelerate's `backward` makes it); the returned loss is not."""
        model.train()
        inputs = self._prepare_inputs(inputs)
        loss = self.compute_loss(model, inputs)
        (loss * self._loss_scale if self._loss_scale != 1.0 else loss).backward()
        return loss.detach()

    # ---- checkpoints ---------------------------------------------------------------------------
    def is_world_process_zero(self):
        d = torch.distributed
        return not (d.is_available() and d.is_initialized()) or d.get_rank() == 0

    def _checkpointer_(self):
        if self._checkpointer is None:
            self._checkpointer = AsyncCheckpointer()
        return self._checkpointer

    def _rng_state(self):
        st = {"python": random.getstate(), "cpu": torch.random.get_rng_state(),
              "ct_dropout": (F_._DropoutState.seed, F_._DropoutState.stream)}
        try:
            import numpy as np
            st["numpy"] = np.random.get_state()
        except ImportError:
            pass
        if torch.cuda.is_available():
            st["cuda"] = torch.cuda.random.get_rng_state()
        return st

    def _load_rng_state(self, checkpoint):
        path = os.path.join(checkpoint, RNG_STATE_NAME)
        if not os.path.isfile(path):
            return
        st = torch.load(path, weights_only=False)
        random.setstate(st["python"])
        torch.random.set_rng_state(st["cpu"])
        if "numpy" in st:
            import numpy as np
            np.random.set_state(st["numpy"])
        if "cuda" in st and torch.cuda.is_available():
            torch.cuda.random.set_rng_state(st["cuda"])
        F_._DropoutState.seed, F_._DropoutState.stream = st.get("ct_dropout", (None, 0))

    def _save_checkpoint(self, model=None, metrics=None):
        """trainer.py:1303-1342. Every rank may call it (like the reference); rank 0 writes. Returns the folder.
        Nothing is on disk yet when it returns — `self.wait_for_checkpoints()` (called at the end of train()) or the
        next save are the synchronisation points; `trainer_state.json` is written last and marks the folder complete."""
        if not self.is_world_process_zero():
            return None
        run_dir = self._arg("output_dir")
        out = os.path.join(run_dir, "%s-%d" % (PREFIX_CHECKPOINT_DIR, self.state.global_step))
        os.makedirs(out, exist_ok=True)
        files = {os.path.join(out, WEIGHTS_NAME): unwrap(model if model is not None else self.model).state_dict()}
        if not self._arg("save_only_model"):
            if self.optimizer is not None:
                files[os.path.join(out, OPTIMIZER_NAME)] = self.optimizer.state_dict()
            if self.lr_scheduler is not None:
                files[os.path.join(out, SCHEDULER_NAME)] = self.lr_scheduler.state_dict()
            files[os.path.join(out, RNG_STATE_NAME)] = self._rng_state()
        state = {"global_step": self.state.global_step, "epoch": self.state.epoch,
                 "log_history": list(self.state.log_history)}
        if metrics is not None:
            state["metrics"] = {k: float(v) for k, v in metrics.items()}
        limit = self._arg("save_total_limit")
        ck = self._checkpointer_()
        ck.save(files, on_done=lambda: (self._write_state(out, state), self._rotate_checkpoints(run_dir, limit)))
        return out

    @staticmethod
    def _write_state(out, state):
        tmp = os.path.join(out, TRAINER_STATE_NAME + ".tmp")
        with open(tmp, "w") as f:
            json.dump(state, f, indent=2, sort_keys=True)
        os.replace(tmp, os.path.join(out, TRAINER_STATE_NAME))

    @staticmethod
    def _sorted_checkpoints(run_dir):
        found = []
        for name in os.listdir(run_dir):
            m = re.fullmatch(PREFIX_CHECKPOINT_DIR + r"-(\d+)", name)
            if m and os.path.isdir(os.path.join(run_dir, name)):
                found.append((int(m.group(1)), os.path.join(run_dir, name)))
        return [p for _, p in sorted(found)]

    @classmethod
    def _rotate_checkpoints(cls, run_dir, limit):
        """trainer.py:1465-1486: keep the `save_total_limit` newest folders. Runs in the writer thread AFTER the new
        checkpoint is complete, so a crash never leaves fewer complete checkpoints than the limit."""
        if limit is None or limit <= 0:
            return
        for old in cls._sorted_checkpoints(run_dir)[:-limit]:
            shutil.rmtree(old, ignore_errors=True)

    def wait_for_checkpoints(self):
        if self._checkpointer is not None:
            self._checkpointer.wait()

    def _load_checkpoint(self, checkpoint):
        """trainer.py:351-379, 1516-1597, 1619-1654: weights, optimizer / scheduler state, trainer state."""
        weights = os.path.join(checkpoint, WEIGHTS_NAME)
        if not os.path.isfile(weights):
            raise ValueError("Can't find a valid checkpoint at %s" % checkpoint)
        target = unwrap(self.model)
        dev = next(target.parameters()).device
        target.load_state_dict(torch.load(weights, map_location=dev), strict=True)
        F_.invalidate_shadows(target)
        opt = self.create_optimizer()
        if os.path.isfile(os.path.join(checkpoint, OPTIMIZER_NAME)):
            opt.load_state_dict(torch.load(os.path.join(checkpoint, OPTIMIZER_NAME), map_location="cpu"))
        if self.lr_scheduler is not None and os.path.isfile(os.path.join(checkpoint, SCHEDULER_NAME)):
            self.lr_scheduler.load_state_dict(torch.load(os.path.join(checkpoint, SCHEDULER_NAME), weights_only=False))
        with open(os.path.join(checkpoint, TRAINER_STATE_NAME)) as f:
            st = json.load(f)
        self.state.global_step = int(st["global_step"])
        self.state.epoch = float(st.get("epoch", 0.0))
        self.state.log_history = list(st.get("log_history", []))

    # ---- public API ----------------------------------------------------------------------------
    def train(self, resume_from_checkpoint=None, **kwargs):
        loader = self.get_train_dataloader()
        opt = self.create_optimizer()
        max_steps = self._arg("max_steps")
        epochs = self._arg("num_train_epochs")
        if len(loader) == 0:
            raise ValueError("Trainer.train: the training dataloader is empty")
        gas = max(1, int(self._arg("gradient_accumulation_steps")))
        steps_per_epoch = max(1, -(-len(loader) // gas))    # optimizer steps; a short tail of micro-batches still steps
        if max_steps is None or max_steps <= 0:
            max_steps = int(math.ceil(epochs * steps_per_epoch))
        log_every = max(1, int(self._arg("logging_steps")))
        strategy = str(self._arg("save_strategy")).lower().replace("intervalstrategy.", "")
        save_every = max(1, int(self._arg("save_steps")))
        epoch, skip = 0, 0
        if resume_from_checkpoint:
            if isinstance(resume_from_checkpoint, bool):
                resume_from_checkpoint = get_last_checkpoint(self._arg("output_dir"))
                if resume_from_checkpoint is None:
                    raise ValueError("No valid checkpoint found in output directory (%s)" % self._arg("output_dir"))
            self._load_checkpoint(resume_from_checkpoint)
            epoch, skip = divmod(self.state.global_step, steps_per_epoch)
        t0, running, last, n_running, step_loss = time.time(), None, float("nan"), 0, None
        self._loss_scale = 1.0 / gas
        ck = None
        while self.state.global_step < max_steps:
            # the order of an epoch depends on (seed, epoch) only, so a resumed run replays nothing to find its place
            if hasattr(loader.sampler, "set_epoch"):
                loader.sampler.set_epoch(epoch)
            if self._train_generator is not None:
                self._train_generator.manual_seed(self._arg("seed") + epoch)
            it = self._host_batches(loader)
            in_epoch = skip
            for _ in range(skip * gas):   # trainer.py:448-451: batches of the interrupted epoch already trained on
                next(it)
            skip = 0
            if resume_from_checkpoint:
                self._load_rng_state(resume_from_checkpoint)   # trainer.py:453
                resume_from_checkpoint = None
            micro = 0
            for inputs in it:
                if micro == 0:
                    opt.zero_grad()
                micro += 1
                last_micro = micro == gas or (in_epoch * gas + micro) == len(loader)   # a short tail still steps
                if not last_micro and hasattr(self.model, "no_sync"):
                    with self.model.no_sync():      # DDP: reduce once per optimizer step
                        loss = self.training_step(self.model, inputs)
                else:
                    loss = self.training_step(self.model, inputs)
                step_loss = loss / gas if step_loss is None else step_loss + loss / gas
                if not last_micro:
                    continue
                micro = 0
                clip = getattr(self.args, "max_grad_norm", None)
                if clip is not None and clip > 0:    # trainer.py:486-493
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), clip)
                if ck is not None:
                    ck.guard()   # a snapshot still reading the live parameters / moments finishes before they change
                opt.step()
                if self.lr_scheduler is not None:
                    self.lr_scheduler.step()
                self.state.global_step += 1
                in_epoch += 1
                self.state.epoch = epoch + in_epoch / max(1, steps_per_epoch)
                running = step_loss if running is None else running + step_loss
                step_loss = None
                n_running += 1
                if self.state.global_step % log_every == 0:
                    last = float(running) / n_running
                    running, n_running = None, 0
                    self.state.log_history.append({"step": self.state.global_step, "loss": last})
                if strategy == "steps" and self.state.global_step % save_every == 0:
                    self._save_checkpoint(self.model)
                    ck = self._checkpointer
                if self.state.global_step >= max_steps:
                    break
            epoch += 1
            if strategy == "epoch" and in_epoch == steps_per_epoch:
                self._save_checkpoint(self.model)
                ck = self._checkpointer
        if running is not None:
            last = float(running) / max(1, n_running)
        self.wait_for_checkpoints()
        return TrainOutput(global_step=self.state.global_step, training_loss=last,
                           metrics={"train_runtime": time.time() - t0, "train_loss": last})

    @torch.no_grad()
    def evaluate(self, eval_dataset=None, ignore_keys=None, metric_key_prefix="eval"):
        """trainer.py:591-739: mean loss over the batches; with `compute_metrics`, the logits (after
        `preprocess_logits_for_metrics(logits, labels)`) and the labels of every batch are gathered on the host and
        handed over as `EvalPrediction(predictions, label_ids)`; metric names get the prefix."""
        if isinstance(eval_dataset, dict):
            metrics = {}
            for name, ds in eval_dataset.items():
                metrics.update(self.evaluate(ds, ignore_keys, "%s_%s" % (metric_key_prefix, name)))
            return metrics
        loader = self.get_eval_dataloader(eval_dataset)
        self.model.eval()
        tot, n, preds, labels = 0.0, 0, [], []
        for inputs in self._host_batches(loader):
            inputs = self._prepare_inputs(inputs)
            loss, outputs = self.compute_loss(self.model, inputs, return_outputs=True)
            tot += float(loss)
            n += 1
            if self.compute_metrics is not None:
                logits = self._logits_of(outputs)
                lab = inputs.get("labels")
                if self.preprocess_logits_for_metrics is not None:
                    logits = self.preprocess_logits_for_metrics(logits, lab)
                preds.append(logits.detach().float().cpu())
                if lab is not None:
                    labels.append(lab.detach().cpu())
        metrics = {}
        if self.compute_metrics is not None and preds:
            metrics = dict(self.compute_metrics(EvalPrediction(
                predictions=torch.cat(preds).numpy(), label_ids=torch.cat(labels).numpy() if labels else None)))
        metrics = {(k if k.startswith(metric_key_prefix + "_") else "%s_%s" % (metric_key_prefix, k)):
                   (v.item() if hasattr(v, "item") else v) for k, v in metrics.items()}
        metrics[metric_key_prefix + "_loss"] = tot / max(n, 1)
        metrics[metric_key_prefix + "_batches"] = n
        return metrics

    @staticmethod
    def _logits_of(outputs):
        """What `prediction_step` keeps (trainer.py:741-787): everything but the loss; the reference models return
        ((loss, logits, hidden), k_v_pasts) — the logits are the second element of the first tuple."""
        if isinstance(outputs, dict):
            return outputs["logits"]
        first = outputs[0]
        if isinstance(first, (tuple, list)):
            return first[1]
        return outputs[1]

    def save_model(self, output_dir=None, _internal_call=False):
        """trainer.py:1347-1404 (plain-module branch: `torch.save(state_dict, pytorch_model.bin)`). On disk when it
        returns."""
        if not self.is_world_process_zero():
            return
        out = output_dir or self._arg("output_dir")
        ck = self._checkpointer_()
        ck.save({os.path.join(out, WEIGHTS_NAME): unwrap(self.model).state_dict()})
        ck.wait()
