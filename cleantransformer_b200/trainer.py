"""Trainer — the class surface of CleanTransformer/trainer/trainer.py (an annotated restatement of the
HuggingFace Trainer that needs accelerate / peft / datasets and has no arithmetic of its own; SURVEY.md
§0 D5). Only its hot-path lines are reproduced: `model(**inputs)` (trainer.py:560), backward (:555) and
`optimizer.step()` (:501), driven by a plain loop over a DataLoader; everything they call runs on the
sm_100a kernels (fused AdamW arena, bucketed P2P DDP). Checkpoint rotation, callbacks, hub upload and
the accelerate plumbing are out of scope.

Accepted `args`: any object with the TrainingArguments attribute names used below (missing ones take the
HF defaults): per_device_train_batch_size, per_device_eval_batch_size, learning_rate, weight_decay,
adam_beta1, adam_beta2, adam_epsilon, num_train_epochs, max_steps, logging_steps, output_dir, seed.
"""
import math
import os
import time
import types

import torch

from .optimizer import TorchAdamW

_DEFAULTS = dict(per_device_train_batch_size=8, per_device_eval_batch_size=8, learning_rate=5e-5,
                 weight_decay=0.0, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, num_train_epochs=3.0,
                 max_steps=-1, logging_steps=500, output_dir="./", seed=42, dataloader_drop_last=False)


class TrainOutput(types.SimpleNamespace):
    pass


class Trainer:
    def __init__(self, model=None, args=None, data_collator=None, train_dataset=None, eval_dataset=None,
                 tokenizer=None, model_init=None, compute_metrics=None, callbacks=None, optimizers=(None, None),
                 preprocess_logits_for_metrics=None):
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
        self.optimizer, self.lr_scheduler = optimizers
        self.state = types.SimpleNamespace(global_step=0, epoch=0.0, log_history=[])
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None

    # ---- helpers -------------------------------------------------------------------------------
    def _arg(self, name):
        return getattr(self.args, name, _DEFAULTS[name])

    def _loader(self, dataset, batch_size, shuffle):
        sampler = None
        if torch.distributed.is_available() and torch.distributed.is_initialized() and shuffle:
            sampler = torch.utils.data.distributed.DistributedSampler(dataset, seed=self._arg("seed"))
        g = torch.Generator()
        g.manual_seed(self._arg("seed"))
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle and sampler is None,
                                           sampler=sampler, collate_fn=self.data_collator,
                                           drop_last=self._arg("dataloader_drop_last"),
                                           generator=g if sampler is None and shuffle else None)

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
        model.train()
        inputs = self._prepare_inputs(inputs)
        loss = self.compute_loss(model, inputs)
        loss.backward()
        return loss.detach()

    # ---- public API ----------------------------------------------------------------------------
    def train(self, resume_from_checkpoint=None, **kwargs):
        if resume_from_checkpoint:
            raise NotImplementedError("checkpoint resume is outside the hot path this package covers")
        loader = self.get_train_dataloader()
        opt = self.create_optimizer()
        max_steps = self._arg("max_steps")
        epochs = self._arg("num_train_epochs")
        if max_steps is None or max_steps <= 0:
            max_steps = int(math.ceil(epochs * len(loader)))
        log_every = max(1, int(self._arg("logging_steps")))
        t0, running, last = time.time(), None, float("nan")
        epoch = 0
        while self.state.global_step < max_steps:
            if hasattr(loader.sampler, "set_epoch"):
                loader.sampler.set_epoch(epoch)
            for inputs in loader:
                opt.zero_grad()
                loss = self.training_step(self.model, inputs)
                opt.step()
                if self.lr_scheduler is not None:
                    self.lr_scheduler.step()
                self.state.global_step += 1
                running = loss if running is None else running + loss
                if self.state.global_step % log_every == 0:
                    last = float(running) / log_every
                    running = None
                    self.state.log_history.append({"step": self.state.global_step, "loss": last})
                if self.state.global_step >= max_steps:
                    break
            epoch += 1
            self.state.epoch = float(epoch)
        if running is not None:
            last = float(running) / max(1, self.state.global_step % log_every)
        return TrainOutput(global_step=self.state.global_step, training_loss=last,
                           metrics={"train_runtime": time.time() - t0, "train_loss": last})

    @torch.no_grad()
    def evaluate(self, eval_dataset=None, ignore_keys=None, metric_key_prefix="eval"):
        loader = self.get_eval_dataloader(eval_dataset)
        self.model.eval()
        tot, n = 0.0, 0
        for inputs in loader:
            inputs = self._prepare_inputs(inputs)
            tot += float(self.compute_loss(self.model, inputs))
            n += 1
        return {metric_key_prefix + "_loss": tot / max(n, 1), metric_key_prefix + "_batches": n}

    def save_model(self, output_dir=None, _internal_call=False):
        out = output_dir or self._arg("output_dir")
        os.makedirs(out, exist_ok=True)
        model = self.model.module if hasattr(self.model, "module") else self.model
        torch.save(model.state_dict(), os.path.join(out, "pytorch_model.bin"))
