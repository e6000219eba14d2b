"""Optimizers — host-side mirror of CleanTransformer/optimizer.py over the fused CUDA kernels.

Two surfaces, because the reference has two:
  * `SGD`, `AdamW`  — the reference's own classes (optimizer.py:12-97): same constructor signature,
    same public attributes (`momentum_buffer`, `rmsp_buffer`, `steps`), same arithmetic — AdamW is
    COUPLED-L2 Adam with the step counter starting at 1, exactly as written there. One deliberate
    difference: the parameter iterable is list()-ed, so passing `model.parameters()` works (the
    reference silently trains nothing when handed a generator, optimizer.py:60,78).
  * `TorchAdamW`     — what the examples actually construct (`from torch.optim import AdamW`,
    examples/ft_bloom.py:19,70): a torch.optim.Optimizer subclass with torch's decoupled-decay
    arithmetic, param_groups and state_dict, backed by ONE kernel launch over a flat arena.

There is no CPU path: parameters must live on a CUDA device.
"""
import torch

from . import ops
from .checkpoint import guard_pending
from .arena import ParamArena, arena_of


def _require_cuda(params):
    for p in params:
        if not p.is_cuda:
            raise RuntimeError("cleantransformer_b200 optimizers run on CUDA tensors only "
                               "(no CPU fallback); move the model to the GPU first")


class SGD:
    """CleanTransformer/optimizer.py:12-50."""

    def __init__(self, params, lr=0.01, momentum=None, dampening=0, weight_decay=None):
        self.params = list(params)
        self.lr = lr
        self.momentum = momentum
        self.dampening = dampening
        self.momentum_buffer = [None for _ in self.params]
        self.weight_decay = weight_decay

    def zero_grad(self):
        for param in self.params:
            if param.grad is not None:
                param.grad = None

    @torch.no_grad()
    def step(self):
        _require_cuda(self.params)
        guard_pending()
        for i, param in enumerate(self.params):
            if param.grad is None:
                continue
            g = param.grad
            if not g.is_contiguous() or g.dtype != torch.float32:
                raise RuntimeError("SGD: gradients must be contiguous fp32")
            first = False
            buf = None
            if self.momentum:
                buf = self.momentum_buffer[i]
                if buf is None:
                    buf = torch.empty_like(param, memory_format=torch.contiguous_format)
                    self.momentum_buffer[i] = buf
                    first = True
            ops.sgd_step(param.data, g, buf, self.lr, self.momentum or 0.0, self.dampening or 0.0,
                         self.weight_decay or 0.0, first)
            param._ct_shadow_ver = -1  # raw-pointer update: invalidate the bf16 shadow


class AdamW:
    """CleanTransformer/optimizer.py:53-97 (coupled L2; bias-corrected; t starts at 1)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = list(params)
        self.lr = lr
        self.beta1, self.beta2 = betas
        self.eps = eps
        self.momentum_buffer = [0 for _ in self.params]
        self.rmsp_buffer = [0 for _ in self.params]
        self.steps = [1 for _ in self.params]
        self.weight_decay = weight_decay
        self._arena = None

    def zero_grad(self):
        for param in self.params:
            if param.grad is not None:
                param.grad = None

    def _ensure_state(self):
        if isinstance(self.momentum_buffer[0], torch.Tensor) if self.params else True:
            return
        _require_cuda(self.params)
        a = arena_of(self.params)
        if a is None and all(getattr(p, "_ct_arena", None) is None for p in self.params):
            a = ParamArena(self.params)
        self._arena = a
        if a is not None:
            a.ensure_state()
            self.momentum_buffer = [a.param_view(p, a.exp_avg) for p in self.params]
            self.rmsp_buffer = [a.param_view(p, a.exp_avg_sq) for p in self.params]
        else:
            self.momentum_buffer = [torch.zeros_like(p) for p in self.params]
            self.rmsp_buffer = [torch.zeros_like(p) for p in self.params]

    @torch.no_grad()
    def step(self):
        if not self.params:
            return
        guard_pending()
        self._ensure_state()
        a = self._arena
        wd = self.weight_decay or 0.0
        if a is not None and a.grads_complete() and len(set(self.steps)) == 1:
            ops.adamw_step(a.flat, a.grad, a.exp_avg, a.exp_avg_sq, self.lr, self.beta1, self.beta2,
                           self.eps, wd, self.steps[0], mode=1, shadow=a.shadow)
            a.mark_shadow_fresh()
            if a.shadow is None:
                for p in self.params:
                    p._ct_shadow_ver = -1
            self.steps = [s + 1 for s in self.steps]
            return
        by_step = {}
        for i, p in enumerate(self.params):
            if p.grad is not None:
                by_step.setdefault(self.steps[i], []).append(i)
        for t, idx in by_step.items():
            ps = [self.params[i].data for i in idx]
            ops.adamw_multi(ps, [self.params[i].grad for i in idx], [self.momentum_buffer[i] for i in idx],
                            [self.rmsp_buffer[i] for i in idx], self.lr, self.beta1, self.beta2, self.eps,
                            wd, t, mode=1)
            for i in idx:
                self.steps[i] += 1
                self.params[i]._ct_shadow_ver = -1


class TorchAdamW(torch.optim.Optimizer):
    """Drop-in for torch.optim.AdamW (decoupled weight decay) as used by examples/ft_bloom*.py.
    Same defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2), same state_dict layout
    (per-parameter 'step', 'exp_avg', 'exp_avg_sq')."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **unused):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._arena = None
        self._arena_checked = False

    def _all_params(self):
        return [p for g in self.param_groups for p in g["params"]]

    def _setup(self):
        if self._arena_checked:
            return
        self._arena_checked = True
        ps = self._all_params()
        _require_cuda(ps)
        a = arena_of(ps)
        if a is None and all(getattr(p, "_ct_arena", None) is None for p in ps) \
                and all(p.dtype == torch.float32 for p in ps):
            a = ParamArena(ps)
        self._arena = a
        if a is not None:
            a.ensure_state()
        for p in ps:
            st = self.state[p]
            if "step" not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32, device="cpu")  # never on the GPU: step() reads it
                if a is not None:
                    st["exp_avg"] = a.param_view(p, a.exp_avg)
                    st["exp_avg_sq"] = a.param_view(p, a.exp_avg_sq)
                else:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        self._adopt_state()

    def _adopt_state(self):
        """Moments that did not come from the arena (torch's load_state_dict hands every parameter freshly
        cloned tensors: trainer resume, examples/ft_bloom_DDP.py:155-156 checkpoints) are copied INTO the arena
        and the state is re-pointed at the arena views, so the flat kernel keeps seeing what state_dict() shows."""
        a = self._arena
        if a is None:
            return
        for p in self._all_params():
            st = self.state.get(p)
            if not st:
                continue
            for key, buf in (("exp_avg", a.exp_avg), ("exp_avg_sq", a.exp_avg_sq)):
                view = a.param_view(p, buf)
                cur = st.get(key)
                if cur is None:
                    st[key] = view
                elif cur.data_ptr() != view.data_ptr():
                    view.copy_(cur.to(device=view.device, dtype=view.dtype).reshape(view.shape))
                    st[key] = view
            if not torch.is_tensor(st.get("step")):
                st["step"] = torch.tensor(float(st.get("step", 0)), dtype=torch.float32, device="cpu")

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        if self._arena_checked:
            self._adopt_state()  # before the first step _setup() does it

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._setup()
        guard_pending()   # an asynchronous checkpoint still reading the live parameters / moments finishes first
        a = self._arena
        groups = self.param_groups
        uniform = all((g["lr"], g["betas"], g["eps"], g["weight_decay"]) ==
                      (groups[0]["lr"], groups[0]["betas"], groups[0]["eps"], groups[0]["weight_decay"])
                      for g in groups)
        ps = self._all_params()
        steps = {int(self.state[p]["step"]) for p in ps}
        if a is not None and uniform and len(steps) == 1 and a.grads_complete():
            g0 = groups[0]
            t = steps.pop() + 1
            ops.adamw_step(a.flat, a.grad, a.exp_avg, a.exp_avg_sq, g0["lr"], g0["betas"][0], g0["betas"][1],
                           g0["eps"], g0["weight_decay"], t, mode=0, grad_scale=grad_scale, shadow=a.shadow)
            a.mark_shadow_fresh()
            for p in ps:
                self.state[p]["step"] += 1
                if a.shadow is None:
                    p._ct_shadow_ver = -1  # no shadow buffer (not a CUDA arena): the cached casts are stale
            return loss
        for g in groups:
            by_step = {}
            for p in g["params"]:
                if p.grad is not None:
                    by_step.setdefault(int(self.state[p]["step"]), []).append(p)
            for t0, plist in by_step.items():
                ops.adamw_multi([p.data for p in plist], [p.grad for p in plist],
                                [self.state[p]["exp_avg"] for p in plist],
                                [self.state[p]["exp_avg_sq"] for p in plist], g["lr"], g["betas"][0],
                                g["betas"][1], g["eps"], g["weight_decay"], t0 + 1, mode=0,
                                grad_scale=grad_scale)
                for p in plist:
                    self.state[p]["step"] += 1
                    p._ct_shadow_ver = -1
        return loss
