"""Flat parameter arena: one contiguous fp32 buffer each for parameters, gradients and optimizer
moments, plus a bf16 shadow of the parameters for the tensor cores.

Memory layout in HBM (N = padded parameter count; Bloom-560M: 559.2 M -> 2.24 GB per f32 buffer):
    params  f32 [N]   every nn.Parameter is re-pointed (p.data) at a 256-byte-aligned slice
    grads   f32 [N]   p.grad views; wgrad kernels write here directly (functional.grad_buffer);
                      the DDP wrapper allocates this buffer from peer-mapped memory so the
                      all-reduce kernel works in place
    exp_avg, exp_avg_sq f32 [N]
    shadow  bf16 [N]  refreshed by the same AdamW pass that updates params
One `ct_adamw_step` launch then covers the whole model: 28 B/param of HBM traffic (+2 B shadow).
"""
import torch

ALIGN = 64  # elements: 256 B for f32, 128 B for bf16 -> every view is TMA-legal
SHADOW_ON_ANY_DEVICE = False  # tests/mock_ops.py sets it so that the shadow bookkeeping can be exercised on CPU


class ParamArena:
    def __init__(self, params, grad_buffer=None, shadow_dtype=torch.bfloat16):
        params = [p for p in params]
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        if not uniq:
            raise ValueError("ParamArena: no parameters (was a generator already exhausted?)")
        dev = uniq[0].device
        for p in uniq:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("ParamArena needs fp32 parameters on one device")
        self.params = uniq
        self.offsets = []
        off = 0
        for p in uniq:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.device = dev
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        if grad_buffer is not None:
            assert grad_buffer.numel() >= off and grad_buffer.dtype == torch.float32
            self.grad = grad_buffer[:off]
            self.grad.zero_()
        else:
            self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.shadow = torch.empty(off, dtype=shadow_dtype, device=dev) \
            if (dev.type == "cuda" or SHADOW_ON_ANY_DEVICE) else None
        self.exp_avg = None
        self.exp_avg_sq = None
        with torch.no_grad():
            for p, o in zip(uniq, self.offsets):
                n = p.numel()
                view = self.flat[o:o + n].view(p.shape)
                view.copy_(p.data)
                old_grad = p.grad
                p.data = view
                p._ct_arena = self
                p._ct_off = o
                p._ct_grad_view = self.grad[o:o + n].view(p.shape)
                if old_grad is not None:
                    p._ct_grad_view.copy_(old_grad)
                    p.grad = p._ct_grad_view
                if self.shadow is not None:
                    p._ct_shadow_view = self.shadow[o:o + n].view(p.shape)
                    p._ct_shadow = None
                    p._ct_shadow_ver = -1

    def ensure_state(self):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)

    def grads_complete(self):
        """True when every parameter's .grad is its arena view (so the flat kernel may run)."""
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != p._ct_grad_view.data_ptr():
                return False
        return True

    def mark_shadow_fresh(self):
        for p in self.params:
            if getattr(p, "_ct_shadow_view", None) is not None:
                p._ct_shadow = p._ct_shadow_view
                p._ct_shadow_ver = p._version
                p._ct_shadow_ptr = p.data_ptr()

    def param_view(self, p, buf):
        o = p._ct_off
        return buf[o:o + p.numel()].view(p.shape)


def arena_of(params):
    """The common arena of `params` if they were all flattened together, else None."""
    a = None
    for p in params:
        pa = getattr(p, "_ct_arena", None)
        if pa is None or (a is not None and pa is not a):
            return None
        a = pa
    if a is None or len(a.params) != len(list(params)):
        return None
    return a
