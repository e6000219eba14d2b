"""CUDA-graph capture of a fixed-shape training step (forward + loss + backward).

One Bloom-560M step is 620 kernel launches from two Python threads (forward in the caller, backward in
autograd's device thread); profiles/r01w_* put ~3 % of the step into the gaps between them. The reference's
examples train with fixed shapes (`padding='max_length'`, examples/ft_bloom.py:125-131), so the whole
forward/backward can be captured once and replayed: the kernels, their arguments (TMA descriptors included — they
are `__grid_constant__` kernel parameters) and every intermediate buffer are then static.

What stays outside the graph: the optimizer step (its bias-correction constants are host scalars that change every
step: one launch) and anything that feeds new data (copied into the static input buffers before the replay).

`DistributedDataParallel(comm="p2p")` is captured with the step: the bucket all-reduces are ordinary kernels on a side
stream that forks from / joins the capture stream, and their epochs live in device memory (csrc/comm.cu), so every
replay synchronises with the peers' replays. Every rank must capture and replay in step. `comm="nccl"` is not
captured (library collectives bring their own graph rules).

Parameters never enter autograd as edges of this package's Functions (functional._anchor), so no stale AccumulateGrad
node — created on another stream by an earlier, un-captured step — can tie the capture to uncaptured work.
"""
import torch


class GraphedTrainStep:
    """`step = GraphedTrainStep(model, example_inputs)`, then per iteration
    `loss = step(**inputs); optimizer.step()`.

    example_inputs: dict of CUDA tensors with the (fixed) shapes / dtypes of every later call; it must make the model
    return a loss (the reference's causal-LM forward returns `((loss, logits, hidden), k_v_pasts)`,
    modeling_bloom.py:218-232; `loss_of` picks it out, default: first element, recursively).
    Gradients land where the un-graphed path puts them (`functional.grad_buffer`: the parameter's `.grad`, i.e. the
    optimizer arena once one exists), overwritten — not accumulated — by every replay."""

    def __init__(self, model, example_inputs, loss_of=None, warmup=3):
        from .ddp import DistributedDataParallel
        if isinstance(model, DistributedDataParallel) and model.world > 1 and not getattr(model, "_peer_mem", False):
            raise RuntimeError("GraphedTrainStep: only the peer-memory collectives (comm='p2p') are captured with the "
                               "step; run comm='nccl' un-graphed")
        for k, v in example_inputs.items():
            if not (torch.is_tensor(v) and v.is_cuda):
                raise RuntimeError("GraphedTrainStep: input '%s' must be a CUDA tensor" % k)
        self.model = model
        self.loss_of = loss_of or self._first
        self.static = {k: v.clone() for k, v in example_inputs.items()}
        params = [p for p in model.parameters() if p.requires_grad]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up off the capture: lazy initialisation, cudaFuncSetAttribute, autotuning
            for _ in range(max(1, warmup)):
                for p in params:
                    p.grad = None
                self.loss_of(model(**self.static)).backward()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None  # the captured kernels must WRITE (beta = 0), not accumulate
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss = self.loss_of(model(**self.static))
            loss.backward()
        self.loss = loss.detach()
        self._grads = [p.grad for p in params]  # keep the captured gradient tensors alive and attached
        self._params = params
        # the graph holds raw addresses: parameters must not move afterwards (TorchAdamW re-points them into its flat
        # arena on its FIRST step — create the optimizer and take one step, or call optimizer._setup(), before capturing)
        self._addr = tuple(p.data_ptr() for p in params)

    @staticmethod
    def _first(out):
        while isinstance(out, (tuple, list)):
            out = out[0]
        return out

    def __call__(self, **inputs):
        for k, v in inputs.items():
            dst = self.static.get(k)
            if dst is None:
                raise KeyError("GraphedTrainStep: unknown input '%s' (captured: %s)" % (k, sorted(self.static)))
            if v.shape != dst.shape or v.dtype != dst.dtype:
                raise ValueError("GraphedTrainStep: input '%s' changed shape/dtype %s/%s -> %s/%s; capture is for "
                                 "fixed shapes" % (k, tuple(dst.shape), dst.dtype, tuple(v.shape), v.dtype))
            dst.copy_(v, non_blocking=True)
        if tuple(p.data_ptr() for p in self._params) != self._addr:
            raise RuntimeError("GraphedTrainStep: parameter storage moved after capture (an optimizer arena was "
                               "created later?); build the optimizer and run optimizer._setup() before capturing")
        for p, g in zip(self._params, self._grads):
            if p.grad is not g:
                p.grad = g  # e.g. after optimizer.zero_grad(set_to_none=True): the replay writes into these tensors
        self.graph.replay()
        return self.loss
