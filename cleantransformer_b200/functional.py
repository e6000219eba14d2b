"""Autograd glue: torch.autograd.Function wrappers around the C-ABI kernels (ops.py).

Numerics contract = the reference under torch.autocast(bfloat16): fp32 master parameters, fp32
residual stream / LayerNorm / softmax statistics / loss, bf16 (or fp16) tensor-core operands with
fp32 accumulation. Parameter gradients are fp32 and are written by the wgrad kernels DIRECTLY into
`param.grad` (or into a flat-arena view of it, see arena.py) instead of being returned to autograd:
that removes autograd's accumulate copies and lets the DDP wrapper see a gradient the moment the
kernel that produced it has been enqueued.
"""
import math
import os
import threading

import torch

from . import ops

COMPUTE_DTYPE = torch.bfloat16


def compute_dtype():
    """bf16 unless the caller is inside torch.autocast(device_type='cuda', dtype=float16)."""
    if torch.is_autocast_enabled():
        d = torch.get_autocast_dtype("cuda")
        if d in (torch.float16, torch.bfloat16):
            return d
    return COMPUTE_DTYPE


# ------------------------------------------------------------------------------------------------
# parameter-side bookkeeping
# ------------------------------------------------------------------------------------------------
def shadow(param, dtype=None):
    """Low-precision copy of an fp32 parameter for the tensor cores (autocast's weight cast),
    cached per parameter version. arena.py points `_ct_shadow_view` into a flat bf16 buffer that the
    fused AdamW kernel refreshes in the same pass as the fp32 update."""
    dtype = dtype or compute_dtype()
    if param.dtype == dtype:
        return param.detach()
    sh = getattr(param, "_ct_shadow", None)
    if sh is not None and sh.dtype == dtype and getattr(param, "_ct_shadow_ver", -1) == param._version \
            and getattr(param, "_ct_shadow_ptr", 0) == param.data_ptr() and sh.device == param.device:
        return sh
    target = getattr(param, "_ct_shadow_view", None)
    if target is not None and (target.dtype != dtype or target.device != param.device):
        target = None
    sh = ops.cast(param.detach(), dtype, out=target)
    param._ct_shadow = sh
    param._ct_shadow_ver = param._version
    param._ct_shadow_ptr = param.data_ptr()
    return sh


SKINNY_ROWS = 32  # csrc/gemm.cu: gemm_skinny_kernel serves M <= 32 with K-major weights


def shadow_kmajor(param, dtype=None):
    """[out, in] low-precision copy of a Conv1D weight stored [in, out] (modeling_gpt.py:32-46), cached per parameter
    version like `shadow`: the q_len = 1 decode step streams every weight once per token and the skinny kernel wants
    each output feature's row contiguous. Costs one extra bf16 copy of the Conv1D weights, made at the first decode."""
    dtype = dtype or compute_dtype()
    sh = getattr(param, "_ct_shadow_t", None)
    if sh is not None and sh.dtype == dtype and getattr(param, "_ct_shadow_t_ver", -1) == param._version \
            and getattr(param, "_ct_shadow_t_ptr", 0) == param.data_ptr() and sh.device == param.device:
        return sh
    sh = shadow(param, dtype).t().contiguous()
    param._ct_shadow_t, param._ct_shadow_t_ver, param._ct_shadow_t_ptr = sh, param._version, param.data_ptr()
    return sh


def invalidate_shadows(module_or_params):
    """Forget the cached low-precision copies. The cache is keyed on (`param._version`, `param.data_ptr()`): in-place
    updates through the parameter itself and `p.data = ...` re-pointing are noticed, but a write through `p.data`
    (`p.data.copy_()`, `p.data.mul_()`: weight surgery, EMA, a foreign optimizer) bumps neither — call this after one.
    The optimizers and the DDP wrapper of this package do it themselves."""
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    for p in params:
        p._ct_shadow_ver = -1


def grad_buffer(param):
    """(f32 tensor the gradient must be written into, accumulate?)."""
    g = param.grad
    if g is not None:
        return g, True
    view = getattr(param, "_ct_grad_view", None)
    if view is None or view.device != param.device:
        view = torch.empty(param.shape, dtype=torch.float32, device=param.device)
    param.grad = view
    return view, False


def note_use(*params):
    """Forward-side count of how many gradient writes a parameter will receive in the coming backward (one per
    use in a Function of this module). The DDP wrapper compares it with the writes it has seen to decide when a
    bucket is complete: a table looked up twice (GPT `segment_ids`, modeling_gpt.py:186-188), a shared module or a
    tied head all work without a declared count. (Called from inside Function.forward, where grad mode is always
    off — so no grad-mode test here; the wrapper resets the counters at the start of each synchronised forward.)"""
    for p in params:
        if p is not None and p.requires_grad:
            p._ct_uses = getattr(p, "_ct_uses", 0) + 1


def grad_written(param):
    """Tell listeners (the DDP wrapper) that a kernel producing this gradient has been enqueued."""
    hooks = getattr(param, "_ct_grad_hooks", None)
    if hooks:
        for h in hooks:
            h(param)


def _as2d(x):
    return x.reshape(-1, x.shape[-1])


def _anchor(params, *tensors):
    """Parameters never enter autograd as inputs of these Functions: their gradients are written in place
    (grad_buffer), so an edge to the parameter's AccumulateGrad node would carry nothing — but the engine still visits
    that node and, at the end of backward, synchronises the caller's stream with the stream the node was CREATED on.
    A node kept alive by an older graph (any loss tensor still referenced) then ties a CUDA-graph capture to
    uncaptured work on another stream (cudaErrorStreamCaptureIsolation). So parameters travel in a tuple, and when no
    tensor input requires grad (embedding of integer ids, first Linear on raw features) a fresh zero-size leaf —
    created on the current stream — makes autograd build the node."""
    if not torch.is_grad_enabled() or any(t is not None and t.requires_grad for t in tensors):
        return None
    for p in params:
        if p is not None and p.requires_grad:
            return torch.empty(0, device=p.device, requires_grad=True)
    return None


def _low(x, dtype):
    """Activation cast to the compute dtype (autocast's input cast)."""
    x = x.contiguous()
    return x if x.dtype == dtype else ops.cast(x, dtype)


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """transformer.py:79-89. Returns y in `out_dtype`, plus an optional second copy in `out2_dtype`
    (e.g. f32 for the residual stream + bf16 for the next GEMM)."""

    @staticmethod
    def forward(ctx, x, wb, eps, out_dtype, out2_dtype, anchor):
        weight, bias = wb
        need = x.requires_grad or weight.requires_grad or bias.requires_grad
        note_use(weight, bias)
        w = weight.detach().reshape(-1)
        b = bias.detach().reshape(-1)
        y, y2, mean, rstd = ops.layernorm_fwd(x.detach(), w, b, eps, out_dtype, out2_dtype, save_stats=need)
        ctx.save_for_backward(x, mean, rstd)
        ctx.weight, ctx.bias = weight, bias
        ctx.x_dtype = x.dtype
        if y2 is None:
            return y
        return y, y2

    @staticmethod
    def backward(ctx, dy, dy2=None):
        x, mean, rstd = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        gw, acc_w = grad_buffer(weight) if weight.requires_grad else (None, False)
        gb, acc_b = grad_buffer(bias) if bias.requires_grad else (None, False)
        if gw is not None and gb is not None and acc_w != acc_b:
            # keep one accumulate flag for the fused kernel: zero the fresh one
            (gw if not acc_w else gb).zero_()
            acc_w = acc_b = True
        dx = ops.layernorm_bwd(dy, x, weight.detach().reshape(-1), mean, rstd,
                               gw.view(-1) if gw is not None else None,
                               gb.view(-1) if gb is not None else None,
                               acc_w if gw is not None else acc_b, dy2=dy2,
                               dx_dtype=torch.float32 if ctx.x_dtype == torch.float32 else ctx.x_dtype)
        if gw is not None:
            grad_written(weight)
        if gb is not None:
            grad_written(bias)
        return dx, None, None, None, None, None


def layer_norm(x, weight, bias, eps, out_dtype=None, out2_dtype=None):
    return LayerNormFn.apply(x, (weight, bias), eps, out_dtype or x.dtype, out2_dtype, _anchor((weight, bias), x))


# ------------------------------------------------------------------------------------------------
# Linear / Conv1D (+ bias, activation, residual)
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = act(x @ W^T + b) (+ residual). W: nn.Linear [out,in] or Conv1D [in,out] (w_in_out)."""

    @staticmethod
    def forward(ctx, x, wb, act, residual, out_dtype, w_in_out, anchor, infer=False):
        """infer: linear() saw grad mode off (in here it always is): nothing is kept for a backward."""
        weight, bias = wb
        cd = compute_dtype()
        x2 = _low(_as2d(x.detach()), cd)
        N = weight.shape[1] if w_in_out else weight.shape[0]
        need = not infer and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad) or
                              (residual is not None and residual.requires_grad))
        if infer and w_in_out and x2.shape[0] <= SKINNY_ROWS and weight.shape[0] % 32 == 0:
            # q_len = 1 decode step: stream a K-major copy of the Conv1D weight (feature rows contiguous)
            w16, w_in_out_k = shadow_kmajor(weight, cd), False
        else:
            w16, w_in_out_k = shadow(weight, cd), w_in_out
        res2 = _as2d(residual.detach()).contiguous() if residual is not None else None
        note_use(weight, bias)
        y, pre = ops.linear_fwd(x2, w16, bias.detach() if bias is not None else None, act, res2, out_dtype,
                                save_preact=(act != ops.ACT_NONE and need), w_in_out=w_in_out_k)
        ctx.save_for_backward(x2, pre)
        ctx.weight, ctx.bias, ctx.act, ctx.w_in_out = weight, bias, act, w_in_out
        ctx.x_shape, ctx.x_dtype, ctx.x_req = x.shape, x.dtype, x.requires_grad
        ctx.has_res = residual is not None
        ctx.res_dtype = residual.dtype if residual is not None else None
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, pre = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        cd = x2.dtype
        dres = None
        if ctx.has_res:
            dres = dy if dy.dtype == ctx.res_dtype else ops.cast(dy.contiguous(), ctx.res_dtype)
        d2 = _as2d(dy).contiguous()
        if ctx.act != ops.ACT_NONE:
            d2 = ops.act_bwd(d2, pre, ctx.act, out_dtype=cd)
        else:
            d2 = _low(d2, cd)
        if weight.requires_grad:
            gw, acc = grad_buffer(weight)
            gbias = None
            if bias is not None and bias.requires_grad:
                gbias, accb = grad_buffer(bias)
                if accb != acc:
                    (gbias if not accb else gw).zero_()
                    acc = True
            ops.linear_wgrad(d2, x2, gw, gbias, accumulate=acc, w_in_out=ctx.w_in_out)
            grad_written(weight)
            if gbias is not None:
                grad_written(bias)
        elif bias is not None and bias.requires_grad:
            gbias, accb = grad_buffer(bias)
            ops.colsum(d2, gbias, accb)
            grad_written(bias)
        dx = None
        if ctx.x_req:
            dx = ops.linear_dgrad(d2, shadow(weight, cd), out_dtype=ctx.x_dtype, w_in_out=ctx.w_in_out)
            dx = dx.view(ctx.x_shape)
        return dx, None, None, dres, None, None, None, None


def linear(x, weight, bias=None, act=ops.ACT_NONE, residual=None, out_dtype=None, w_in_out=False):
    if out_dtype is None:
        out_dtype = residual.dtype if residual is not None else compute_dtype()
    return LinearFn.apply(x, (weight, bias), act, residual, out_dtype, w_in_out, _anchor((weight, bias), x, residual),
                          not torch.is_grad_enabled())


# ------------------------------------------------------------------------------------------------
# Attention core
# ------------------------------------------------------------------------------------------------
# Dropout: one counter-based generator for every site (include/ct_b200.h: "dropout")
# ------------------------------------------------------------------------------------------------
class _DropoutState:
    """seed: torch's CPU generator seed at first use (so `torch.manual_seed` governs the masks as it does for the
    reference), or `manual_dropout_seed`. Every dropout call of a forward pass takes the next `stream` number; the
    backward of that call reuses it, so no mask is ever stored."""
    seed = None
    stream = 0


def manual_dropout_seed(seed):
    _DropoutState.seed = int(seed)
    _DropoutState.stream = 0


def next_dropout(p):
    """(p, seed, stream) for one dropout site call; refuses CUDA-graph capture (the stream number is a host counter that
    a replay would not advance: every replay would drop the same elements)."""
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        raise RuntimeError("dropout > 0 cannot be captured in a CUDA graph (host-side mask counter); run the step un-graphed")
    if _DropoutState.seed is None:
        _DropoutState.seed = int(torch.initial_seed())
    _DropoutState.stream = (_DropoutState.stream + 1) & 0xFFFFFFFF
    return (float(p), _DropoutState.seed, _DropoutState.stream)


class DropoutFn(torch.autograd.Function):
    """torch.nn.Dropout on a hidden-state tensor, fused with the residual add that follows it where there is one
    (transformer.py:109-116, modeling_gpt.py:136,150-153, modeling_bert.py:253-262, modeling_bloom.py dropout_add)."""

    @staticmethod
    def forward(ctx, x, residual, drop, out_dtype):
        ctx.drop = drop
        ctx.x_dtype = x.dtype
        ctx.has_res = residual is not None
        ctx.res_dtype = residual.dtype if residual is not None else None
        return ops.dropout(x.detach(), drop[0], drop[1], drop[2], residual.detach() if residual is not None else None,
                           out_dtype)

    @staticmethod
    def backward(ctx, dy):
        p, seed, stream = ctx.drop
        dx = ops.dropout(dy, p, seed, stream, None, ctx.x_dtype)
        dres = None
        if ctx.has_res:
            dres = dy if dy.dtype == ctx.res_dtype else ops.cast(dy.contiguous(), ctx.res_dtype)
        return dx, dres, None, None


def dropout(x, p, training, residual=None, out_dtype=None):
    """residual + dropout(x) (or dropout(x)); identity (+ residual) when not training or p == 0 — the residual add then
    runs through the same kernel with p = 0 (no ATen add on the path: transformer.py:109-111 has no GEMM to carry it)."""
    od = out_dtype or (residual.dtype if residual is not None else x.dtype)
    if not training or p <= 0:
        if residual is None:
            return x if x.dtype == od else x.to(od)
        return DropoutFn.apply(x, residual, (0.0, 0, 0), od)
    return DropoutFn.apply(x, residual, next_dropout(p), od)


LAYOUT_BLOOM = "bloom"        # fused [B,S,H,3,D]   (modeling_bloom.py:81-82)
LAYOUT_GPT = "gpt"            # fused [B,S,3,H,D]   (modeling_gpt.py:72)


def split_packed(qkv, n_head, layout):
    """Strided [B,H,S,D] views of q, k, v inside a packed projection output."""
    B, S, three_h = qkv.shape
    D = three_h // (3 * n_head)
    if layout == LAYOUT_BLOOM:
        t = qkv.view(B, S, n_head, 3, D)
        return [t[:, :, :, i, :].permute(0, 2, 1, 3) for i in range(3)]
    t = qkv.view(B, S, 3, n_head, D)
    return [t[:, :, i, :, :].permute(0, 2, 1, 3) for i in range(3)]


class PackedAttentionFn(torch.autograd.Function):
    """softmax(QK^T*scale + bias/masks) V on a packed QKV tensor (training / prefill, no cache)."""

    @staticmethod
    def forward(ctx, qkv, n_head, layout, scale, causal, causal_fill, kbias2, first_valid, drop=None):
        """drop: None or next_dropout(p) — attention-probability dropout inside the kernel."""
        qkv_d = qkv.detach().contiguous()
        q, k, v = split_packed(qkv_d, n_head, layout)
        o, lse2 = ops.attn_fwd(q, k, v, scale, causal, causal_fill, kbias2, first_valid,
                               need_lse=qkv.requires_grad, dropout=drop)
        ctx.save_for_backward(qkv_d, o, lse2, kbias2, first_valid)
        ctx.cfg = (n_head, layout, scale, causal, causal_fill, drop)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse2, kbias2, first_valid = ctx.saved_tensors
        n_head, layout, scale, causal, causal_fill, drop = ctx.cfg
        dqkv = torch.empty_like(qkv)
        q, k, v = split_packed(qkv, n_head, layout)
        dq, dk, dv = split_packed(dqkv, n_head, layout)
        do = do.contiguous()
        if do.dtype != qkv.dtype:
            do = ops.cast(do, qkv.dtype)
        ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, causal, causal_fill, kbias2, first_valid, dropout=drop)
        return dqkv, None, None, None, None, None, None, None, None


class SeparateAttentionFn(torch.autograd.Function):
    """Same, for three separate [B,S,H*D] projections (transformer.py:37-57)."""

    @staticmethod
    def forward(ctx, q, k, v, n_head, scale, causal, causal_fill, kbias2, first_valid, drop=None):
        B, Sq, HD = q.shape
        D = HD // n_head
        qd, kd, vd = [t.detach().contiguous() for t in (q, k, v)]
        q4, k4, v4 = [t.view(B, t.shape[1], n_head, D).permute(0, 2, 1, 3) for t in (qd, kd, vd)]
        need = q.requires_grad or k.requires_grad or v.requires_grad
        o, lse2 = ops.attn_fwd(q4, k4, v4, scale, causal, causal_fill, kbias2, first_valid, need_lse=need, dropout=drop)
        ctx.save_for_backward(qd, kd, vd, o, lse2, kbias2, first_valid)
        ctx.cfg = (n_head, scale, causal, causal_fill, drop)
        return o

    @staticmethod
    def backward(ctx, do):
        qd, kd, vd, o, lse2, kbias2, first_valid = ctx.saved_tensors
        n_head, scale, causal, causal_fill, drop = ctx.cfg
        B, Sq, HD = qd.shape
        D = HD // n_head
        v4 = lambda t: t.view(B, t.shape[1], n_head, D).permute(0, 2, 1, 3)
        dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
        do = do.contiguous()
        if do.dtype != qd.dtype:
            do = ops.cast(do, qd.dtype)
        ops.attn_bwd(do, v4(qd), v4(kd), v4(vd), o, lse2, v4(dq), v4(dk), v4(dv), scale, causal,
                     causal_fill, kbias2, first_valid, dropout=drop)
        return dq, dk, dv, None, None, None, None, None, None, None


def attention_cached(q4, k4, v4, scale, causal, causal_fill, kbias2, first_valid, drop=None):
    """Inference with a KV cache: q4 [B,H,Sq,D], k4/v4 [B,H,Sk,D] (any strides). No autograd. `drop`: the reference's
    generate() on a model left in train mode applies its attention dropout here as well."""
    o, _ = ops.attn_fwd(q4, k4, v4, scale, causal, causal_fill, kbias2, first_valid, need_lse=False, dropout=drop)
    return o


# ------------------------------------------------------------------------------------------------
# Embedding and LM loss
# ------------------------------------------------------------------------------------------------
class EmbeddingFn(torch.autograd.Function):
    """sum_k weight_k[ids_k] (token + position + segment tables): modeling_bloom.py:190,
    modeling_gpt.py:169-188, modeling_bert.py:297-300. Gradients are scattered straight into
    weight.grad."""

    @staticmethod
    def forward(ctx, ids, weights, padding_idx0, anchor):
        shape = torch.broadcast_shapes(*[i.shape for i in ids])
        out = None
        for i, w in zip(ids, weights):
            i = i.expand(shape).contiguous()
            out = ops.embedding_fwd(i, w.detach(), out, accumulate=out is not None)
        ctx.ids = [i.expand(shape).contiguous() for i in ids]
        ctx.weights = weights
        note_use(*weights)
        ctx.padding_idx0 = padding_idx0
        return out

    @staticmethod
    def backward(ctx, dout):
        _embedding_scatter(ctx.ids, ctx.weights, dout, ctx.padding_idx0)
        return None, None, None, None


def _embedding_scatter(ids, weights, dout, padding_idx0):
    """The autograd scatter of the gathers: every table's gradient += its token rows of `dout`."""
    dout = dout.contiguous()
    if dout.dtype != torch.float32:
        dout = ops.cast(dout, torch.float32)
    for k, (i, w) in enumerate(zip(ids, weights)):
        if not w.requires_grad:
            continue
        g, acc = grad_buffer(w)
        if not acc:
            g.zero_()
        pad = padding_idx0 if k == 0 else -1
        override = getattr(w, "_ct_embedding_bwd_override", None)  # ddp.py: sparse exchange of a tied table
        if override is not None:
            override(w, i, dout, g, pad)
        else:
            ops.embedding_bwd(i, dout, g, pad)
        grad_written(w)


def embedding_sum(ids_list, weight_list, padding_idx0=-1):
    return EmbeddingFn.apply(tuple(ids_list), tuple(weight_list), padding_idx0, _anchor(weight_list))


# SURVEY §8 f N3: the gather(s) and the LayerNorm that follows them as ONE kernel (Bloom's word_embeddings_layernorm,
# BERT's embedding_post). Opt-in (CT_FUSED_EMBED_LN=1) — parity-tested on the GPU, never benchmarked at length.
FUSED_EMBED_LN = os.environ.get("CT_FUSED_EMBED_LN", "0") == "1"


class EmbeddingLNFn(torch.autograd.Function):
    """LayerNorm(sum_k weight_k[ids_k]): modeling_bloom.py:190-191, modeling_bert.py:297-301. The backward is the
    LayerNorm backward (transformer.py:79-89) on the saved sum followed by the scatter of EmbeddingFn."""

    @staticmethod
    def forward(ctx, ids, weights, padding_idx0, ln_wb, eps, out_dtype, anchor):
        gamma, beta = ln_wb
        shape = torch.broadcast_shapes(*[i.shape for i in ids])
        ids = [i.expand(shape).contiguous() for i in ids]
        need = any(w.requires_grad for w in weights) or gamma.requires_grad or beta.requires_grad
        note_use(*weights)
        note_use(gamma, beta)
        emb, y, _, mean, rstd = ops.embedding_layernorm_fwd(ids, [w.detach() for w in weights], gamma.detach().reshape(-1),
                                                            beta.detach().reshape(-1), eps, out_dtype, None, save=need)
        if need:
            ctx.save_for_backward(emb, mean, rstd)
        ctx.ids, ctx.weights, ctx.padding_idx0 = ids, weights, padding_idx0
        ctx.gamma, ctx.beta = gamma, beta
        return y

    @staticmethod
    def backward(ctx, dy):
        emb, mean, rstd = ctx.saved_tensors
        gamma, beta = ctx.gamma, ctx.beta
        gw, acc_w = grad_buffer(gamma) if gamma.requires_grad else (None, False)
        gb, acc_b = grad_buffer(beta) if beta.requires_grad else (None, False)
        if gw is not None and gb is not None and acc_w != acc_b:
            (gw if not acc_w else gb).zero_()
            acc_w = acc_b = True
        demb = ops.layernorm_bwd(dy, emb, gamma.detach().reshape(-1), mean, rstd,
                                 gw.view(-1) if gw is not None else None, gb.view(-1) if gb is not None else None,
                                 acc_w if gw is not None else acc_b, dx_dtype=torch.float32)
        if gw is not None:
            grad_written(gamma)
        if gb is not None:
            grad_written(beta)
        _embedding_scatter(ctx.ids, ctx.weights, demb, ctx.padding_idx0)
        return None, None, None, None, None, None, None


def embedding_layer_norm(ids_list, weight_list, ln_weight, ln_bias, eps, padding_idx0=-1, out_dtype=torch.float32):
    """LayerNorm(embedding_sum(...)): one kernel when FUSED_EMBED_LN is on and the shape fits, else the two calls."""
    if FUSED_EMBED_LN and ops.embedding_layernorm_ok(weight_list, ln_weight):
        return EmbeddingLNFn.apply(tuple(ids_list), tuple(weight_list), padding_idx0, (ln_weight, ln_bias), eps,
                                   out_dtype, _anchor(tuple(weight_list) + (ln_weight, ln_bias)))
    return layer_norm(embedding_sum(ids_list, weight_list, padding_idx0), ln_weight, ln_bias, eps, out_dtype=out_dtype)


class LMLossFn(torch.autograd.Function):
    """modeling_bloom.py:224-230: shifted cross entropy, mean over B*(S-1) positions. One pass over
    the logits produces the loss and dlogits."""

    @staticmethod
    def forward(ctx, logits, labels, shift):
        B, S, V = logits.shape
        l2 = logits.detach().reshape(B * S, V)
        loss, dl = ops.cross_entropy_fwd(l2, labels.reshape(-1), S=S, shift=shift,
                                         want_dlogits=logits.requires_grad)
        ctx.dl = dl
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        dl = ctx.dl
        ctx.dl = None
        ops.scale_by_scalar(dl, dloss.contiguous().float())
        return dl.view(ctx.shape), None, None


def lm_loss(logits, labels, shift=True):
    return LMLossFn.apply(logits, labels, shift)


FUSED_LM_STATS = os.environ.get("CT_FUSED_LM_STATS", "0") != "0"


class LMHeadLossFn(torch.autograd.Function):
    """lm_head + shifted cross entropy as one node (modeling_bloom.py:220-230): the logits GEMM leaves the softmax
    statistics of every row next to the logits, so the loss kernel is a single streaming pass (SURVEY §8 f, N1, first
    half). Returns (loss, logits); the logits are an output for the caller's tuple only (non-differentiable here —
    the node back-propagates the loss). Opt-in (CT_FUSED_LM_STATS=1): not yet run on a GPU."""

    @staticmethod
    def forward(ctx, hidden, wb, labels, shift, anchor):
        weight = wb[0]
        cd = compute_dtype()
        B, S, H = hidden.shape
        x2 = _low(_as2d(hidden.detach()), cd)
        w16 = shadow(weight, cd)
        logits, stats = ops.lm_head_logits_with_stats(x2, w16)
        need = hidden.requires_grad or weight.requires_grad
        note_use(weight)
        loss, dl = ops.cross_entropy_fwd_stats(logits, labels.reshape(-1), stats, S=S, shift=shift, want_dlogits=need)
        ctx.save_for_backward(x2, dl)
        ctx.weight = weight
        ctx.x_shape, ctx.x_dtype, ctx.x_req = hidden.shape, hidden.dtype, hidden.requires_grad
        logits = logits.view(B, S, -1)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        x2, dl = ctx.saved_tensors
        weight = ctx.weight
        ops.scale_by_scalar(dl, dloss.contiguous().float())
        if weight.requires_grad:
            gw, acc = grad_buffer(weight)
            ops.linear_wgrad(dl, x2, gw, None, accumulate=acc)
            grad_written(weight)
        dx = None
        if ctx.x_req:
            dx = ops.linear_dgrad(dl, shadow(weight, x2.dtype), out_dtype=ctx.x_dtype).view(ctx.x_shape)
        return dx, None, None, None, None


def lm_head_loss(hidden, weight, labels, shift=True):
    """(loss, logits) of the tied LM head; the fused-statistics node when it is enabled and the shape allows it."""
    B, S, H = hidden.shape
    if FUSED_LM_STATS and ops.lm_head_stats_ok(B * S, weight.shape[0], compute_dtype()):
        return LMHeadLossFn.apply(hidden, (weight,), labels, shift, _anchor((weight,), hidden))
    logits = linear(hidden, weight)
    return lm_loss(logits, labels, shift=shift), logits


def reject_head_mask(head_mask):
    """The reference multiplies the softmax weights by `head_mask` when one is passed (transformer.py:48-50,
    modeling_bloom.py:112-113, modeling_gpt.py:95-96); the fused kernels never materialise those weights. Refuse
    instead of silently computing something else (same policy as dropout with p > 0)."""
    if head_mask is not None:
        raise NotImplementedError("head_mask is not supported by the fused attention kernels (the softmax weights are "
                                  "never materialised); pass head_mask=None")


def default_scale(head_dim):
    return 1.0 / math.sqrt(head_dim)


# ------------------------------------------------------------------------------------------------
# Fused pre-LN block (Bloom / GPT-2 wiring): one autograd node per block
# ------------------------------------------------------------------------------------------------
SAVE_ACT_GRAD = os.environ.get("CT_SAVE_ACT_GRAD", "0") != "0"


class PreLNBlockFn(torch.autograd.Function):
    """x -> x + proj(attn(qkv(LN1 x)))  -> ... + W2 act(W1 LN2(.)):
    modeling_bloom.py:142-159 (apply_residual_connection_post_layernorm = False) and the 'gpt2' branch
    of modeling_gpt.py:147-152. Compared with composing the op-level Functions this removes, per layer,
    the stand-alone activation-gradient kernel (fused into the 4h->h dgrad epilogue), the two residual
    gradient adds (fused into the LayerNorm backward via dx_add) and ~10 autograd nodes.

    `spec` = dict(ln1, qkv, proj, ln2, fc1, fc2: modules holding .weight/.bias; w_in_out: Conv1D
    layout flag; n_head, layout, scale, causal, causal_fill, act). Parameter gradients are written in
    place (functional.grad_buffer), only the hidden-state gradient flows through autograd."""

    @staticmethod
    def forward(ctx, x, spec, kbias2, first_valid):
        cd = compute_dtype()
        io = spec["w_in_out"]
        B, S, H = x.shape
        xd = x.detach().contiguous()
        x2 = xd.view(B * S, H)
        ln1, _, mean1, rstd1 = ops.layernorm_fwd(x2, spec["ln1"].weight.detach(), spec["ln1"].bias.detach(),
                                                 spec["ln1"].eps, out_dtype=cd)
        qkv, _ = ops.linear_fwd(ln1, shadow(spec["qkv"].weight, cd), spec["qkv"].bias.detach(), out_dtype=cd,
                                w_in_out=io)
        q, k, v = split_packed(qkv.view(B, S, -1), spec["n_head"], spec["layout"])
        o, lse2 = ops.attn_fwd(q, k, v, spec["scale"], spec["causal"], spec["causal_fill"], kbias2, first_valid)
        att, _ = ops.linear_fwd(o.view(B * S, H), shadow(spec["proj"].weight, cd), spec["proj"].bias.detach(),
                                residual=x2, out_dtype=torch.float32, w_in_out=io)
        ln2, _, mean2, rstd2 = ops.layernorm_fwd(att, spec["ln2"].weight.detach(), spec["ln2"].bias.detach(),
                                                 spec["ln2"].eps, out_dtype=cd)
        # tanh-GELU: the forward epilogue saves gelu'(h) (same tanh as gelu(h)) in place of h, so the 4h->h dgrad
        # epilogue only multiplies (SAVE_ACT_GRAD; `pre` then holds the derivative, see backward)
        act_f = ops.ACT_GELU_TANH_SAVE_GRAD if (SAVE_ACT_GRAD and spec["act"] == ops.ACT_GELU_TANH) else spec["act"]
        h4, pre = ops.linear_fwd(ln2, shadow(spec["fc1"].weight, cd), spec["fc1"].bias.detach(), act=act_f,
                                 out_dtype=cd, save_preact=True, w_in_out=io)
        ctx.pre_is_grad = act_f == ops.ACT_GELU_TANH_SAVE_GRAD
        out, _ = ops.linear_fwd(h4, shadow(spec["fc2"].weight, cd), spec["fc2"].bias.detach(), residual=att,
                                out_dtype=torch.float32, w_in_out=io)
        ctx.save_for_backward(x2, mean1, rstd1, ln1, qkv, o, lse2, att, mean2, rstd2, ln2, pre, h4, kbias2,
                              first_valid)
        ctx.spec = spec
        ctx.shape = (B, S, H)
        for name in ("ln1", "qkv", "proj", "ln2", "fc1", "fc2"):
            note_use(spec[name].weight, spec[name].bias)
        ctx.mark_non_differentiable(k, v)
        ctx.set_materialize_grads(False)  # no zero-filled [B,H,S,D] gradients for the k/v outputs
        return out.view(B, S, H), k, v

    @staticmethod
    def _linear_bwd(lin, dy16, x16, io, cd, need_dx=True, actgrad_src=None, actgrad_act=ops.ACT_NONE,
                    bias_done=False):
        """wgrad (+ bias gradient unless `bias_done`: the LayerNorm backward that produced dy already
        wrote its column sums into bias.grad) and dgrad of one Linear / Conv1D."""
        w, b = lin.weight, lin.bias
        if w.requires_grad:
            gw, acc = grad_buffer(w)
            gb = None
            if b is not None and b.requires_grad and not bias_done:
                gb, accb = grad_buffer(b)
                if accb != acc:
                    (gb if not accb else gw).zero_()
                    acc = True
            ops.linear_wgrad(dy16, x16, gw, gb, accumulate=acc, w_in_out=io)
            grad_written(w)
            if gb is not None:
                grad_written(b)
        if not need_dx:
            return None
        return ops.linear_dgrad(dy16, shadow(w, cd), out_dtype=cd, actgrad_src=actgrad_src,
                                actgrad_act=actgrad_act, w_in_out=io)

    @staticmethod
    def _ln_bwd(ln, dy16, x, mean, rstd, dx_add, low_dtype=None, bias_of=None):
        """LayerNorm backward + residual-path add. Returns (dx f32, dx in `low_dtype` or None). With
        `bias_of` (the Linear whose `residual + Linear(.)` output is this LayerNorm's input) the same
        pass also writes that Linear's bias gradient = column sums of dx."""
        gw, acc_w = grad_buffer(ln.weight) if ln.weight.requires_grad else (None, False)
        gb, acc_b = grad_buffer(ln.bias) if ln.bias.requires_grad else (None, False)
        if gw is not None and gb is not None and acc_w != acc_b:
            (gw if not acc_w else gb).zero_()
            acc_w = acc_b = True
        dxsum, acc_s = None, False
        if bias_of is not None and bias_of.bias is not None and bias_of.bias.requires_grad:
            dxsum, acc_s = grad_buffer(bias_of.bias)
        out = ops.layernorm_bwd(dy16, x, ln.weight.detach(), mean, rstd, gw, gb, acc_w if gw is not None else acc_b,
                                dx_add=dx_add, dx_dtype=torch.float32, dx2_dtype=low_dtype,
                                dxsum=dxsum, dxsum_accumulate=acc_s)
        if gw is not None:
            grad_written(ln.weight)
        if gb is not None:
            grad_written(ln.bias)
        if dxsum is not None:
            grad_written(bias_of.bias)
        return out if low_dtype is not None else (out, None)

    @staticmethod
    def backward(ctx, g_out, _gk, _gv):
        (x2, mean1, rstd1, ln1, qkv, o, lse2, att, mean2, rstd2, ln2, pre, h4, kbias2, first_valid) = ctx.saved_tensors
        spec = ctx.spec
        B, S, H = ctx.shape
        io = spec["w_in_out"]
        cd = ln1.dtype
        # the block above (later in forward order) hands over the low-precision copy of the very tensor
        # autograd passes us, written by its LayerNorm backward in the same pass as the f32 gradient
        g16 = _take_handoff(g_out, cd)
        g_out = g_out.contiguous().view(B * S, H)
        if g_out.dtype != torch.float32:
            g_out = ops.cast(g_out, torch.float32)
        if g16 is None:
            g16 = ops.cast(g_out, cd)
        else:
            g16 = g16.view(B * S, H)
        # FFN: out = att + fc2(act(fc1(ln2)))
        d_pre = PreLNBlockFn._linear_bwd(spec["fc2"], g16, h4, io, cd, actgrad_src=pre,
                                         actgrad_act=ops.ACT_GRAD_PRECOMPUTED if ctx.pre_is_grad else spec["act"])
        d_ln2 = PreLNBlockFn._linear_bwd(spec["fc1"], d_pre, ln2, io, cd)
        # LN2 backward + residual path; also emits bf16(g_att) and proj.bias.grad = colsum(g_att)
        g_att, g16b = PreLNBlockFn._ln_bwd(spec["ln2"], d_ln2, att, mean2, rstd2, g_out, low_dtype=cd,
                                           bias_of=spec["proj"])
        # attention: att = x + proj(attn(qkv(ln1)))
        d_o = PreLNBlockFn._linear_bwd(spec["proj"], g16b, o.view(B * S, H), io, cd, bias_done=True)
        dqkv = torch.empty_like(qkv)
        q, k, v = split_packed(qkv.view(B, S, -1), spec["n_head"], spec["layout"])
        dq, dk, dv = split_packed(dqkv.view(B, S, -1), spec["n_head"], spec["layout"])
        ops.attn_bwd(d_o.view(B, S, H), q, k, v, o, lse2, dq, dk, dv, spec["scale"], spec["causal"],
                     spec["causal_fill"], kbias2, first_valid)
        d_ln1 = PreLNBlockFn._linear_bwd(spec["qkv"], dqkv, ln1, io, cd)
        g_x, g_x16 = PreLNBlockFn._ln_bwd(spec["ln1"], d_ln1, x2, mean1, rstd1, g_att, low_dtype=cd)
        g_x = g_x.view(B, S, H)
        _put_handoff(g_x, g_x16)
        return g_x, None, None, None


# One-slot hand-over of a gradient's low-precision copy between consecutive PreLNBlockFn backward
# nodes. The consumer only uses it when autograd passes the IDENTICAL tensor object on (no
# accumulation with another consumer's gradient, no hook rewrote it); holding the f32 tensor here
# keeps its storage alive, so the identity test cannot be fooled by a recycled allocation.
_HANDOFF = threading.local()


def _put_handoff(g32, g16):
    _HANDOFF.slot = (g32, g32._version, g16)


def _take_handoff(g, dtype):
    slot = getattr(_HANDOFF, "slot", None)
    _HANDOFF.slot = None
    if slot is None:
        return None
    g32, ver, g16 = slot
    if g32 is g and g._version == ver and g16 is not None and g16.dtype == dtype:
        return g16
    return None
