"""Plain-torch stand-ins for `cleantransformer_b200.ops` (TEST INFRASTRUCTURE, CPU only).

The host layer above the C ABI — autograd Functions, residual wiring, gradient bookkeeping, tied weights, the DDP
hooks — is ordinary Python that only ever touches tensors through `ops.*`. Patching those entry points with
reference arithmetic lets the `-m "not gpu"` suite execute that Python (forward AND backward) on the build
machine and compare it with the golden vectors of the real reference. Nothing here ships: the product still
refuses CPU tensors (`ops._req_cuda`), and these mocks follow the documented contract of each wrapper in ops.py /
include/ct_b200.h, not the kernels' internals.
"""
import contextlib
import math

import torch

LOG2E = 1.4426950408889634
FLT_MAX = 3.4028234663852886e38


def _act(x, act):
    if act in (3, 5):
        return x * 0.5 * (1 + torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x)))
    if act == 1:
        return torch.relu(x)
    if act == 2:
        return 0.5 * x * (1 + torch.erf(x * 0.70710678118654752))
    if act == 4:
        return torch.tanh(x)
    return x


def _act_grad(x, act):
    if act == 6:
        return x
    if act in (3, 5):
        t = torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x))
        return 0.5 * x * ((1 - t * t) * (0.79788456 + 0.1070322243 * x * x)) + 0.5 * (1 + t)
    if act == 1:
        return (x > 0).to(x.dtype)
    if act == 2:
        return 0.5 * (1 + torch.erf(x * 0.70710678118654752)) + x * 0.3989422804014327 * torch.exp(-0.5 * x * x)
    if act == 4:
        return 1 - torch.tanh(x) ** 2
    return torch.ones_like(x)


def layernorm_fwd(x, gamma, beta, eps, out_dtype=None, out2_dtype=None, save_stats=True):
    cols = gamma.numel()
    x2 = x.contiguous().view(-1, cols).float()
    mean = x2.mean(-1)
    var = ((x2 - mean[:, None]) ** 2).mean(-1)
    rstd = 1.0 / torch.sqrt(var + eps)
    y32 = ((x2 - mean[:, None]) * rstd[:, None] * gamma + beta).view(x.shape)
    y = y32.to(out_dtype or x.dtype) if out_dtype is not False else None
    y2 = y32.to(out2_dtype) if out2_dtype is not None else None
    return y, y2, (mean if save_stats else None), (rstd if save_stats else None)


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, accumulate, dy2=None, dx_add=None,
                  dx_dtype=torch.float32, dx2_dtype=None, dxsum=None, dxsum_accumulate=False):
    cols = gamma.numel()
    x2 = x.contiguous().view(-1, cols).float()
    g = torch.zeros_like(x2)
    if dy is not None:
        g = g + dy.reshape(-1, cols).float()
    if dy2 is not None:
        g = g + dy2.reshape(-1, cols).float()
    xhat = (x2 - mean[:, None]) * rstd[:, None]
    dxh = g * gamma
    dx = rstd[:, None] * (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True))
    if dx_add is not None:
        dx = dx + dx_add.reshape(-1, cols).float()
    for buf, val, acc in ((dgamma, (g * xhat).sum(0), accumulate), (dbeta, g.sum(0), accumulate),
                          (dxsum, dx.sum(0), dxsum_accumulate)):
        if buf is not None:
            if acc:
                buf.add_(val)
            else:
                buf.copy_(val)
    out = dx.view(x.shape).to(dx_dtype)
    return out if dx2_dtype is None else (out, dx.view(x.shape).to(dx2_dtype))


def cast(src, dtype, out=None):
    if out is None:
        return src.to(dtype)
    out.copy_(src)
    return out


def colsum(x2d, out, accumulate):
    s = x2d.float().sum(0)
    if accumulate:
        out.add_(s)
    else:
        out.copy_(s)


def act_fwd(x, act, out_dtype=None):
    return _act(x.float(), act).to(out_dtype or x.dtype)


def act_bwd(dy, x, act, out_dtype=None):
    return (dy.float() * _act_grad(x.float(), act)).to(out_dtype or dy.dtype)


def gemm(A, B, M, N, K, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, alpha=1.0, beta=0.0, bias=None,
         act=0, preact=None, actgrad_src=None, actgrad_act=0, residual=None, impl=0, row_stats=None):
    a = (A.t() if a_mn else A).float()[:M, :K]
    b = (B.t() if b_mn else B).float()[:N, :K]
    t = alpha * (a @ b.t())
    if bias is not None:
        t = t + bias.float()
    if preact is not None:
        preact.copy_(_act_grad(t, 3) if act == 5 else t)
    t = _act(t, act)
    if actgrad_src is not None:
        t = t * _act_grad(actgrad_src.float(), actgrad_act)
    if residual is not None:
        t = t + residual.float()
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype)
    if beta != 0.0:
        t = t + beta * out.float()
    out.copy_(t)
    if row_stats is not None:  # [slots, M, 2]: (max2, sum2) of the stored values per (256-column tile, chunk parity)
        v = out.float() * LOG2E
        row_stats[..., 0] = float("-inf"); row_stats[..., 1] = 0.0
        for tile in range((N + 255) // 256):
            for hf in range(2):
                cols = [c for ch in range(hf, 8, 2) for c in range(tile * 256 + ch * 32, tile * 256 + ch * 32 + 32) if c < N]
                if cols:
                    sub = v[:, cols]
                    mx = sub.max(-1).values
                    row_stats[2 * tile + hf, :, 0] = mx
                    row_stats[2 * tile + hf, :, 1] = torch.exp2(sub - mx[:, None]).sum(-1)
    return out


def attn_mask_prep(attention_mask, n_head, mode, slopes=None):
    m = attention_mask.float()
    B, Sk = m.shape
    if mode == 0:
        pos = (torch.cumsum(m, -1) - 1) * m
        kb = slopes.float()[None, :, None] * pos[:, None, :] * LOG2E
        kb = torch.where(m[:, None, :] == 1, kb, torch.tensor(float("-inf")))
    elif mode == 1:
        kb = ((1 - m) * torch.finfo(torch.float32).min * LOG2E)[:, None, :]
    else:
        kb = ((1 - m) * -10000.0 * LOG2E)[:, None, :]
    fv = torch.where(m.sum(-1) > 0, m.argmax(-1), torch.full((B,), Sk)).to(torch.int32)
    return kb.contiguous(), fv


def _attn_core(q, k, v, scale, causal, causal_fill, kbias2, dropout=None):
    Sq, Sk = q.shape[2], k.shape[2]
    s2 = (q.float() @ k.float().transpose(2, 3)) * (scale * LOG2E)
    kb = kbias2[:, :, None, :] if kbias2 is not None else 0.0
    s2 = s2 + kb
    if causal:
        i = torch.arange(Sq)[:, None]
        j = torch.arange(Sk)[None, :]
        fill = torch.full_like(s2, causal_fill * LOG2E if causal_fill > -1e30 else float("-inf")) + kb
        s2 = torch.where(j > i + (Sk - Sq), fill.detach(), s2)
    s2 = s2.clamp_min(-FLT_MAX)
    mx = s2.max(-1, keepdim=True).values.detach()
    e = torch.exp2(s2 - mx)
    l = e.sum(-1, keepdim=True)
    pr = e / l
    if dropout is not None and dropout[0] > 0:  # after the softmax, survivors scaled (include/ct_b200.h: "dropout")
        from oracle import ct_oracle as O
        keep = O.dropout_mask_attention(q.shape[0], q.shape[1], Sq, Sk, *dropout)
        pr = pr * keep / (1.0 - dropout[0])
    o = pr @ v.float()
    return o, (mx + torch.log2(l)).squeeze(-1)


def attn_fwd(q, k, v, scale, causal=False, causal_fill=-FLT_MAX, kbias2=None, first_valid=None, need_lse=True, impl=0,
             seq_len_dev=None, dropout=None, kv_new=None):
    B, H, Sq, D = q.shape
    if seq_len_dev is not None:  # captured decode step: the key count is a device scalar, k / v are the whole capacity
        n = int(seq_len_dev[0])
        assert Sq == 1 and n <= k.shape[2]
        if kv_new is not None:  # the kernel stores the new token's rows at n - 1 before attending
            k[:, :, n - 1:n] = kv_new[0]
            v[:, :, n - 1:n] = kv_new[1]
        k, v = k[:, :, :n], v[:, :, :n]
        kbias2 = kbias2[:, :, :n] if kbias2 is not None else None
    o, lse2 = _attn_core(q, k, v, scale, causal, causal_fill, kbias2, dropout)
    return o.transpose(1, 2).reshape(B, Sq, H * D).to(q.dtype), (lse2 if need_lse else None)


def attn_bwd(dout, q, k, v, o, lse2, dq, dk, dv, scale, causal=False, causal_fill=-FLT_MAX, kbias2=None,
             first_valid=None, impl=0, dropout=None):
    B, H, Sq, D = q.shape
    with torch.enable_grad():
        qr, kr, vr = [t.detach().float().clone().requires_grad_(True) for t in (q, k, v)]
        out, _ = _attn_core(qr, kr, vr, scale, causal, causal_fill, kbias2, dropout)
        out.backward(dout.float().view(B, Sq, H, D).transpose(1, 2))
    dq.copy_(qr.grad); dk.copy_(kr.grad); dv.copy_(vr.grad)


def dropout(x, p, seed, rng_stream, residual=None, out_dtype=None):
    from oracle import ct_oracle as O
    keep = O.dropout_mask_elementwise(x.shape, p, seed, rng_stream)
    y = x.float() * keep / (1.0 - p)
    if residual is not None:
        y = y + residual.float()
    return y.to(out_dtype or x.dtype)


def embedding_fwd(ids, weight, out=None, accumulate=False):
    e = weight[ids]
    if out is None:
        return e.float()
    if accumulate:
        out.add_(e)
    else:
        out.copy_(e)
    return out


def embedding_layernorm_fwd(ids_list, weights, gamma, beta, eps, out_dtype=torch.float32, out2_dtype=None, save=True):
    emb = None
    for i, w in zip(ids_list, weights):
        emb = w[i].float() if emb is None else emb + w[i]
    y, y2, mean, rstd = layernorm_fwd(emb, gamma, beta, eps, out_dtype, out2_dtype, save_stats=save)
    return (emb if save else None), y, y2, mean, rstd


def embedding_bwd(ids, dout, dweight, padding_idx=-1):
    flat = ids.reshape(-1)
    d = dout.reshape(-1, dweight.shape[1]).float()
    keep = (flat != padding_idx) & (flat >= 0) & (flat < dweight.shape[0])
    dweight.index_add_(0, flat[keep], d[keep])


def cross_entropy_fwd(logits2d, labels, S=0, shift=False, ignore_index=-100, want_dlogits=True):
    rows, V = logits2d.shape
    tgt = labels.reshape(-1).clone()
    if shift:
        t2 = torch.full_like(tgt, -100)
        t2[:-1] = tgt[1:]
        t2[torch.arange(rows) % S == S - 1] = -100
        tgt = t2
    valid = (tgt != ignore_index) & (tgt >= 0) & (tgt < V)
    cnt = max(int(valid.sum()), 1)
    x = logits2d.float()
    lse = torch.logsumexp(x, -1)
    safe = tgt.clamp(0, V - 1)
    loss = ((lse - x.gather(1, safe[:, None]).squeeze(1)) * valid).sum() / cnt
    dl = None
    if want_dlogits:
        p = torch.softmax(x, -1)
        p[torch.arange(rows), safe] -= 1.0
        dl = (p * valid[:, None] / cnt).to(logits2d.dtype)
    return loss.float(), dl


def cross_entropy_fwd_stats(logits2d, labels, row_stats, S=0, shift=False, ignore_index=-100, want_dlogits=True):
    loss, dl = cross_entropy_fwd(logits2d, labels, S, shift, ignore_index, want_dlogits)
    m2, s2 = row_stats[..., 0].double(), row_stats[..., 1].double()
    lse_stats = torch.logsumexp(m2 * math.log(2) + torch.log(s2.clamp_min(1e-300)), 0)
    assert torch.allclose(lse_stats, torch.logsumexp(logits2d.double(), -1), atol=1e-4), "row statistics disagree"
    return loss, dl


def scale_by_scalar(x, scalar_f32):
    x.mul_(float(scalar_f32))


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, mode=0, grad_scale=1.0, shadow=None):
    """mode 0 = torch.optim.AdamW (decoupled decay), mode 1 = the reference's class (optimizer.py:78-95: coupled L2,
    g rewritten, m_hat / (sqrt(v_hat) + eps)); include/ct_b200.h: ct_adamw_step."""
    gg = g * grad_scale
    if mode == 1:
        gg = gg + weight_decay * p
        g.copy_(gg)
    else:
        p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(gg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    if mode == 1:
        p.sub_(lr * (m / bc1) / (torch.sqrt(v / bc2) + eps))
    else:
        p.sub_((lr / bc1) * m / (torch.sqrt(v) / math.sqrt(bc2) + eps))
    if shadow is not None:
        shadow.copy_(p)


def adamw_multi(ps, gs, ms, vs, lr, beta1, beta2, eps, weight_decay, step, mode=0, grad_scale=1.0, shadows=None):
    for i in range(len(ps)):
        adamw_step(ps[i], gs[i], ms[i], vs[i], lr, beta1, beta2, eps, weight_decay, step, mode, grad_scale,
                   shadows[i] if shadows else None)


def sgd_step(p, g, buf, lr, momentum, dampening, weight_decay, first_step):
    """optimizer.py:28-50: g += wd*p; buf = first ? g : momentum*buf + (1-dampening)*g; g = buf; p -= lr*g."""
    if weight_decay:
        g.add_(p, alpha=weight_decay)
    if momentum:
        if first_step:
            buf.copy_(g)
        else:
            buf.mul_(momentum).add_(g, alpha=1 - dampening)
        g.copy_(buf)
    p.sub_(g, alpha=lr)


def kv_cache_append(past, new):
    """ops.kv_cache_append grows a preallocated cache in place; the observable result is the concatenation (a view of a
    buffer with at least ops.KV_CACHE_MIN_CAP rows, reachable as `._ct_cache_base`)."""
    from cleantransformer_b200 import ops
    cat = new if past is None else torch.cat([past, new], 2)
    B, H, t, D = cat.shape
    base = None
    if past is None and ops.KV_PREALLOC:  # a cached decode plan's buffer (generation.py)
        cand = ops.KV_PREALLOC.popleft()
        if cand.shape[:2] == (B, H) and cand.shape[3] == D and cand.shape[2] >= t and cand.dtype == cat.dtype:
            base = cand
    if base is None:
        base = torch.zeros(B, H, max(t, ops.KV_CACHE_MIN_CAP[0]), D, dtype=cat.dtype)
    base[:, :, :t] = cat
    view = base[:, :, :t]
    view._ct_cache_base = base
    return view


def kv_append_dev(base, new, len_dev):
    n, s = int(len_dev[0]), new.shape[2]
    if n - s >= 0 and n <= base.shape[2]:
        base[:, :, n - s:n] = new


def greedy_step(logits2d, alive, end_ids, pad_id, ids_out, cur_ids, pos_ids, state, sampled=None):
    """include/ct_b200.h: ct_greedy_step (generation_util.py:86-101; `sampled` replaces the argmax for do_sample)."""
    tok = torch.argmax(logits2d.float(), dim=-1) if sampled is None else sampled
    nxt = tok * alive + pad_id * (1 - alive)
    if end_ids is not None and end_ids.numel():
        hit = (nxt[None, :] == end_ids[:, None]).any(dim=0)
        alive.mul_((~hit).long())
    col = int(state[1])
    ids_out[:, col] = nxt
    cur_ids.copy_(nxt)
    if pos_ids is not None:
        pos_ids.add_(1)
    state[0] += 1
    state[1] += 1
    state[2] = int(alive.sum())
    if int(state[2]) == 0 and int(state[3]) < 0:
        state[3] = col + 1


def lm_head_stats_ok(M, V, dtype=None):
    return True


def lm_head_logits_with_stats(x2d, w):
    M, K = x2d.shape
    V = w.shape[0]
    stats = torch.empty((2 * ((V + 255) // 256), M, 2), dtype=torch.float32)
    return gemm(x2d, w, M, V, K, out_dtype=x2d.dtype, row_stats=stats), stats


@contextlib.contextmanager
def patched(compute_dtype=torch.float32):
    """Route cleantransformer_b200.ops through the stand-ins above and run the host layer in `compute_dtype`."""
    from cleantransformer_b200 import functional, ops
    names = ["layernorm_fwd", "layernorm_bwd", "cast", "colsum", "act_fwd", "act_bwd", "gemm", "attn_mask_prep",
             "attn_fwd", "attn_bwd", "embedding_fwd", "embedding_bwd", "embedding_layernorm_fwd", "cross_entropy_fwd", "cross_entropy_fwd_stats",
             "scale_by_scalar", "lm_head_stats_ok", "lm_head_logits_with_stats", "adamw_step", "adamw_multi", "sgd_step", "kv_cache_append",
             "kv_append_dev", "greedy_step", "dropout"]
    saved = {n: getattr(ops, n) for n in names}
    from cleantransformer_b200 import arena, generation, optimizer
    saved_gen = generation._on_device, generation._capture
    saved_req, saved_cd, saved_oreq = ops._req_cuda, functional.COMPUTE_DTYPE, optimizer._require_cuda
    saved_sh = arena.SHADOW_ON_ANY_DEVICE
    try:
        arena.SHADOW_ON_ANY_DEVICE = True
        for n in names:
            setattr(ops, n, globals()[n])
        ops._req_cuda = lambda *ts: None
        optimizer._require_cuda = lambda ps: None
        generation._on_device = lambda t: True          # the captured-decode host logic runs with ...
        generation._capture = lambda step: (step, 1)    # ... the "replay" simply calling the step
        functional.COMPUTE_DTYPE = compute_dtype
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        ops._req_cuda = saved_req
        optimizer._require_cuda = saved_oreq
        generation._on_device, generation._capture = saved_gen
        functional.COMPUTE_DTYPE = saved_cd
        arena.SHADOW_ON_ANY_DEVICE = saved_sh
