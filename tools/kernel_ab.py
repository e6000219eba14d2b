"""A/B of kernel variants at the Bloom-560M bench shapes: parity error against a torch fp32 restatement
and CUDA-event time per launch (inputs > L2 are rotated between launches), for
  LayerNorm backward  (LN_BWD_IMPL 1 = warp-per-row, 0 = row spread over cols/4 threads)
  attention forward / backward (default kernels; occupancy diagnostic)
  GEMM epilogues      (GEMM_EPI_IMPL 1 = generic, 2 = specialised with row-per-thread residual loads, 0 = specialised)
Prints one JSON line per measurement; never asserts (a failing variant shows up as a large error or an
`error` field), so one GPU visit tells everything.   python tools/kernel_ab.py [ln] [attn] [gemm]
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cleantransformer_b200 import ops  # noqa: E402
from oracle import ct_oracle as O  # noqa: E402

DEV = "cuda"
LOG2E = 1.4426950408889634


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def out(**kw):
    print(json.dumps(kw), flush=True)


def ln_ab():
    T, H = 8192, 1024
    torch.manual_seed(0)
    sets = []
    for _ in range(4):  # rotate 4 input sets (4 x 134 MB > L2)
        sets.append((torch.randn(T, H, device=DEV), torch.randn(T, H, device=DEV).bfloat16(),
                     torch.randn(T, H, device=DEV)))
    w = torch.randn(H, device=DEV); b = torch.randn(H, device=DEV)
    x, dy, extra = sets[0]
    y, _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5, out_dtype=torch.bfloat16)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    O.layernorm(xr, wr, br, 1e-5).backward(dy.float())
    want = xr.grad + extra
    stats = [ops.layernorm_fwd(s[0], w, b, 1e-5, out_dtype=torch.bfloat16)[2:] for s in sets]
    for impl in (1, 2, 0):
        prev = ops.set_option("LN_BWD_IMPL", impl)
        try:
            dg = torch.empty(H, device=DEV); db = torch.empty(H, device=DEV); cs = torch.empty(H, device=DEV)
            dx, dxl = ops.layernorm_bwd(dy, x, w, mean, rstd, dg, db, False, dx_add=extra, dx2_dtype=torch.bfloat16,
                                        dxsum=cs)
            errs = dict(dx=rel(dx, want), dx_bf16=rel(dxl, want), dgamma=rel(dg, wr.grad), dbeta=rel(db, br.grad),
                        dxsum=rel(cs, want.sum(0)))
            i = [0]

            def run_plain():
                xs, dys, ex = sets[i[0] % 4]; mu, rs = stats[i[0] % 4]; i[0] += 1
                ops.layernorm_bwd(dys, xs, w, mu, rs, dg, db, False, dx_add=ex)

            def run_fused():
                xs, dys, ex = sets[i[0] % 4]; mu, rs = stats[i[0] % 4]; i[0] += 1
                ops.layernorm_bwd(dys, xs, w, mu, rs, dg, db, False, dx_add=ex, dx2_dtype=torch.bfloat16, dxsum=cs)

            out(kernel="ln_bwd", impl=impl, err=errs, us_plain=timeit(run_plain), us_with_bf16_and_colsum=timeit(run_fused),
                algorithmic_MB=(T * H * (2 + 4 + 4 + 4)) / 1e6)
        except Exception as ex:  # noqa: BLE001
            out(kernel="ln_bwd", impl=impl, error=repr(ex)[:300])
        finally:
            ops.set_option("LN_BWD_IMPL", prev)


def _attn_oracle(q, k, v, scale, causal, causal_fill, kb2):
    Sq, Sk = q.shape[2], k.shape[2]
    s2 = (q.float() @ k.float().transpose(2, 3)) * (scale * LOG2E)
    kb = kb2[:, :, None, :] if kb2 is not None else 0.0
    s2 = s2 + kb
    if causal:
        i = torch.arange(Sq, device=q.device)[:, None]; j = torch.arange(Sk, device=q.device)[None, :]
        fill = torch.full_like(s2, causal_fill * LOG2E if causal_fill > -1e30 else float("-inf")) + kb
        s2 = torch.where(j > i + (Sk - Sq), fill, s2)
    s2 = s2.clamp_min(-3.4028234663852886e38)
    mx = s2.max(-1, keepdim=True).values
    e = torch.exp2(s2 - mx)
    o = (e / e.sum(-1, keepdim=True)) @ v.float()
    return o.transpose(1, 2).reshape(q.shape[0], Sq, -1)


def attn_ab():
    import ctypes
    from cleantransformer_b200 import _lib
    f, b = ctypes.c_int(0), ctypes.c_int(0)
    det = (ctypes.c_int * 8)()
    _lib.check(_lib.load().ct_attn_occupancy(ctypes.byref(f), ctypes.byref(b), det), "ct_attn_occupancy")
    out(kernel="attention", occupancy_ctas_per_sm=dict(forward=f.value, backward=b.value),
        forward_detail=dict(regs=det[0], static_smem=det[1], dynamic_smem=det[2],
                            ctas_per_sm_at_smem_minus_0_1_2_4_16KB=[det[3 + i] for i in range(5)]))
    cases = [  # name, B, H, S, mask mode, padding, fill
        ("bloom_bench_8x16x1024", 8, 16, 1024, 0, False, -ops.FLT_MAX),
        ("bloom_rightpad_2x16x1024", 2, 16, 1024, 0, True, -ops.FLT_MAX),
        ("gpt_nobias_2x16x1024", 2, 16, 1024, None, False, -1e4),
    ]
    for name, B, H, S, mode, pad, fill in cases:
        torch.manual_seed(1)
        D = 64
        qkv = torch.randn(B, S, H, 3, D, device=DEV).bfloat16()
        q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
        kb2 = fv = None
        if mode is not None:
            mask = torch.ones(B, S, dtype=torch.long, device=DEV)
            if pad:
                for b in range(B):
                    mask[b, S - 100 - 300 * b:] = 0
            kb2, fv = ops.attn_mask_prep(mask, H, mode, O.alibi_slopes(H).to(DEV))
        scale = 1.0 / math.sqrt(D)
        nb = min(B, 2)  # oracle on a slice (S^2 fp32 tensors)
        qr, kr, vr = [t[:nb].float().detach().requires_grad_(True) for t in (q, k, v)]
        ref = _attn_oracle(qr, kr, vr, scale, True, fill, kb2[:nb].expand(nb, H, S) if kb2 is not None else None)
        do = torch.randn(B, S, H * D, device=DEV).bfloat16()
        ref.backward(do[:nb].float())
        impls = ("bwd_8_compute_warps", "bwd_16_compute_warps")  # ATTN_BWD_IMPL 1 / 2 = default (same forward)
        if os.environ.get("CT_AB_ONLY_DEFAULT"):
            impls = impls[1:]
        for impl in impls:
            prev = ops.set_option("ATTN_BWD_IMPL", 2 if impl.startswith("bwd_16") else 1)
            try:
                o, lse2 = ops.attn_fwd(q, k, v, scale, True, fill, kb2, fv)
                dqkv = torch.zeros_like(qkv)
                dq, dk, dv = [dqkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
                ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, True, fill, kb2, fv)
                torch.cuda.synchronize()
                errs = dict(o=rel(o[:nb], ref), dq=rel(dq[:nb], qr.grad), dk=rel(dk[:nb], kr.grad), dv=rel(dv[:nb], vr.grad),
                            finite=bool(torch.isfinite(o.float()).all() and torch.isfinite(dqkv.float()).all()))
                us_f = timeit(lambda: ops.attn_fwd(q, k, v, scale, True, fill, kb2, fv))
                us_b = timeit(lambda: ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, True, fill, kb2, fv))
                flop_f = 4.0 * B * H * S * S * D / 2  # causal-counted
                out(kernel="attention", case=name, impl=impl, err=errs, us_fwd=us_f, us_bwd_incl_delta_and_dq_convert=us_b,
                    tflops_fwd=flop_f / us_f / 1e6, tflops_bwd=2.5 * flop_f / us_b / 1e6)
            except Exception as ex:  # noqa: BLE001
                out(kernel="attention", case=name, impl=impl, error=repr(ex)[:300])
            finally:
                ops.set_option("ATTN_BWD_IMPL", prev)


def gemm_ab():
    T, H = 8192, 1024
    torch.manual_seed(2)
    x = torch.randn(T, H, device=DEV).bfloat16()
    w1 = (torch.randn(4 * H, H, device=DEV) * 0.05).bfloat16(); b1 = torch.randn(4 * H, device=DEV)
    w2 = (torch.randn(H, 4 * H, device=DEV) * 0.05).bfloat16(); b2 = torch.randn(H, device=DEV)
    wq = (torch.randn(3 * H, H, device=DEV) * 0.05).bfloat16(); bq = torch.randn(3 * H, device=DEV)
    wo = (torch.randn(H, H, device=DEV) * 0.05).bfloat16(); bo = torch.randn(H, device=DEV)
    res = torch.randn(T, H, device=DEV)
    h4 = torch.randn(T, 4 * H, device=DEV).bfloat16()
    pre = torch.randn(T, 4 * H, device=DEV).bfloat16()
    dy = torch.randn(T, H, device=DEV).bfloat16()
    xf, hf = x.float(), h4.float()

    def gelu(t):
        return t * 0.5 * (1 + torch.tanh(0.79788456 * t * (1 + 0.044715 * t * t)))

    def gelu_grad(t):
        th = torch.tanh(0.79788456 * t * (1 + 0.044715 * t * t))
        return 0.5 * t * ((1 - th * th) * (0.79788456 + 0.1070322243 * t * t)) + 0.5 * (1 + th)

    ref_pre = xf @ w1.float().t() + b1
    jobs = [
        ("qkv_fwd_bias", 2.0 * T * 3 * H * H, lambda: ops.linear_fwd(x, wq, bq)[0], lambda: xf @ wq.float().t() + bq, 4e-3),
        ("ffn1_fwd_bias_gelu_preact", 2.0 * T * 4 * H * H,
         lambda: ops.linear_fwd(x, w1, b1, act=ops.ACT_GELU_TANH, save_preact=True)[0], lambda: gelu(ref_pre), 4e-3),
        ("ffn2_fwd_bias_residual_f32", 2.0 * T * 4 * H * H,
         lambda: ops.linear_fwd(h4, w2, b2, residual=res, out_dtype=torch.float32)[0],
         lambda: hf @ w2.float().t() + b2 + res, 1e-4),
        ("proj_fwd_bias_residual_f32", 2.0 * T * H * H,
         lambda: ops.linear_fwd(x, wo, bo, residual=res, out_dtype=torch.float32)[0],
         lambda: xf @ wo.float().t() + bo + res, 1e-4),
        ("ffn2_dgrad_actgrad", 2.0 * T * 4 * H * H,
         lambda: ops.linear_dgrad(dy, w2, actgrad_src=pre, actgrad_act=ops.ACT_GELU_TANH),
         lambda: (dy.float() @ w2.float()) * gelu_grad(pre.float()), 4e-3),
        ("ffn1_dgrad_plain", 2.0 * T * 4 * H * H, lambda: ops.linear_dgrad(h4, w1), lambda: hf @ w1.float(), 4e-3),
    ]
    for name, flop, fn, ref_fn, tol in jobs:
        ref = ref_fn()
        for impl in (1, 2, 0):
            prev = ops.set_option("GEMM_EPI_IMPL", impl)
            try:
                got = fn()
                torch.cuda.synchronize()
                us = timeit(fn)
                out(kernel="gemm", case=name, impl=impl, err=rel(got, ref), tol=tol, us=us, tflops=flop / us / 1e6)
            except Exception as ex:  # noqa: BLE001
                out(kernel="gemm", case=name, impl=impl, error=repr(ex)[:300])
            finally:
                ops.set_option("GEMM_EPI_IMPL", prev)
        del ref


def wgrad_ab():
    """split-K rule of the 2-CTA wgrad GEMMs (GEMM_SPLITK 1 = first-generation rule, 0 = wave-quantisation rule)."""
    T, H = 8192, 1024
    torch.manual_seed(4)
    for name, N, K in (("ffn2_wgrad", H, 4 * H), ("ffn1_wgrad", 4 * H, H), ("qkv_wgrad", 3 * H, H), ("proj_wgrad", H, H)):
        dy = torch.randn(T, N, device=DEV).bfloat16()
        x = torch.randn(T, K, device=DEV).bfloat16()
        ref = dy.float().t() @ x.float()
        refb = dy.float().sum(0)
        for rule in (1, 0):
            for overlap in (False,):
                prev = ops.set_option("GEMM_SPLITK", rule)
                try:
                    dw = torch.empty(N, K, device=DEV); db = torch.empty(N, device=DEV)
                    ops.linear_wgrad(dy, x, dw, db, accumulate=False)
                    torch.cuda.synchronize()
                    err, errb = rel(dw, ref), rel(db, refb)
                    dw.fill_(1.0); db.fill_(1.0)
                    ops.linear_wgrad(dy, x, dw, db, accumulate=True)
                    torch.cuda.synchronize()
                    err_acc = max(rel(dw - 1.0, ref), rel(db - 1.0, refb))
                    us = timeit(lambda: ops.linear_wgrad(dy, x, dw, db, accumulate=False))
                    out(kernel="wgrad", case=name, splitk_rule=rule, overlap_colsum=overlap, err=err, err_bias=errb,
                        err_accumulate=err_acc, us_gemm_plus_colsum=us, tflops=2.0 * T * N * K / us / 1e6)
                except Exception as ex:  # noqa: BLE001
                    out(kernel="wgrad", case=name, splitk_rule=rule, overlap_colsum=overlap, error=repr(ex)[:300])
                finally:
                    ops.set_option("GEMM_SPLITK", prev)
        del dy, x, ref


def ce_ab():
    rows, V = 8192, 250880
    torch.manual_seed(3)
    logits = torch.empty(rows, V, device=DEV, dtype=torch.bfloat16).normal_(0, 2)
    labels = torch.randint(0, V, (rows,), device=DEV)
    ref_rows = 64
    lr = logits[:ref_rows].float()
    ref_lse = torch.logsumexp(lr, -1)
    ref_loss_rows = ref_lse - lr.gather(1, labels[:ref_rows, None]).squeeze(1)
    for impl in (1, 3):
        prev = ops.set_option("CE_IMPL", impl)
        try:
            loss, dl = ops.cross_entropy_fwd(logits, labels, S=1024, shift=True)
            torch.cuda.synchronize()
            p = torch.softmax(lr, -1)
            tg = labels.view(8, 1024)[:, 1:]
            cnt = tg.numel()
            exp = p.clone()
            for r in range(ref_rows - 1):
                exp[r, labels[r + 1]] -= 1.0
            err = rel(dl[:ref_rows - 1].float() * cnt, exp[:ref_rows - 1])
            us = timeit(lambda: ops.cross_entropy_fwd(logits, labels, S=1024, shift=True))
            out(kernel="ce", impl=impl, loss=float(loss), err_dlogits=err, us=us,
                algorithmic_GBs=2.0 * rows * V * 2 / us / 1e3)
            del dl
        except Exception as ex:  # noqa: BLE001
            out(kernel="ce", impl=impl, error=repr(ex)[:300])
        finally:
            ops.set_option("CE_IMPL", prev)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ln", "attn", "gemm"]
    ops.device_check(0)
    for wname in which:
        {"ln": ln_ab, "attn": attn_ab, "gemm": gemm_ab, "ce": ce_ab, "wgrad": wgrad_ab}[wname]()
