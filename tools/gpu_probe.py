"""GPU probe: run each kernel check in its own process (a device trap must not poison the rest).
Usage: python tools/gpu_probe.py [case ...]   -> appends JSON lines to gpurun_out/probe.jsonl
"""
import json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def rel(a, b):
    import torch
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def case_ln():
    import torch
    from cleantransformer_b200 import ops
    torch.manual_seed(0)
    res = {}
    for cols in (1024, 768, 24):
        x = torch.randn(4096, cols, device="cuda")
        g = torch.randn(cols, device="cuda"); b = torch.randn(cols, device="cuda")
        y, y2, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, out_dtype=torch.float32, out2_dtype=torch.bfloat16)
        ref = torch.nn.functional.layer_norm(x, (cols,), g, b, 1e-5)
        res["fwd%d" % cols] = rel(y, ref); res["fwd_bf%d" % cols] = rel(y2, ref)
        dy = torch.randn_like(x)
        xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-5).backward(dy)
        dg = torch.empty(cols, device="cuda"); db = torch.empty(cols, device="cuda")
        dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db, False)
        res["dx%d" % cols] = rel(dx, xr.grad); res["dg%d" % cols] = rel(dg, gr.grad); res["db%d" % cols] = rel(db, br.grad)
    return res


def case_adamw():
    import torch
    from cleantransformer_b200 import ops
    torch.manual_seed(0)
    n = 1 << 20
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda")
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=0.01)
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    sh = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda")
        pr.grad = g.clone(); opt.step()
        ops.adamw_step(p, g.clone(), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, mode=0, shadow=sh)
    return {"p": rel(p, pr.detach()), "shadow": rel(sh, pr.detach())}


def _gemm_case(M, N, K, a_mn, b_mn, impl, epi=False):
    import torch
    from cleantransformer_b200 import ops
    torch.manual_seed(1)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    Af = (A.float().t() if a_mn else A.float()); Bf = (B.float().t() if b_mn else B.float())
    ref = Af @ Bf.t()
    kw = {}
    if epi:
        bias = torch.randn(N, device="cuda"); resid = torch.randn(M, N, device="cuda")
        pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(bias=bias, act=ops.ACT_GELU_TANH, preact=pre, residual=resid)
        t = ref + bias
        ref2 = torch.nn.functional.gelu(t, approximate="tanh") + resid
    out = ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, impl=impl, **kw)
    torch.cuda.synchronize()
    if epi:
        return {"out": rel(out, ref2), "pre": rel(pre, t)}
    return {"out": rel(out, ref)}


def case_gemm_simt():
    r = {}
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            r["%d%d" % (a_mn, b_mn)] = _gemm_case(200, 136, 72, a_mn, b_mn, 2)["out"]
    r["epi"] = _gemm_case(200, 136, 72, 0, 0, 2, epi=True)
    return r


def mk_tc(M, N, K, a_mn, b_mn, epi=False, impl=1):
    def f():
        return _gemm_case(M, N, K, a_mn, b_mn, impl, epi)
    return f


def case_gemm_perf_2cta():
    import torch
    from cleantransformer_b200 import ops
    res = {}
    for (M, N, K) in [(8192, 3072, 1024), (8192, 4096, 1024), (8192, 1024, 4096), (8192, 8192, 8192), (8192, 32768, 1024)]:
        A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        r = {}
        for impl in (1, 3):
            for _ in range(3):
                ops.gemm(A, B, M, N, K, out=out, impl=impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.gemm(A, B, M, N, K, out=out, impl=impl)
            e1.record(); torch.cuda.synchronize()
            r["impl%d_tflops" % impl] = 2 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9
        res["%dx%dx%d" % (M, N, K)] = r
    return res


def case_gemm_wgrad_splitk():
    import torch
    from cleantransformer_b200 import ops
    torch.manual_seed(2)
    M, N, K = 4096, 1024, 1024  # tokens, out, in
    dy = torch.randn(M, N, device="cuda").bfloat16(); x = torch.randn(M, K, device="cuda").bfloat16()
    dw = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
    ops.linear_wgrad(dy, x, dw, db, accumulate=False)
    ops.linear_wgrad(dy, x, dw, db, accumulate=True)
    ref = 2 * (dy.float().t() @ x.float())
    return {"dw": rel(dw, ref), "db": rel(db, 2 * dy.float().sum(0))}


def case_gemm_perf():
    import torch
    from cleantransformer_b200 import ops
    res = {}
    for (M, N, K) in [(8192, 3072, 1024), (8192, 4096, 1024), (8192, 1024, 4096), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(A, B, M, N, K, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(A, B, M, N, K, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10):
            torch.matmul(A, B.t(), out=out)
        e1.record(); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        res["%dx%dx%d" % (M, N, K)] = {"ct_tflops": 2 * M * N * K / ms / 1e9, "cublas_tflops": 2 * M * N * K / ms_t / 1e9}
    return res



def _attn_ref(q, k, v, scale, causal, causal_fill, kb2):
    """fp32 torch restatement of the ct_b200.h score definition. q,k,v [B,H,S,D] (any dtype)."""
    import torch
    FLT_MAX = 3.4028234663852886e38
    LOG2E = 1.4426950408889634
    qf, kf, vf = q.float(), k.float(), v.float()
    Sq, Sk = q.shape[2], k.shape[2]
    s2 = (qf @ kf.transpose(2, 3)) * (scale * LOG2E)
    kb = kb2[:, :, None, :] if kb2 is not None else 0.0
    s2 = s2 + kb
    if causal:
        i = torch.arange(Sq, device=q.device)[:, None]; j = torch.arange(Sk, device=q.device)[None, :]
        fut = j > i + (Sk - Sq)
        fill = torch.full_like(s2, causal_fill * LOG2E if causal_fill > -1e30 else float("-inf")) + kb
        s2 = torch.where(fut, fill, s2)
    s2 = s2.clamp_min(-FLT_MAX)
    m = s2.max(-1, keepdim=True).values
    e = torch.exp2(s2 - m)
    l = e.sum(-1, keepdim=True)
    o = (e / l) @ vf
    return o.transpose(1, 2).reshape(q.shape[0], Sq, -1), (m + torch.log2(l)).squeeze(-1)


def _attn_case(B, H, Sq, Sk, D, causal, mode, impl, bwd=False, pad="none", causal_fill=None):
    import torch
    from cleantransformer_b200 import ops
    torch.manual_seed(7)
    dev = "cuda"
    qkv = (torch.randn(B, Sk, H, 3, D, device=dev) * 1.0).bfloat16()
    q = qkv[:, Sk - Sq:, :, 0, :].permute(0, 2, 1, 3); k = qkv[..., 1, :].permute(0, 2, 1, 3); v = qkv[..., 2, :].permute(0, 2, 1, 3)
    kb2 = fv = None
    if mode is not None:
        mask = torch.ones(B, Sk, dtype=torch.long, device=dev)
        for b in range(B):
            n = Sk - (b * 37) % (Sk // 2)
            if pad == "right": mask[b, n:] = 0
            if pad == "left": mask[b, :Sk - n] = 0
        slopes = None
        if mode == 0:
            from oracle import ct_oracle as O
            slopes = O.alibi_slopes(H).to(dev)
        kb2, fv = ops.attn_mask_prep(mask, H, mode, slopes)
    cf = causal_fill if causal_fill is not None else -ops.FLT_MAX
    scale = 1.0 / D ** 0.5
    o, lse2 = ops.attn_fwd(q, k, v, scale, causal, cf, kb2, fv, impl=impl)
    torch.cuda.synchronize()
    kbe = kb2.expand(B, H, Sk) if kb2 is not None else None
    qr, kr, vr = [t.float().detach().requires_grad_(True) for t in (q, k, v)]
    oref, lref = _attn_ref(qr, kr, vr, scale, causal, cf, kbe)
    res = {"o": rel(o, oref), "lse": float((lse2 - lref).abs().max())}
    if bwd:
        do = (torch.randn_like(o.float()) * 1.0).bfloat16()
        dqkv = torch.zeros_like(qkv)
        dq = dqkv[:, Sk - Sq:, :, 0, :].permute(0, 2, 1, 3); dk = dqkv[..., 1, :].permute(0, 2, 1, 3); dv = dqkv[..., 2, :].permute(0, 2, 1, 3)
        ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, causal, cf, kb2, fv, impl=impl)
        torch.cuda.synchronize()
        oref.backward(do.float())
        res.update({"dq": rel(dq, qr.grad), "dk": rel(dk, kr.grad), "dv": rel(dv, vr.grad)})
    return res


def mk_attn(*a, **kw):
    def f():
        return _attn_case(*a, **kw)
    return f


def case_attn_perf():
    import torch
    from cleantransformer_b200 import ops
    B, H, S, D = 8, 16, 1024, 64
    qkv = torch.randn(B, S, H, 3, D, device="cuda").bfloat16()
    q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    mask = torch.ones(B, S, dtype=torch.long, device="cuda")
    from oracle import ct_oracle as O
    kb2, fv = ops.attn_mask_prep(mask, H, 0, O.alibi_slopes(H).cuda())
    scale = 0.125
    res = {}
    def timeit(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    o, lse2 = ops.attn_fwd(q, k, v, scale, True, -ops.FLT_MAX, kb2, fv)
    ms = timeit(lambda: ops.attn_fwd(q, k, v, scale, True, -ops.FLT_MAX, kb2, fv))
    fl = 4 * B * H * S * S * D / 2
    res["fwd_causal_ms"] = ms; res["fwd_causal_tflops"] = fl / ms / 1e9
    ms = timeit(lambda: ops.attn_fwd(q, k, v, scale, False))
    res["fwd_dense_ms"] = ms; res["fwd_dense_tflops"] = 2 * fl / ms / 1e9
    do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
    dq, dk, dv = [dqkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    ms = timeit(lambda: ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, True, -ops.FLT_MAX, kb2, fv))
    res["bwd_causal_ms"] = ms; res["bwd_causal_tflops"] = 2.5 * fl / ms / 1e9
    import torch.nn.functional as F
    qs, ks, vs = [t.contiguous() for t in (q, k, v)]
    ms = timeit(lambda: F.scaled_dot_product_attention(qs, ks, vs, is_causal=True))
    res["sdpa_fwd_causal_ms"] = ms
    return res



def case_ce_embed():
    import torch
    from cleantransformer_b200 import ops
    torch.manual_seed(3)
    res = {}
    B, S, V = 3, 17, 1000
    logits = (torch.randn(B, S, V, device="cuda") * 3).bfloat16()
    labels = torch.randint(0, V, (B, S), device="cuda")
    loss, dl = ops.cross_entropy_fwd(logits.view(B * S, V), labels.view(-1), S=S, shift=True)
    lr = logits.float().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1))
    ref.backward()
    res["loss_shift"] = abs(float(loss) - float(ref)) / abs(float(ref))
    res["dlogits_shift"] = rel(dl.view(B, S, V), lr.grad)
    lab2 = labels.clone().view(-1); lab2[::5] = -100
    lf = torch.randn(B * S, 777, device="cuda")
    loss2, dl2 = ops.cross_entropy_fwd(lf, lab2.clamp(max=776), S=0, shift=False)
    lr2 = lf.clone().requires_grad_(True)
    ref2 = torch.nn.functional.cross_entropy(lr2, lab2.clamp(max=776)); ref2.backward()
    res["loss_plain"] = abs(float(loss2) - float(ref2)) / abs(float(ref2)); res["dlogits_plain"] = rel(dl2, lr2.grad)
    W = torch.randn(500, 64, device="cuda"); ids = torch.randint(0, 500, (4, 9), device="cuda")
    out = ops.embedding_fwd(ids, W)
    res["emb_fwd"] = rel(out, W[ids])
    dout = torch.randn(4, 9, 64, device="cuda"); dW = torch.zeros_like(W)
    ops.embedding_bwd(ids, dout, dW)
    Wr = W.clone().requires_grad_(True); torch.nn.functional.embedding(ids, Wr).backward(dout)
    res["emb_bwd"] = rel(dW, Wr.grad)
    return res

CASES = {
    "ln": case_ln,
    "adamw": case_adamw,
    "gemm_simt": case_gemm_simt,
    "tc_128x128x64_kk": mk_tc(128, 128, 64, 0, 0),
    "tc_128x256x256_kk": mk_tc(128, 256, 256, 0, 0),
    "tc_512x512x512_kk": mk_tc(512, 512, 512, 0, 0),
    "tc_512x512x512_km": mk_tc(512, 512, 512, 0, 1),
    "tc_512x512x512_mk": mk_tc(512, 512, 512, 1, 0),
    "tc_512x512x512_mm": mk_tc(512, 512, 512, 1, 1),
    "tc_ragged_kk": mk_tc(200, 136, 72, 0, 0),
    "tc_ragged_mm": mk_tc(200, 136, 72, 1, 1),
    "tc_big_kk": mk_tc(8192, 3072, 1024, 0, 0),
    "tc_epi": mk_tc(1024, 1024, 512, 0, 0, True),
    "tc_wgrad_splitk": case_gemm_wgrad_splitk,
    "gemm_perf": case_gemm_perf,
    "tc2_256x256x64_kk": mk_tc(256, 256, 64, 0, 0, impl=3),
    "tc2_512x512x512_kk": mk_tc(512, 512, 512, 0, 0, impl=3),
    "tc2_512x512x512_km": mk_tc(512, 512, 512, 0, 1, impl=3),
    "tc2_512x512x512_mk": mk_tc(512, 512, 512, 1, 0, impl=3),
    "tc2_512x512x512_mm": mk_tc(512, 512, 512, 1, 1, impl=3),
    "tc2_ragged_kk": mk_tc(200, 136, 72, 0, 0, impl=3),
    "tc2_ragged_mm": mk_tc(712, 520, 200, 1, 1, impl=3),
    "tc2_big_kk": mk_tc(8192, 3072, 1024, 0, 0, impl=3),
    "tc2_epi": mk_tc(1024, 1024, 512, 0, 0, True, impl=3),
    "gemm_perf_2cta": case_gemm_perf_2cta,
    "attn_simt_d8_bloom": mk_attn(2, 8, 12, 12, 8, True, 0, 2, bwd=True, pad="right"),
    "attn_simt_d12_gpt": mk_attn(3, 4, 8, 8, 12, True, 1, 2, bwd=True, pad="left", causal_fill=-1e4),
    "attn_simt_d64_bert": mk_attn(2, 4, 40, 40, 64, False, 2, 2, bwd=True, pad="right"),
    "attn_simt_decode": mk_attn(2, 4, 1, 33, 64, True, 1, 2, pad="left", causal_fill=-1e4),
    "attn_tc_plain": mk_attn(2, 4, 256, 256, 64, False, None, 1),
    "attn_tc_causal": mk_attn(2, 4, 384, 384, 64, True, None, 1),
    "attn_tc_bloom_ragged": mk_attn(3, 4, 300, 300, 64, True, 0, 1, pad="right"),
    "attn_tc_gpt_left": mk_attn(3, 4, 300, 300, 64, True, 1, 1, pad="left", causal_fill=-1e4),
    "attn_tc_bert": mk_attn(2, 4, 512, 512, 64, False, 2, 1, pad="right"),
    "attn_tc_prefill_off": mk_attn(2, 4, 128, 384, 64, True, 1, 1, pad="none", causal_fill=-1e4),
    "attn_tc_bwd_plain": mk_attn(2, 4, 256, 256, 64, False, None, 1, bwd=True),
    "attn_tc_bwd_causal": mk_attn(2, 4, 384, 384, 64, True, None, 1, bwd=True),
    "attn_tc_bwd_bloom_ragged": mk_attn(3, 4, 300, 300, 64, True, 0, 1, bwd=True, pad="right"),
    "attn_tc_bwd_gpt_left": mk_attn(3, 4, 300, 300, 64, True, 1, 1, bwd=True, pad="left", causal_fill=-1e4),
    "attn_tc_bwd_big": mk_attn(2, 16, 1024, 1024, 64, True, 0, 1, bwd=True, pad="right"),
    "attn_perf": case_attn_perf,
    "ce_embed": case_ce_embed,
}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        name = sys.argv[2]
        try:
            r = {"case": name, "ok": True, "result": CASES[name]()}
        except Exception as ex:  # noqa
            r = {"case": name, "ok": False, "error": repr(ex)[:2000]}
        print("PROBE_RESULT " + json.dumps(r), flush=True)
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    with open(os.path.join(OUT, "probe.jsonl"), "a") as f:
        for name in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=180)
                line = [l for l in p.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
                r = json.loads(line[-1][len("PROBE_RESULT "):]) if line else {"case": name, "ok": False, "error": "no result", "stdout": p.stdout[-1500:], "stderr": p.stderr[-1500:]}
                if not r.get("ok"):
                    r["stderr"] = p.stderr[-1500:]; r["stdout"] = p.stdout[-800:]
            except subprocess.TimeoutExpired:
                r = {"case": name, "ok": False, "error": "timeout"}
            r["sec"] = round(time.time() - t0, 1)
            f.write(json.dumps(r) + "\n"); f.flush()
            print(json.dumps(r), flush=True)
