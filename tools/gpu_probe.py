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


def mk_tc(M, N, K, a_mn, b_mn, epi=False):
    def f():
        return _gemm_case(M, N, K, a_mn, b_mn, 1, epi)
    return f


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
