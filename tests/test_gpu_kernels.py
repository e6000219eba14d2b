"""GPU parity, kernel level: every C-ABI entry point against the oracle (oracle/ct_oracle.py, pinned
to the reference by tests/test_oracle_golden.py) on seeded inputs, plus size-independent properties
at BASELINE.json's full sizes (linearity of the GEMM, softmax rows summing to one through the
attention output, all-ones LayerNorm invariants).

Tolerances: err = ||x - ref||_inf / ||ref||_inf.
  fp32 kernels (LayerNorm, AdamW, SGD, CE on fp32, embedding, fp32-output GEMM): <= 1e-5 .. 1e-4
  kernels that round their OUTPUT to bf16: <= 4e-3 (one bf16 quantum, 2^-8, of the tensor maximum)
  attention (P rounded to bf16 before P.V, like the reference's autocast matmul): <= 5e-3
"""
import math
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
FLT_MAX = 3.4028234663852886e38
LOG2E = 1.4426950408889634


def _ops():
    from cleantransformer_b200 import ops
    return ops


@pytest.mark.parametrize("impl", [0, 2, 1], ids=["v2", "v2generic", "v1"])
@pytest.mark.parametrize("cols,rows", [(1024, 1000), (768, 1000), (128, 1000), (24, 1000), (1024, 8192), (256, 3)])
def test_layernorm_fwd_bwd_vs_oracle(cols, rows, impl):
    """impl 0 = default backward (row spread over cols/4 threads; compile-time variant on the training dtypes),
    2 = the same design with run-time dtypes, 1 = warp-per-row backward."""
    from oracle import ct_oracle as O
    ops = _ops()
    prev = ops.set_option("LN_BWD_IMPL", impl)
    try:
        torch.manual_seed(0)
        x = torch.randn(rows, cols, device=DEV)
        w = torch.randn(cols, device=DEV); b = torch.randn(cols, device=DEV)
        y, y2, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5, out_dtype=torch.float32, out2_dtype=torch.bfloat16)
        xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
        ref = O.layernorm(xr, wr, br, 1e-5)
        assert rel_err(y, ref) < 1e-5
        assert rel_err(y2, ref) < 4e-3
        dy = torch.randn_like(x)
        ref.backward(dy)
        dg = torch.empty(cols, device=DEV); db = torch.empty(cols, device=DEV)
        extra = torch.randn_like(x)
        dx = ops.layernorm_bwd(dy, x, w, mean, rstd, dg, db, False, dx_add=extra)
        assert rel_err(dx - extra, xr.grad) < 1e-4
        assert rel_err(dg, wr.grad) < 1e-4 and rel_err(db, br.grad) < 1e-4
        ops.layernorm_bwd(dy, x, w, mean, rstd, dg, db, True)  # accumulate
        assert rel_err(dg, 2 * wr.grad) < 1e-4
        # two incoming gradients (fp32 residual consumer + bf16 GEMM consumer)
        dy2 = torch.randn_like(x).bfloat16()
        dx2 = ops.layernorm_bwd(dy, x, w, mean, rstd, None, None, False, dy2=dy2)
        xr.grad = None
        O.layernorm(xr, w, b, 1e-5).backward(dy + dy2.float())
        assert rel_err(dx2, xr.grad) < 1e-4
        # fused by-products: low-precision copy of dx and its column sums (a bias gradient), bf16 dy
        dyl = dy.bfloat16()
        xr.grad = None
        O.layernorm(xr, w, b, 1e-5).backward(dyl.float())
        want = xr.grad + extra
        cs = torch.full((cols,), 7.0, device=DEV)
        dxf, dxl = ops.layernorm_bwd(dyl, x, w, mean, rstd, dg, db, False, dx_add=extra,
                                     dx2_dtype=torch.bfloat16, dxsum=cs, dxsum_accumulate=False)
        assert rel_err(dxf, want) < 1e-4 and dxl.dtype == torch.bfloat16 and rel_err(dxl, want) < 4e-3
        assert torch.equal(dxl, dxf.bfloat16())
        assert rel_err(cs, want.sum(0)) < 1e-4
        ops.layernorm_bwd(dyl, x, w, mean, rstd, None, None, False, dx_add=extra, dxsum=cs, dxsum_accumulate=True)
        assert rel_err(cs, 2 * want.sum(0)) < 1e-4
    finally:
        ops.set_option("LN_BWD_IMPL", prev)


def test_layernorm_bert_eps_and_empty():
    ops = _ops()
    x = torch.randn(7, 768, device=DEV)
    y, _, _, _ = ops.layernorm_fwd(x, torch.ones(768, device=DEV), torch.zeros(768, device=DEV), 1e-12)
    assert rel_err(y, torch.nn.functional.layer_norm(x, (768,), eps=1e-12)) < 1e-5
    e, _, _, _ = ops.layernorm_fwd(torch.empty(0, 768, device=DEV), torch.ones(768, device=DEV),
                                   torch.zeros(768, device=DEV), 1e-5)
    assert e.shape == (0, 768)


@pytest.mark.parametrize("mode", [0, 1])
def test_adamw_flat_and_multi_vs_oracle(mode):
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(1)
    n = 100003  # not a multiple of 4: exercises the tail
    p = torch.randn(n + 1, device=DEV)[:n + 1]
    p0 = torch.randn(n, device=DEV)
    stepf = O.adamw_torch_step if mode == 0 else O.adamw_reference_step
    pr, mr, vr = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pk = torch.zeros(n + 61, device=DEV)[:n]  # 16B-aligned base
    pk.copy_(p0)
    m = torch.zeros_like(pk); v = torch.zeros_like(pk)
    sh = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    for t in range(1, 4):
        g = torch.randn(n, device=DEV)
        pr, g_after, mr, vr = stepf(pr, g.clone(), mr, vr, t, lr=1e-2, weight_decay=0.05)
        gk = g.clone()
        ops.adamw_step(pk, gk, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.05, t, mode=mode, shadow=sh)
        assert rel_err(pk, pr) < 1e-5 and rel_err(m, mr) < 1e-5 and rel_err(v, vr) < 1e-5
        if mode == 1:
            assert rel_err(gk, g_after) < 1e-6  # the reference rewrites g (optimizer.py:80-81)
    assert rel_err(sh, pr) < 4e-3
    # multi-tensor entry point incl. a tensor whose base is not 16-byte aligned
    ts = [torch.randn(s, device=DEV) for s in (5, 1024, 77)] + [torch.randn(130, device=DEV)[1:]]
    gs = [torch.randn_like(t) for t in ts]
    ms = [torch.zeros_like(t) for t in ts]; vs = [torch.zeros_like(t) for t in ts]
    refs = [stepf(t.clone(), g.clone(), torch.zeros_like(t), torch.zeros_like(t), 1, lr=1e-2, weight_decay=0.05)[0]
            for t, g in zip(ts, gs)]
    ops.adamw_multi(ts, gs, ms, vs, 1e-2, 0.9, 0.999, 1e-8, 0.05, 1, mode=mode)
    for t, r in zip(ts, refs):
        assert rel_err(t, r) < 1e-5


def test_sgd_vs_oracle():
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(2)
    p = torch.randn(5000, device=DEV); pr = p.clone(); buf = torch.empty_like(p); bufr = None
    for t in range(3):
        g = torch.randn_like(p)
        pr, _, bufr = O.sgd_reference_step(pr, g.clone(), bufr, lr=0.01, momentum=0.9, dampening=0.1, weight_decay=0.01)
        ops.sgd_step(p, g.clone(), buf, 0.01, 0.9, 0.1, 0.01, t == 0)
        assert rel_err(p, pr) < 1e-5


MAJORS = [(0, 0), (0, 1), (1, 0), (1, 1)]


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("shape", [(512, 512, 512), (200, 136, 72), (128, 3072, 1024), (8, 64, 48), (712, 520, 200)])
@pytest.mark.parametrize("impl", [1, 2, 3])
def test_gemm_all_operand_majors(a_mn, b_mn, shape, impl):
    ops = _ops()
    M, N, K = shape
    if impl in (1, 3) and ((a_mn and M % 8) or (b_mn and N % 8) or (not a_mn and K % 8) or (not b_mn and K % 8)):
        pytest.skip("TMA needs 16-byte aligned leading dimensions")
    torch.manual_seed(3)
    A = torch.randn((K, M) if a_mn else (M, K), device=DEV).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=DEV).bfloat16()
    ref = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    out = ops.gemm(A, B, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32, impl=impl)
    assert rel_err(out, ref) < 1e-5
    out16 = ops.gemm(A, B, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.bfloat16, impl=impl)
    assert rel_err(out16, ref) < 4e-3


@pytest.mark.parametrize("act,name", [(1, "relu"), (2, "gelu"), (3, "gelu_new"), (4, "tanh")])
def test_linear_epilogues_vs_oracle(act, name):
    """y = act(x W^T + b) + residual with the pre-activation saved, then dgrad with act' fused."""
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(4)
    M, N, K = 384, 512, 256
    x = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    b = torch.randn(N, device=DEV); res = torch.randn(M, N, device=DEV)
    y, pre = ops.linear_fwd(x, w, b, act=act, residual=res, out_dtype=torch.float32, save_preact=True)
    t = O.linear(x.float(), w.float(), b)
    a = torch.tanh(t) if name == "tanh" else O.activation(t, name)
    assert rel_err(pre, t) < 4e-3
    assert rel_err(y, a + res) < 2e-3  # tanh.approx in the tanh-GELU epilogue: <= 5e-4 abs
    # fused activation-gradient in the next layer's dgrad: dx = (dy W2) * act'(pre)
    w2 = (torch.randn(64, N, device=DEV) * 0.1).bfloat16(); dy = torch.randn(M, 64, device=DEV).bfloat16()
    dx = ops.linear_dgrad(dy, w2, out_dtype=torch.float32, actgrad_src=pre, actgrad_act=act)
    pr = pre.float().clone().requires_grad_(True)
    (torch.tanh(pr) if name == "tanh" else O.activation(pr, name)).backward(dy.float() @ w2.float())
    assert rel_err(dx, pr.grad) < 3e-3
    # Conv1D layout ([in, out] weight, modeling_gpt.py:32-46)
    wc = w.t().contiguous()
    yc, _ = ops.linear_fwd(x, wc, b, out_dtype=torch.float32, w_in_out=True)
    assert rel_err(yc, O.conv1d(x.float(), wc.float(), b)) < 1e-5


@pytest.mark.parametrize("variant", ["v1", "v2", "v2row"])
@pytest.mark.parametrize("M,N,K", [(1024, 512, 256), (520, 1056, 320), (2048, 1024, 1024)])
def test_gemm_specialised_epilogues_vs_oracle(M, N, K, variant):
    """The compile-time epilogue kinds of the 2-CTA kernel (v2; GEMM_EPI_IMPL=0) and the generic run-time
    epilogue (v1) on the four hot combinations: bias->bf16, bias+GELU+saved pre-activation, bias+f32
    residual -> f32, dgrad x GELU'(pre). M=520 exercises ragged row tiles, N=1056 a ragged 256-column tile."""
    from oracle import ct_oracle as O
    ops = _ops()
    # "v2row": compile-time epilogues with the first form of the f32 residual fetch (row per thread); "v2" fetches it
    # in the coalesced write-out pattern and adds it after the transpose through the staging tile
    prev = ops.set_option("GEMM_EPI_IMPL", {"v1": 1, "v2": 0, "v2row": 2}[variant])
    try:
        torch.manual_seed(11)
        x = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
        b = torch.randn(N, device=DEV); res = torch.randn(M, N, device=DEV)
        t = O.linear(x.float(), w.float(), b)
        y, _ = ops.linear_fwd(x, w, b)                                            # bias -> bf16
        assert rel_err(y, t) < 4e-3
        y0, _ = ops.linear_fwd(x, w, None)                                        # no bias -> bf16
        assert rel_err(y0, x.float() @ w.float().t()) < 4e-3
        g, pre = ops.linear_fwd(x, w, b, act=ops.ACT_GELU_TANH, save_preact=True)  # bias + GELU + pre-activation
        assert rel_err(pre, t) < 4e-3 and rel_err(g, O.activation(t, "gelu_new")) < 4e-3
        r, _ = ops.linear_fwd(x, w, b, residual=res, out_dtype=torch.float32)      # bias + residual -> f32
        assert rel_err(r, t + res) < 1e-5
        r0, _ = ops.linear_fwd(x, w, None, residual=res, out_dtype=torch.float32)
        assert rel_err(r0, x.float() @ w.float().t() + res) < 1e-5
        dy = torch.randn(M, K, device=DEV).bfloat16()                              # dgrad [M,K]x[K->N] * GELU'(pre)
        w2 = (torch.randn(K, N, device=DEV) * 0.1).bfloat16()                      # Linear N -> K: weight [K, N]
        dx = ops.linear_dgrad(dy, w2, actgrad_src=pre, actgrad_act=ops.ACT_GELU_TANH)
        pr = pre.float().clone().requires_grad_(True)
        O.activation(pr, "gelu_new").backward(dy.float() @ w2.float())
        assert rel_err(dx, pr.grad) < 5e-3
        # forward that saves gelu'(t) instead of t (one tanh for both) + backward that only multiplies
        g2, dpre = ops.linear_fwd(x, w, b, act=ops.ACT_GELU_TANH_SAVE_GRAD, save_preact=True)
        assert torch.equal(g2, g)
        tt = t.clone().requires_grad_(True)
        O.activation(tt, "gelu_new").sum().backward()
        assert rel_err(dpre, tt.grad) < 4e-3
        dx2 = ops.linear_dgrad(dy, w2, actgrad_src=dpre, actgrad_act=ops.ACT_GRAD_PRECOMPUTED)
        pr2 = t.clone().requires_grad_(True)  # reference at the unrounded pre-activation
        O.activation(pr2, "gelu_new").backward(dy.float() @ w2.float())
        assert rel_err(dx2, pr2.grad) < 8e-3
    finally:
        ops.set_option("GEMM_EPI_IMPL", prev)


def test_wgrad_splitk_accumulate_and_bias():
    ops = _ops()
    torch.manual_seed(5)
    M, N, K = 4096, 1024, 1024
    dy = torch.randn(M, N, device=DEV).bfloat16(); x = torch.randn(M, K, device=DEV).bfloat16()
    dw = torch.empty(N, K, device=DEV); db = torch.empty(N, device=DEV)
    ops.linear_wgrad(dy, x, dw, db, accumulate=False)
    ref = dy.float().t() @ x.float()
    assert rel_err(dw, ref) < 1e-5 and rel_err(db, dy.float().sum(0)) < 1e-5
    ops.linear_wgrad(dy, x, dw, db, accumulate=True)
    assert rel_err(dw, 2 * ref) < 1e-5
    dwc = torch.empty(K, N, device=DEV)
    ops.linear_wgrad(dy, x, dwc, None, w_in_out=True)
    assert rel_err(dwc, ref.t()) < 1e-5


def test_gemm_linearity_at_full_size():
    """Size-independent property at the Bloom-560M FFN shape: (A1 + A2) B == A1 B + A2 B exactly in
    fp32 accumulation up to rounding; checked on a 1/64 sample of the output."""
    ops = _ops()
    torch.manual_seed(6)
    M, N, K = 8192, 4096, 1024
    A1 = torch.randint(-4, 5, (M, K), device=DEV).bfloat16(); A2 = torch.randint(-4, 5, (M, K), device=DEV).bfloat16()
    B = torch.randint(-4, 5, (N, K), device=DEV).bfloat16()
    s = ops.gemm((A1 + A2), B, M, N, K, out_dtype=torch.float32)
    a = ops.gemm(A1, B, M, N, K, out_dtype=torch.float32); b = ops.gemm(A2, B, M, N, K, out_dtype=torch.float32)
    assert torch.equal(s, a + b)  # small-integer inputs: every partial sum is exact in fp32
    assert torch.equal(s[::64, ::64], ((A1 + A2).float()[::64] @ B.float()[::64].t()))


def _attn_oracle(q, k, v, scale, causal, causal_fill, kb2):
    """Score definition of include/ct_b200.h in plain torch (fp32). q,k,v [B,H,S,D]."""
    Sq, Sk = q.shape[2], k.shape[2]
    s2 = (q.float() @ k.float().transpose(2, 3)) * (scale * LOG2E)
    kb = kb2[:, :, None, :] if kb2 is not None else 0.0
    s2 = s2 + kb
    if causal:
        i = torch.arange(Sq, device=q.device)[:, None]; j = torch.arange(Sk, device=q.device)[None, :]
        fill = torch.full_like(s2, causal_fill * LOG2E if causal_fill > -1e30 else float("-inf")) + kb
        s2 = torch.where(j > i + (Sk - Sq), fill, s2)
    s2 = s2.clamp_min(-FLT_MAX)
    mx = s2.max(-1, keepdim=True).values
    e = torch.exp2(s2 - mx)
    o = (e / e.sum(-1, keepdim=True)) @ v.float()
    return o.transpose(1, 2).reshape(q.shape[0], Sq, -1)


ATT_CASES = [
    # B, H, Sq, Sk, D, causal, mask mode, pad side, causal_fill, impl
    (2, 8, 12, 12, 8, True, 0, "right", -FLT_MAX, 2),      # golden-test head size (SIMT)
    (3, 4, 8, 8, 12, True, 1, "left", -1e4, 2),
    (2, 4, 1, 33, 64, True, 1, "left", -1e4, 2),            # q_len = 1 decode against a cache
    (2, 4, 256, 256, 64, False, None, "none", -FLT_MAX, 1),
    (3, 4, 300, 300, 64, True, 0, "right", -FLT_MAX, 1),    # Bloom: ALiBi + causal + right padding, ragged
    (3, 4, 300, 300, 64, True, 1, "left", -1e4, 1),         # GPT: -1e4 replace + finfo.min, left padding
    (2, 4, 512, 512, 64, False, 2, "right", -FLT_MAX, 1),   # BERT additive mask
    (2, 4, 128, 384, 64, True, 1, "none", -1e4, 1),         # prefill chunk against a longer cache
    (2, 4, 640, 640, 64, True, 0, "none", -FLT_MAX, 1),     # Bloom, no padding: interior + aligned-diagonal tiles
    (2, 4, 512, 512, 64, True, None, "none", -FLT_MAX, 1),  # causal, no per-key bias at all
    (2, 4, 384, 384, 64, True, None, "none", -1e4, 1),      # causal with the finite GPT fill, no bias
    (2, 2, 256, 640, 64, True, 0, "none", -FLT_MAX, 1),     # Sq < Sk: diagonal offset by 384 (aligned)
    (2, 2, 200, 440, 64, True, 0, "right", -FLT_MAX, 1),    # diagonal offset (240) not a multiple of 128
]
ATT_PARAMS = ATT_CASES


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("B,H,Sq,Sk,D,causal,mode,pad,cfill,impl", ATT_PARAMS)
def test_attention_fwd_bwd_vs_oracle(B, H, Sq, Sk, D, causal, mode, pad, cfill, impl, dtype):
    if dtype == torch.float16 and (impl == 2 or Sq == 300):
        pytest.skip("f16 operands: tcgen05 path only, one shape per mask family")
    _attention_case(B, H, Sq, Sk, D, causal, mode, pad, cfill, impl, dtype)


def _attention_case(B, H, Sq, Sk, D, causal, mode, pad, cfill, impl, dtype):
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(7)
    qkv = torch.randn(B, Sk, H, 3, D, device=DEV).to(dtype)
    q = qkv[:, Sk - Sq:, :, 0, :].permute(0, 2, 1, 3); k = qkv[..., 1, :].permute(0, 2, 1, 3); v = qkv[..., 2, :].permute(0, 2, 1, 3)
    kb2 = fv = None
    if mode is not None:
        mask = torch.ones(B, Sk, dtype=torch.long, device=DEV)
        for b in range(B):
            n = Sk - (b * 37) % (Sk // 2 + 1)
            if pad == "right": mask[b, n:] = 0
            if pad == "left": mask[b, :Sk - n] = 0
        slopes = O.alibi_slopes(H).to(DEV) if mode == 0 else None
        kb2, fv = ops.attn_mask_prep(mask, H, mode, slopes)
        if mode == 0:  # the prepared bias must equal the reference's alibi + fill construction
            al = O.build_alibi_tensor(mask, H, torch.float32).view(B, H, Sk)
            exp = torch.where(mask[:, None, :] == 1, al * LOG2E, torch.tensor(float("-inf"), device=DEV))
            assert torch.allclose(kb2, exp, rtol=1e-6, atol=1e-6)
    scale = 1.0 / math.sqrt(D)
    o, lse2 = ops.attn_fwd(q, k, v, scale, causal, cfill, kb2, fv, impl=impl)
    kbe = kb2.expand(B, H, Sk) if kb2 is not None else None
    qr, kr, vr = [t.float().detach().requires_grad_(True) for t in (q, k, v)]
    ref = _attn_oracle(qr, kr, vr, scale, causal, cfill, kbe)
    assert rel_err(o, ref) < 5e-3
    if Sq == 1:
        return
    do = torch.randn_like(ref).to(dtype)
    dqkv = torch.zeros_like(qkv)
    dq = dqkv[:, Sk - Sq:, :, 0, :].permute(0, 2, 1, 3); dk = dqkv[..., 1, :].permute(0, 2, 1, 3); dv = dqkv[..., 2, :].permute(0, 2, 1, 3)
    ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, causal, cfill, kb2, fv, impl=impl)
    ref.backward(do.float())
    assert rel_err(dq, qr.grad) < 8e-3 and rel_err(dk, kr.grad) < 8e-3 and rel_err(dv, vr.grad) < 8e-3


def test_attention_matches_reference_bloom_layer(golden):
    """End to end against the oracle's restatement of modeling_bloom.py:76-124 (alibi.baddbmm,
    masked_fill(finfo.min), softmax, PV) on a d=64 configuration."""
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(8)
    B, S, H, D = 2, 200, 4, 64
    hid = H * D
    x = torch.randn(B, S, hid, device=DEV)
    wqkv = torch.randn(3 * hid, hid, device=DEV) * 0.05; bqkv = torch.randn(3 * hid, device=DEV) * 0.1
    mask = torch.ones(B, S, dtype=torch.long, device=DEV); mask[1, 150:] = 0
    alibi = O.build_alibi_tensor(mask, H, torch.float32)
    mb = O.bloom_attn_mask(mask, (B, S))
    eye = torch.eye(hid, device=DEV)
    ref, _ = O.bloom_attention(x, torch.zeros_like(x), alibi, mb, wqkv, bqkv, eye, torch.zeros(hid, device=DEV), H)
    qkv = (x.view(-1, hid) @ wqkv.t() + bqkv).bfloat16().view(B, S, 3 * hid)
    from cleantransformer_b200 import functional as F
    q, k, v = F.split_packed(qkv, H, F.LAYOUT_BLOOM)
    kb2, fv = ops.attn_mask_prep(mask, H, 0, O.alibi_slopes(H).to(DEV))
    o, _ = ops.attn_fwd(q, k, v, 1.0 / math.sqrt(D), True, -FLT_MAX, kb2, fv)
    assert rel_err(o, ref) < 8e-3


def test_attention_rows_sum_to_one_at_full_size():
    """Property at BASELINE.json's size (B=8,H=16,S=1024,d=64): with V = 1 every output element is
    the softmax row sum, i.e. exactly 1 up to bf16 rounding of P; checks every tile incl. the diagonal."""
    ops = _ops()
    from oracle import ct_oracle as O
    torch.manual_seed(9)
    B, H, S, D = 8, 16, 1024, 64
    qkv = torch.randn(B, S, H, 3, D, device=DEV).bfloat16()
    qkv[..., 2, :] = 1.0
    q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    mask = torch.ones(B, S, dtype=torch.long, device=DEV); mask[3, 700:] = 0
    kb2, fv = ops.attn_mask_prep(mask, H, 0, O.alibi_slopes(H).to(DEV))
    o, lse2 = ops.attn_fwd(q, k, v, 0.125, True, -FLT_MAX, kb2, fv)
    assert float((o.float() - 1).abs().max()) < 8e-3
    assert torch.isfinite(lse2).all()


def test_cross_entropy_and_embedding_vs_oracle():
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(10)
    B, S, V = 3, 17, 1000
    logits = (torch.randn(B, S, V, device=DEV) * 3).bfloat16()
    labels = torch.randint(0, V, (B, S), device=DEV)
    loss, dl = ops.cross_entropy_fwd(logits.view(B * S, V), labels.view(-1), S=S, shift=True)
    lr = logits.float().clone().requires_grad_(True)
    ref = O.shifted_lm_loss(lr, labels)
    ref.backward()
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < 1e-5
    assert rel_err(dl.view(B, S, V), lr.grad) < 4e-3
    lf = torch.randn(50, 777, device=DEV); lab = torch.randint(0, 777, (50,), device=DEV); lab[::5] = -100
    loss2, dl2 = ops.cross_entropy_fwd(lf, lab, S=0, shift=False)
    lr2 = lf.clone().requires_grad_(True)
    ref2 = torch.nn.functional.cross_entropy(lr2, lab); ref2.backward()
    assert abs(float(loss2) - float(ref2)) / abs(float(ref2)) < 1e-5 and rel_err(dl2, lr2.grad) < 1e-5
    W = torch.randn(500, 64, device=DEV); ids = torch.randint(0, 500, (4, 9), device=DEV); ids[0, 0] = 0
    assert torch.equal(ops.embedding_fwd(ids, W), W[ids])
    dout = torch.randn(4, 9, 64, device=DEV); dW = torch.zeros_like(W)
    ops.embedding_bwd(ids, dout, dW, padding_idx=0)
    Wr = W.clone().requires_grad_(True)
    torch.nn.functional.embedding(ids, Wr, padding_idx=0).backward(dout)
    assert rel_err(dW, Wr.grad) < 1e-6


@pytest.mark.parametrize("impl", [1, 2, 3], ids=["twopass", "cluster", "l2resident"])
@pytest.mark.parametrize("rows,S,V,shift", [(51, 17, 1000, True), (40, 0, 8 * 4099, False), (24, 8, 250880, True),
                                            (9, 3, 16, True)])
def test_cross_entropy_variants_vs_torch(impl, rows, S, V, shift):
    """CE_IMPL 1 = two-pass kernel, 2 = row resident in the shared memory of a 4-CTA cluster (one HBM pass), 3 = two
    passes with one row per SM in flight (second read from L2); bf16
    logits: ragged slices (V/8 not a multiple of the cluster size or of the chunk), a full Bloom vocabulary row,
    slices shorter than one chunk and CTAs without any element, ignored targets."""
    ops = _ops()
    torch.manual_seed(12)
    logits = (torch.randn(rows, V, device=DEV) * 4).bfloat16()
    labels = torch.randint(0, V, (rows,), device=DEV)
    labels[::7] = -100
    prev = ops.set_option("CE_IMPL", impl)
    try:
        loss, dl = ops.cross_entropy_fwd(logits, labels, S=S, shift=shift)
        torch.cuda.synchronize()
    finally:
        ops.set_option("CE_IMPL", prev)
    lr = logits.float().clone().requires_grad_(True)
    if shift:
        x = lr.view(rows // S, S, V)[:, :-1].reshape(-1, V); t = labels.view(rows // S, S)[:, 1:].reshape(-1)
    else:
        x, t = lr, labels
    ref = torch.nn.functional.cross_entropy(x, t)
    ref.backward()
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < 1e-5
    assert rel_err(dl, lr.grad) < 4e-3
    assert torch.isfinite(dl.float()).all()


def test_cast_colsum_activations():
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(11)
    x = torch.randn(1000, 333, device=DEV) * 2
    assert torch.equal(ops.cast(x, torch.bfloat16), x.bfloat16())
    out = torch.zeros(333, device=DEV)
    ops.colsum(x, out, False)
    assert rel_err(out, x.sum(0)) < 1e-5
    assert rel_err(ops.act_fwd(x, ops.ACT_GELU_TANH), O.gelu_tanh_bloom(x)) < 1e-3
    g = torch.randn_like(x)
    assert rel_err(ops.act_bwd(g, x, ops.ACT_GELU_TANH), O.gelu_tanh_bloom_back(g, x)) < 2e-3
    assert rel_err(ops.act_fwd(x, ops.ACT_GELU_ERF), O.gelu_erf(x)) < 1e-5


def test_kv_cache_append_matches_concat():
    """In-place cache growth (ct_kv_append) == the reference's torch.concat along the time axis
    (modeling_bloom.py:88-92, modeling_gpt.py:76-80), across a capacity re-allocation."""
    ops = _ops()
    torch.manual_seed(12)
    B, H, D = 2, 3, 64
    ref = None
    cache = None
    for step, s in enumerate([5, 1, 1, 300, 1]):
        new = torch.randn(B, s, H, 3, D, device=DEV).bfloat16()[..., 1, :].permute(0, 2, 1, 3)  # strided view
        ref = new.contiguous() if ref is None else torch.cat((ref, new), dim=2)
        cache = ops.kv_cache_append(cache, new)
        assert cache.shape == ref.shape
        assert torch.equal(cache, ref)
    assert cache._ct_cache_base.shape[2] >= cache.shape[2]


@pytest.mark.skipif(not os.environ.get("CT_TEST_EXPERIMENTAL"),
                    reason="opt-in (CT_TEST_EXPERIMENTAL=1); the same path runs at the full 8192 x 250880 shape in "
                           "test_gpu_parity_shapes.py::test_config2_...[fused_lm_stats]")
@pytest.mark.parametrize("M,V,K", [(512, 1024, 256), (640, 2080, 320), (1024, 250880 // 8, 128)])
def test_lm_head_row_stats_and_streaming_cross_entropy(M, V, K):
    """The logits GEMM's per-row softmax statistics reproduce logsumexp of the STORED bf16 logits, and the one-pass loss
    kernel that consumes them equals the two-pass one (ragged last 256-column tile: V = 2080; ignored targets)."""
    ops = _ops()
    torch.manual_seed(13)
    x = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(V, K, device=DEV) * 0.2).bfloat16()
    logits, stats = ops.lm_head_logits_with_stats(x, w)
    ref_logits, _ = ops.linear_fwd(x, w)
    assert torch.equal(logits, ref_logits)
    m2, s2 = stats[..., 0], stats[..., 1]                      # [slots, M]
    lse2 = torch.logsumexp(m2.double() * math.log(2) + torch.log(s2.double().clamp_min(1e-300)), dim=0) / math.log(2)
    ref = torch.logsumexp(logits.double(), dim=1) / math.log(2)
    assert float((lse2 - ref).abs().max()) < 2e-4
    S = 64
    labels = torch.randint(0, V, (M,), device=DEV)
    labels[::11] = -100
    loss_s, dl_s = ops.cross_entropy_fwd_stats(logits, labels, stats, S=S, shift=True)
    prev = ops.set_option("CE_IMPL", 1)
    try:
        loss_r, dl_r = ops.cross_entropy_fwd(logits, labels, S=S, shift=True)
    finally:
        ops.set_option("CE_IMPL", prev)
    assert abs(float(loss_s) - float(loss_r)) <= 2e-5 * abs(float(loss_r))
    assert rel_err(dl_s, dl_r) < 4e-3


@pytest.mark.parametrize("B,H,S,causal,mode", [(2, 4, 384, True, 0), (1, 3, 300, False, 2), (2, 2, 1024, True, None)])
def test_attention_backward_variants_agree_bit_for_bit(B, H, S, causal, mode):
    """ATTN_BWD_IMPL: the default backward (16 compute warps, one 32-query chunk each, heaviest key tiles first) and the
    8-warp kernel it replaced run the same arithmetic per element in the same order: dK / dV identical bits, dQ identical
    up to the order of its fp32 atomics."""
    from oracle import ct_oracle as O
    ops = _ops()
    torch.manual_seed(3)
    D = 64
    qkv = torch.randn(B, S, H, 3, D, device=DEV).bfloat16()
    q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    kb2 = fv = None
    if mode is not None:
        mask = torch.ones(B, S, dtype=torch.long, device=DEV)
        mask[0, S - 70:] = 0
        kb2, fv = ops.attn_mask_prep(mask, H, mode, O.alibi_slopes(H).to(DEV) if mode == 0 else None)
    scale = 0.125
    o, lse2 = ops.attn_fwd(q, k, v, scale, causal, -FLT_MAX, kb2, fv)
    do = torch.randn_like(o)
    outs = []
    for impl in (1, 0):
        prev = ops.set_option("ATTN_BWD_IMPL", impl)
        try:
            d = torch.zeros_like(qkv)
            dq, dk, dv = [d[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
            ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, scale, causal, -FLT_MAX, kb2, fv)
            torch.cuda.synchronize()
            outs.append((dq.clone(), dk.clone(), dv.clone()))
        finally:
            ops.set_option("ATTN_BWD_IMPL", prev)
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert rel_err(outs[0][0], outs[1][0].float()) < 1e-2  # bf16 quantum after differently ordered fp32 atomics
