"""ncu driver: the dominant kernels at Bloom-560M shapes, a few launches each (QKV / FFN GEMMs with
their epilogues, a slice of the LM head, attention fwd/bwd)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cleantransformer_b200 import ops
from oracle import ct_oracle as O
dev = "cuda"
T, H = 8192, 1024
x = torch.randn(T, H, device=dev).bfloat16()
w_qkv = torch.randn(3 * H, H, device=dev).bfloat16(); b_qkv = torch.randn(3 * H, device=dev)
w1 = torch.randn(4 * H, H, device=dev).bfloat16(); b1 = torch.randn(4 * H, device=dev)
w2 = torch.randn(H, 4 * H, device=dev).bfloat16(); b2 = torch.randn(H, device=dev)
res = torch.randn(T, H, device=dev)
wv = torch.randn(32768, H, device=dev).bfloat16()
for _ in range(2):
    qkv, _ = ops.linear_fwd(x, w_qkv, b_qkv)                                      # launch A: QKV fwd
    h4, pre = ops.linear_fwd(x, w1, b1, act=ops.ACT_GELU_TANH, save_preact=True)  # B: FFN1 + GELU + preact
    out, _ = ops.linear_fwd(h4, w2, b2, residual=res, out_dtype=torch.float32)    # C: FFN2 + residual (fp32 out)
    dpre = ops.linear_dgrad(x, w2, actgrad_src=pre, actgrad_act=ops.ACT_GELU_TANH)  # D: dgrad + act' (K=1024 -> 4096)
    dw = torch.empty(4 * H, H, device=dev); db = torch.empty(4 * H, device=dev)
    ops.linear_wgrad(h4, x, dw, db)                                               # E: wgrad FFN1 (+colsum)
    lg, _ = ops.linear_fwd(x, wv)                                                 # F: LM-head slice (N=32768)
B, Hh, S, D = 8, 16, 1024, 64
q3 = torch.randn(B, S, Hh, 3, D, device=dev).bfloat16()
q, k, v = [q3[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
mask = torch.ones(B, S, dtype=torch.long, device=dev)
kb2, fv = ops.attn_mask_prep(mask, Hh, 0, O.alibi_slopes(Hh).cuda())
for _ in range(2):
    o, lse2 = ops.attn_fwd(q, k, v, 0.125, True, -ops.FLT_MAX, kb2, fv)
    do = torch.randn_like(o); dq3 = torch.empty_like(q3)
    dq, dk, dv = [dq3[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, 0.125, True, -ops.FLT_MAX, kb2, fv)
torch.cuda.synchronize()
print("done")
