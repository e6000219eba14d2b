"""Tiny driver for ncu: a few attention fwd/bwd launches at the Bloom-560M shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cleantransformer_b200 import ops
from oracle import ct_oracle as O
B, H, S, D = 8, 16, 1024, 64
qkv = torch.randn(B, S, H, 3, D, device="cuda").bfloat16()
q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
mask = torch.ones(B, S, dtype=torch.long, device="cuda")
kb2, fv = ops.attn_mask_prep(mask, H, 0, O.alibi_slopes(H).cuda())
for _ in range(3):
    o, lse2 = ops.attn_fwd(q, k, v, 0.125, True, -ops.FLT_MAX, kb2, fv)
    do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
    dq, dk, dv = [dqkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, 0.125, True, -ops.FLT_MAX, kb2, fv)
torch.cuda.synchronize()
print("done")
