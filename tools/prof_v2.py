"""ncu driver for the second-generation kernels at the Bloom-560M bench shapes: one warm-up launch and
one profiled launch each of LayerNorm backward, attention forward and attention backward.
  ncu --set full --clock-control none --import-source on -k regex:'ln_bwd_cta|attn_fwd_tc2|attn_bwd_tc2' \
      -s 3 -c 3 -o gpurun_out/prof_v2 python tools/prof_v2.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cleantransformer_b200 import ops
from oracle import ct_oracle as O

dev = "cuda"
T, H = 8192, 1024
x = torch.randn(T, H, device=dev); dy = torch.randn(T, H, device=dev).bfloat16(); extra = torch.randn(T, H, device=dev)
w = torch.randn(H, device=dev); b = torch.randn(H, device=dev)
_, _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5, out_dtype=torch.bfloat16)
dg = torch.empty(H, device=dev); db = torch.empty(H, device=dev); cs = torch.empty(H, device=dev)
B, Hh, S, D = 8, 16, 1024, 64
qkv = torch.randn(B, S, Hh, 3, D, device=dev).bfloat16()
q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
mask = torch.ones(B, S, dtype=torch.long, device=dev)
kb2, fv = ops.attn_mask_prep(mask, Hh, 0, O.alibi_slopes(Hh).cuda())
do = torch.randn(B, S, Hh * D, device=dev).bfloat16()
dq3 = torch.empty_like(qkv)
dq, dk, dv = [dq3[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
for _ in range(2):
    ops.layernorm_bwd(dy, x, w, mean, rstd, dg, db, False, dx_add=extra, dx2_dtype=torch.bfloat16, dxsum=cs)
    o, lse2 = ops.attn_fwd(q, k, v, 0.125, True, -ops.FLT_MAX, kb2, fv)
    ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, 0.125, True, -ops.FLT_MAX, kb2, fv)
torch.cuda.synchronize()
print("done")
