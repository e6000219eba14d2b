"""Tiny driver for ncu: the largest launch of the dominant kernel (LM-head forward GEMM, M=8192 N=250880 K=1024)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cleantransformer_b200 import ops
x = torch.randn(8192, 1024, device="cuda").bfloat16()
w = (torch.randn(250880, 1024, device="cuda") * 0.02).bfloat16()
for _ in range(3):
    y, _ = ops.linear_fwd(x, w)
torch.cuda.synchronize()
print("done", float(y[0, 0]))
