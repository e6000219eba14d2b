"""Times (CUDA events, L2 flushed between calls) and exercises the skinny decode GEMM on the GPT-2-medium / Bloom-560M
decode shapes; run under `ncu --set full -k regex:gemm_skinny` to capture the kernels. Usage: python tools/skinny_prof.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from cleantransformer_b200 import ops
    shapes = [("h->h", 32, 1024, 1024), ("qkv", 32, 3072, 1024), ("fc1", 32, 4096, 1024), ("fc2", 32, 1024, 4096),
              ("lm_head gpt2", 32, 50257, 1024), ("lm_head bloom", 32, 250880, 1024)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for name, M, N, K in shapes:
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        bias = torch.randn(N, device="cuda")
        y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for impl in (4, 1):
            ts = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(x, w, M, N, K, out=y, bias=bias, impl=impl)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            us = sorted(ts[1:])[len(ts[1:]) // 2]
            out.append({"shape": name, "M": M, "N": N, "K": K, "impl": "skinny" if impl == 4 else "tcgen05 128-row",
                        "us": us, "weight_GBps": N * K * 2 / us / 1e3})
            print(out[-1])
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
