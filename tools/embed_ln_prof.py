"""A/B of the model preamble (SURVEY §8 f N3): ct_embedding_fwd (x tables) + ct_layernorm_fwd vs the fused
ct_embedding_layernorm_fwd, at the Bloom-560M (configs[1]) and BERT-base (configs[4]) shapes. CUDA events around
each variant, L2 flushed (256 MB write) before every timed call, 3 warm-ups + 20 timed calls; training form (the sum,
mean and rstd are saved for the backward). Prints one JSON line per shape and writes gpurun_out/<tag>_embed_ln.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleantransformer_b200 import ops  # noqa: E402

DEV = "cuda"


def timed(fn, flush, n=20, warm=3):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def case(name, B, S, H, vocabs, eps):
    torch.manual_seed(0)
    tables = [torch.randn(v, H, device=DEV) * 0.02 for v in vocabs]
    ids = [torch.randint(0, v, (B, S), device=DEV) for v in vocabs]
    gamma, beta = torch.ones(H, device=DEV), torch.zeros(H, device=DEV)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=DEV)

    def two():
        e = None
        for i, t in zip(ids, tables):
            e = ops.embedding_fwd(i, t, e, accumulate=e is not None)
        return ops.layernorm_fwd(e, gamma, beta, eps, torch.float32, None)

    def one():
        return ops.embedding_layernorm_fwd(ids, tables, gamma, beta, eps, torch.float32, None)

    y_two, y_one = two()[0], one()[1]
    err = float((y_two - y_one).abs().max())
    t_two, t_one = timed(two, flush), timed(one, flush)
    T = B * S
    alg = T * H * 4 * (len(vocabs) + 2)          # table rows read, sum written, y written
    out = {"case": name, "tokens": T, "H": H, "tables": len(vocabs), "two_kernels_us": 1e3 * t_two,
           "fused_us": 1e3 * t_one, "fused_GBps_algorithmic": alg / (t_one * 1e-3) / 1e9,
           "launches": [len(vocabs) + 1, 1], "max_abs_diff": err}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r04a"
    res = [case("bloom-560m preamble (8x1024, H=1024, V=250880)", 8, 1024, 1024, [250880], 1e-5),
           case("bert-base preamble (64x512, H=768, word+segment+position)", 64, 512, 768, [30522, 2, 512], 1e-12)]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/%s_embed_ln.json" % tag, "w") as f:
        json.dump(res, f, indent=1)
