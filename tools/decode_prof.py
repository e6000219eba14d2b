"""Per-kernel time of the captured GPT-2-medium decode step (bench.py --workload gpt2_decode shape), from CUPTI via
torch.profiler: which kernels the 1-token step spends its time in. Usage: python tools/decode_prof.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from cleantransformer_b200.models import modeling_gpt as mg
    L, NH, E, V, P, NEW, B = 24, 16, 1024, 50257, 32, 128, 32
    cfg = dict(vocab_size=V, n_embd=E, n_positions=1024, n_layer=L, n_head=NH, n_ctx=1024, afn="gelu_new")
    model = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version="gpt2").cuda().eval()
    model._tie_weights()
    ids = torch.randint(1, V, (B, P), device="cuda")
    mask = torch.ones(B, P, dtype=torch.long, device="cuda")
    gc = {"beam_size": 1, "do_sample": False, "max_gen_len": NEW - 2, "end_ids": None, "pad_id": 0}
    for _ in range(2):
        model.generate(ids, attention_mask=mask, generation_configs=gc)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        model.generate(ids, attention_mask=mask, generation_configs=gc)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.name and "Memcpy" not in ev.name and "Memset" not in ev.name:
            r = rows.setdefault(ev.name[:110], [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time
    total = sum(r[1] for r in rows.values())
    table = sorted(((n, c, t) for n, (c, t) in rows.items()), key=lambda x: -x[2])
    out = {"new_tokens": NEW, "batch": B, "kernel_us_total": total, "kernel_us_per_token_step": total / NEW,
           "graph_launches_per_step": getattr(model, "_ct_decode_graph_launches", None),
           "kernels": [{"name": n, "calls": c, "us_total": t, "us_per_call": t / c, "share": t / total} for n, c, t in table[:25]]}
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
