"""Where does a GPT-2-medium greedy generation (bench.py --workload gpt2_decode shape) spend its time, generation by
generation: prefill / first eager step / graph capture / replays (CUDA events), host wall clock, allocator activity.
Usage: python tools/decode_timing.py [out.json] [n_generations]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from cleantransformer_b200.models import modeling_gpt as mg
    L, NH, E, V, P, NEW, B = 24, 16, 1024, 50257, 32, 512, 32
    cfg = dict(vocab_size=V, n_embd=E, n_positions=1024, n_layer=L, n_head=NH, n_ctx=1024, afn="gelu_new")
    model = mg.GPTLMHeadModel(mg.GPTConfig(**cfg), version="gpt2").cuda().eval()
    model._tie_weights()
    ids = torch.randint(1, V, (B, P), device="cuda")
    mask = torch.ones(B, P, dtype=torch.long, device="cuda")
    gc = {"beam_size": 1, "do_sample": False, "max_gen_len": NEW - 2, "end_ids": None, "pad_id": 0}
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    # (replay_issue_host_ms: host wall clock to issue the 510 graph launches of a generation; close to `replays` = the
    # device time means the launch queue throttled the host, i.e. the device is the bottleneck; far below = host is idle)
    model._ct_decode_trace = []
    rows = []
    for g in range(n):
        st0 = torch.cuda.memory_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.generate(ids, attention_mask=mask, generation_configs=gc)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        st1 = torch.cuda.memory_stats()
        r = dict(model._ct_decode_trace[-1])
        r["wall_ms"] = wall
        r["cudaMallocs"] = st1.get("num_device_alloc", 0) - st0.get("num_device_alloc", 0)
        r["cudaFrees"] = st1.get("num_device_free", 0) - st0.get("num_device_free", 0)
        r["reserved_MB"] = st1.get("reserved_bytes.all.current", 0) / 2 ** 20
        rows.append(r)
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
