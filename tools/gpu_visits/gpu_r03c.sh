#!/bin/bash
# r03c (1 GPU): captured decode with sampling: decode tests + a sampling throughput line
TAG=${1:-r03c}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/${TAG}_tests.log | cut -c1-250
timeout 600 python - > $OUT/${TAG}_sampling_bench.log 2>&1 <<'PY'
import os, sys, time, json, torch
sys.path.insert(0, os.getcwd())
from cleantransformer_b200.models import modeling_gpt as mg
L, NH, E, V, P, NEW, B = 24, 16, 1024, 50257, 32, 512, 32
m = mg.GPTLMHeadModel(mg.GPTConfig(vocab_size=V, n_embd=E, n_positions=1024, n_layer=L, n_head=NH, n_ctx=1024, afn="gelu_new"), version="gpt2").cuda().eval()
m._tie_weights()
ids = torch.randint(1, V, (B, P), device="cuda"); mask = torch.ones(B, P, dtype=torch.long, device="cuda")
gc = {"beam_size": 1, "do_sample": True, "max_gen_len": NEW - 2, "end_ids": None, "pad_id": 0, "temperature": 0.8, "top_k": 10, "top_p": 0.8}
res = {}
for graph in ("1", "0"):
    os.environ["CT_DECODE_GRAPH"] = graph
    n = 2 if graph == "1" else 1
    m.generate(ids, attention_mask=mask, generation_configs=gc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        out = m.generate(ids, attention_mask=mask, generation_configs=gc)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    res["captured" if graph == "1" else "host_loop"] = {"tokens_per_s": B * NEW / dt, "ms_per_generation": dt * 1e3}
print(json.dumps(res))
PY
echo "sampling rc=$?"; tail -2 $OUT/${TAG}_sampling_bench.log | cut -c1-300
date
