#!/bin/bash
# GPU visit r01f: attention backward variants (parity + timing), then a short bench per variant.
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_smi.txt
echo "== attention parity, new variants"; date
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention and (v2t or v3)" > $OUT/${TAG}_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -4 $OUT/${TAG}_attn_tests.log
echo "== attention A/B"; date
timeout 200 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab rc=$?"
cut -c1-330 $OUT/${TAG}_ab_attn.jsonl
tail -3 $OUT/${TAG}_ab_attn.err
for v in 2 3 4; do
  echo "== bench ATTN_BWD_IMPL=$v"; date
  CT_ATTN_BWD_IMPL=$v timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_bwd$v.json 2> $OUT/${TAG}_bench_bwd$v.err; echo "bench rc=$?"
  grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bwd$v.json | head -1
  grep -o '"loss": [0-9.]*' $OUT/${TAG}_bench_bwd$v.json | head -1
done
date
