#!/bin/bash
# r02z (1 GPU): 16-warp attention backward without the per-tile CTA barrier: parity + A/B + ncu of that kernel + bench
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
echo "== attention tests"; date
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py -m gpu -q -x -k "attention or dropout" > $OUT/${TAG}_attn_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_attn_tests.log | cut -c1-250
echo "== attention A/B"; date
timeout 600 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attention.jsonl 2> $OUT/${TAG}_ab_attention.err; echo "ab rc=$?"; grep -o '"case": "[^"]*", "impl": "[^"]*"\|"us_bwd_incl_delta_and_dq_convert": [0-9.]*\|"error": "[^"]*"\|"dq": [0-9.e-]*' $OUT/${TAG}_ab_attention.jsonl | paste - - - | cut -c1-260
echo "== ncu backward (16 warps only)"; date
CT_AB_ONLY_DEFAULT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc2 -c 1 -o $OUT/${TAG}_attn_bwd16_ncu -f python tools/kernel_ab.py attn > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
echo "== bench"; date
timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -2
date
