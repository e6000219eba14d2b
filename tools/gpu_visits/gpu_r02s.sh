#!/bin/bash
# r02s (1 GPU): attention backward, 16 compute warps + heaviest-first order as the default: A/B, ncu --set full, suite, bench
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
echo "== attention A/B"; date
timeout 600 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attention.jsonl 2> $OUT/${TAG}_ab_attention.err; echo "ab rc=$?"; grep -o '"case": "[^"]*", "impl": "[^"]*"\|"us_fwd": [0-9.]*\|"us_bwd_incl_delta_and_dq_convert": [0-9.]*\|"error": "[^"]*"\|"d[qkv]": [0-9.e-]*' $OUT/${TAG}_ab_attention.jsonl | paste - - - - - - | cut -c1-260
echo "== ncu backward"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc2 -c 2 -o $OUT/${TAG}_attn_bwd_ncu -f python tools/kernel_ab.py attn > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -4 $OUT/${TAG}_tests.log | cut -c1-250
echo "== bench"; date
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -3
grep -o '"clocks": {[^}]*}' $OUT/${TAG}_bench.json
date
