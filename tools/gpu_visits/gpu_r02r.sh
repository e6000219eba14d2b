#!/bin/bash
# r02r (1 GPU): attention backward with 16 compute warps (ATTN_BWD_IMPL=2): A/B + parity, then the suite and bench with it
TAG=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
echo "== attention A/B"; date
timeout 600 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attention.jsonl 2> $OUT/${TAG}_ab_attention.err; echo "ab rc=$?"; grep -o '"case": "[^"]*", "impl": "[^"]*"\|"us_fwd": [0-9.]*\|"us_bwd_incl_delta_and_dq_convert": [0-9.]*\|"error": "[^"]*"\|"d[qkv]": [0-9.e-]*' $OUT/${TAG}_ab_attention.jsonl | paste - - - - - - | cut -c1-260; tail -3 $OUT/${TAG}_ab_attention.err | cut -c1-300
echo "== attention tests with the 16-warp backward"; date
CT_ATTN_BWD_IMPL=2 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_parity_shapes.py -m gpu -q -x -k "attention or dropout or config1 or config2 or config5" > $OUT/${TAG}_attn_tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_attn_tests.log | cut -c1-250
echo "== bench with the 16-warp backward"; date
CT_ATTN_BWD_IMPL=2 timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > $OUT/${TAG}_bench_bwd16.json 2> $OUT/${TAG}_bench_bwd16.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bwd16.json | head -2
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > $OUT/${TAG}_bench_bwd8.json 2> $OUT/${TAG}_bench_bwd8.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bwd8.json | head -2
date
