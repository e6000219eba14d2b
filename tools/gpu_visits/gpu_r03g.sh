#!/bin/bash
# r03g (1 GPU): cached decode plan: decode tests, decode timing (plan reuse), decode bench
TAG=${1:-r03g}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -12 $OUT/${TAG}_tests.log | cut -c1-250
timeout 300 python tools/decode_timing.py $OUT/${TAG}_decode_timing.json 4 > $OUT/${TAG}_decode_timing.log 2>&1; echo "timing rc=$?"; cat $OUT/${TAG}_decode_timing.log | cut -c1-330
timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 --no-eager-baseline > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
grep -o '"ms_per_generation": {[^}]*}' $OUT/${TAG}_bench_gpt2_decode.json
date
