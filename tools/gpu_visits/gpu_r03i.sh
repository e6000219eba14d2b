#!/bin/bash
# r03i (1 GPU): complete decode bench line at HEAD (with the eager reference on the same GPU)
TAG=${1:-r03i}
OUT=gpurun_out
mkdir -p $OUT
timeout 240 python bench.py --workload gpt2_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -4
date
