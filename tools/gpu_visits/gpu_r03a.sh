#!/bin/bash
# r03a (1 GPU): bench with the asynchronous loss read-back in the e2e arm
TAG=${1:-r03a}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -3
grep -o '"loss": [0-9.]*, "loss_e2e": [0-9.]*' $OUT/${TAG}_bench.json
timeout 900 python bench.py --no-graph --no-cpu-baseline --no-eager-baseline --no-kernel-table > $OUT/${TAG}_bench_nograph.json 2> $OUT/${TAG}_bench_nograph.err; echo "bench (no graph) rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_nograph.json | head -2
date
