#!/bin/bash
# r02p (1 GPU): compile-time skinny variants + NVML clock sampler: decode tests, decode / default bench arms
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
echo "== decode tests"; date
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > $OUT/${TAG}_new_tests.log 2>&1; echo "new rc=$?"; tail -3 $OUT/${TAG}_new_tests.log | cut -c1-250
echo "== bench arms"; date
timeout 600 python bench.py --workload gpt2_decode --steps 4 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
grep -o '"ms_per_generation": {[^}]*}' $OUT/${TAG}_bench_gpt2_decode.json
grep -o '"clocks": {[^}]*}' $OUT/${TAG}_bench_gpt2_decode.json
echo "== decode profile"; date
timeout 600 python tools/decode_prof.py $OUT/${TAG}_decode_prof.json > $OUT/${TAG}_decode_prof.log 2>&1; echo "prof rc=$?"
echo "== bench (default line)"; date
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -3
grep -o '"clocks": {[^}]*}' $OUT/${TAG}_bench.json
date
