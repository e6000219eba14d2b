#!/bin/bash
# r02n (1 GPU): model-level dropout parity, decode with fused KV append + leaner skinny GEMM: tests, profile, bench arms
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
echo "== new tests"; date
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_dropout.py tests/test_gpu_parity_shapes.py -m gpu -q -x -k "decode or dropout or skinny or conv1d or greedy" > $OUT/${TAG}_new_tests.log 2>&1; echo "new rc=$?"; tail -30 $OUT/${TAG}_new_tests.log | cut -c1-250
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -8 $OUT/${TAG}_tests.log | cut -c1-250
echo "== skinny timing"; date
timeout 300 python tools/skinny_prof.py $OUT/${TAG}_skinny.json > $OUT/${TAG}_skinny.log 2>&1; echo "skinny rc=$?"; grep skinny $OUT/${TAG}_skinny.log | cut -c1-200
echo "== decode profile"; date
timeout 600 python tools/decode_prof.py $OUT/${TAG}_decode_prof.json > $OUT/${TAG}_decode_prof.log 2>&1; echo "prof rc=$?"
echo "== bench arms"; date
timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
grep -o '"ms_per_generation": {[^}]*}' $OUT/${TAG}_bench_gpt2_decode.json
echo "== bench (default line)"; date
timeout 600 python bench.py --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -2
date
