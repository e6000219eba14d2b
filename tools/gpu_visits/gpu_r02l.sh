#!/bin/bash
# r02l (1 GPU): skinny weight-streaming GEMM for the decode step: decode tests, suite, decode profile, bench arms
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
echo "== decode tests"; date
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > $OUT/${TAG}_decode_tests.log 2>&1; echo "decode rc=$?"; tail -25 $OUT/${TAG}_decode_tests.log
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -8 $OUT/${TAG}_tests.log
echo "== opt-in tests"; date
CT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "graphed_train_step or fused_lm_head" > $OUT/${TAG}_optin_tests.log 2>&1; echo "optin rc=$?"; tail -15 $OUT/${TAG}_optin_tests.log
echo "== decode profile"; date
timeout 600 python tools/decode_prof.py $OUT/${TAG}_decode_prof.json > $OUT/${TAG}_decode_prof.log 2>&1; echo "prof rc=$?"; head -60 $OUT/${TAG}_decode_prof.log | cut -c1-200
echo "== bench arms"; date
timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
CT_DECODE_GRAPH=0 timeout 600 python bench.py --workload gpt2_decode --steps 1 --warmup 3 --no-eager-baseline > $OUT/${TAG}_bench_gpt2_decode_nograph.json 2> $OUT/${TAG}_bench_gpt2_decode_nograph.err; echo "gpt2_decode (no graph) rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode_nograph.json | head -2
echo "== bench (default line)"; date
timeout 600 python bench.py --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -2
date
