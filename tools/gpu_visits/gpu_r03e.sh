#!/bin/bash
# r03e (1 GPU): final record of the round at HEAD: whole GPU suite, smoke, default bench + reference arm, decode / BERT arms
TAG=${1:-r03e}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -4
grep -o '"clocks": {[^}]*}' $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -4
timeout 600 python bench.py --workload bert_cls --steps 10 --warmup 3 > $OUT/${TAG}_bench_bert_cls.json 2> $OUT/${TAG}_bench_bert_cls.err; echo "bert_cls rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bert_cls.json | head -3
date
