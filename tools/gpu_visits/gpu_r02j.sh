#!/bin/bash
# r02j (1 GPU): HEAD after pruning: whole GPU suite, smoke, pipe probe, extra bench arms, bench + reference arm
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
echo "== pipes probe"; date
timeout 120 ./tools/_probe/pipes > $OUT/${TAG}_pipes.txt 2>&1; echo "pipes rc=$?"; cat $OUT/${TAG}_pipes.txt | head -40
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -5 $OUT/${TAG}_tests.log
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench arms"; date
timeout 600 python bench.py --workload gpt2_decode --steps 2 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
timeout 600 python bench.py --workload bert_cls --steps 10 --warmup 3 > $OUT/${TAG}_bench_bert_cls.json 2> $OUT/${TAG}_bench_bert_cls.err; echo "bert_cls rc=$?"; tail -2 $OUT/${TAG}_bench_bert_cls.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bert_cls.json | head -3
echo "== bench (default line) + reference arm"; date
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_reference.json | head -1
date
