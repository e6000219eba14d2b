#!/bin/bash
# r02o (1 GPU): decode timing per phase, decode tests with the 16-warp skinny tile, bench arms (decode, bert_cls with dropout)
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
echo "== decode tests"; date
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > $OUT/${TAG}_new_tests.log 2>&1; echo "new rc=$?"; tail -4 $OUT/${TAG}_new_tests.log | cut -c1-250
echo "== decode timing"; date
timeout 600 python tools/decode_timing.py $OUT/${TAG}_decode_timing.json 8 > $OUT/${TAG}_decode_timing.log 2>&1; echo "timing rc=$?"; cat $OUT/${TAG}_decode_timing.log | cut -c1-300
echo "== skinny timing"; date
timeout 300 python tools/skinny_prof.py $OUT/${TAG}_skinny.json > $OUT/${TAG}_skinny.log 2>&1; echo "skinny rc=$?"; grep skinny $OUT/${TAG}_skinny.log | cut -c1-200
echo "== decode profile"; date
timeout 600 python tools/decode_prof.py $OUT/${TAG}_decode_prof.json > $OUT/${TAG}_decode_prof.log 2>&1; echo "prof rc=$?"
echo "== bench arms"; date
timeout 600 python bench.py --workload gpt2_decode --steps 4 --warmup 3 --no-eager-baseline > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"; tail -2 $OUT/${TAG}_bench_gpt2_decode.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
grep -o '"ms_per_generation": {[^}]*}' $OUT/${TAG}_bench_gpt2_decode.json
timeout 600 python bench.py --workload bert_cls --steps 10 --warmup 3 > $OUT/${TAG}_bench_bert_cls.json 2> $OUT/${TAG}_bench_bert_cls.err; echo "bert_cls rc=$?"; tail -2 $OUT/${TAG}_bench_bert_cls.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bert_cls.json | head -3
CT_BENCH_BERT_DROPOUT=0 timeout 600 python bench.py --workload bert_cls --steps 10 --warmup 3 --no-eager-baseline > $OUT/${TAG}_bench_bert_cls_nodrop.json 2> $OUT/${TAG}_bench_bert_cls_nodrop.err; echo "bert_cls (p=0) rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bert_cls_nodrop.json | head -2
date
