#!/bin/bash
# r02w (1 GPU): HEAD evidence: whole GPU suite, smoke, default bench + reference arm, extra arms, decode profile,
# ncu launch list of one un-graphed training step, ncu --set full of the 16-warp attention backward
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log | cut -c1-250
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log | cut -c1-200
echo "== bench (default line) + reference arm"; date
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
echo "== extra arms"; date
timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 > $OUT/${TAG}_bench_gpt2_decode.json 2> $OUT/${TAG}_bench_gpt2_decode.err; echo "gpt2_decode rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_gpt2_decode.json | head -3
timeout 600 python bench.py --workload bert_cls --steps 10 --warmup 3 > $OUT/${TAG}_bench_bert_cls.json 2> $OUT/${TAG}_bench_bert_cls.err; echo "bert_cls rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bert_cls.json | head -3
timeout 600 python tools/decode_prof.py $OUT/${TAG}_decode_prof.json > $OUT/${TAG}_decode_prof.log 2>&1; echo "decode prof rc=$?"
echo "== ncu launch list (one un-graphed step)"; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $OUT/${TAG}_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-kernel-table > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu --set full, attention backward (16 warps)"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tc2_kernel<1, 7>" -c 1 -o $OUT/${TAG}_attn_bwd16_ncu -f python tools/kernel_ab.py attn > $OUT/${TAG}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"; tail -2 $OUT/${TAG}_ncu_attn.log | cut -c1-200
date
