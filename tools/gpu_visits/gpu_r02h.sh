#!/bin/bash
# r02h (1 GPU): attention backward v7 (TMA reduce / TMA store / ALU bf16 packing) + forward ALU packing
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
echo "== attention tests"; date
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > $OUT/${TAG}_attn_tests.log 2>&1; echo "attn rc=$?"; tail -4 $OUT/${TAG}_attn_tests.log
echo "== attention A/B"; date
timeout 300 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab rc=$?"
cut -c1-400 $OUT/${TAG}_ab_attn.jsonl | head -8
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -5 $OUT/${TAG}_tests.log
echo "== bench"; date
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -1
CT_ATTN_BWD_IMPL=4 CT_ATTN_FWD_IMPL=3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-kernel-table > $OUT/${TAG}_bench_oldattn.json 2> /dev/null; echo "bench old attention rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_oldattn.json | head -1
echo "== ncu attention (new defaults)"; date
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc4|attn_bwd_tc7" -s 4 -c 2 -o $OUT/${TAG}_attn -f python tools/attn_prof.py > $OUT/${TAG}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"; tail -2 $OUT/${TAG}_ncu_attn.log
date
