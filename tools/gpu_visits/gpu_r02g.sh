#!/bin/bash
# r02g (1 GPU): attention occupancy detail + A/B, ncu --set full of the default attention kernels and of the LM-head
# GEMM (traffic), whole GPU suite, bench
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
echo "== attention A/B"; date
timeout 300 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab rc=$?"
cut -c1-330 $OUT/${TAG}_ab_attn.jsonl | head -4
echo "== ncu attention"; date
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc4|attn_bwd_tc2" -s 4 -c 2 -o $OUT/${TAG}_attn -f python tools/attn_prof.py > $OUT/${TAG}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"; tail -2 $OUT/${TAG}_ncu_attn.log
echo "== ncu LM-head GEMM traffic"; date
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 -s 2 -c 1 --csv --log-file $OUT/${TAG}_ncu_lmhead.csv python tools/lmhead_prof.py > $OUT/${TAG}_ncu_lmhead.log 2>&1; echo "ncu lmhead rc=$?"; tail -4 $OUT/${TAG}_ncu_lmhead.csv
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -5 $OUT/${TAG}_tests.log
echo "== bench + launch list"; date
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-kernel-table > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
date
