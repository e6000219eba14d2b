#!/bin/bash
# Round-end GPU visit: whole GPU suite, bench line, knob cross-checks, eager-PyTorch arm on the same box,
# ncu launch list of one step, ncu --set full of the attention kernels.
TAG=${1:-r01z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== tests"; date
timeout 600 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -4 $OUT/${TAG}_tests.log
echo "== bench"; date
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 2600 $OUT/${TAG}_bench.json
for cfg in "CT_CE_IMPL=3" "CT_ATTN_BWD_IMPL=5" "CT_GEMM_SPLITK=1"; do
  name=$(echo "$cfg" | tr '=' '_')
  echo "== bench $cfg"; date
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> /dev/null; echo "rc=$?"
  grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_$name.json | head -1
done
echo "== eager PyTorch arm (reference path on this box)"; date
timeout 150 python bench.py --impl eager --steps 3 --warmup 3 > $OUT/${TAG}_bench_eager.json 2> $OUT/${TAG}_bench_eager.err; echo "eager rc=$?"
cut -c1-400 $OUT/${TAG}_bench_eager.json; tail -2 $OUT/${TAG}_bench_eager.err
echo "== launches"; date
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/step_prof.py > $OUT/${TAG}_launches.log 2>&1; echo "launches rc=$?"
echo "== ncu full: attention kernels"; date
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_tc2|attn_bwd_tc2' -s 2 -c 2 \
    -f -o $OUT/${TAG}_attn_full python tools/prof_v2.py > $OUT/${TAG}_attn_full.log 2>&1; echo "ncu full rc=$?"
date
ls -la $OUT | tail -15
