#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of one step, ncu full capture (CSV raw
# page) of a 2-layer step. Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01c'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== tests"; date
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -5 $OUT/${TAG}_tests.log
echo "== bench"; date
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/${TAG}_bench.json
echo "== launches"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/step_prof.py > $OUT/${TAG}_launches.log 2>&1; echo "launches rc=$?"
echo "== ncu full (2 layers)"; date
LAYERS=2 timeout 900 ncu --set full --clock-control none --profile-from-start off --csv --page raw \
    --log-file $OUT/${TAG}_full.csv python tools/step_prof.py > $OUT/${TAG}_full.log 2>&1; echo "full rc=$?"
date
ls -la $OUT
