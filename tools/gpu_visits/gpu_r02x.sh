#!/bin/bash
# r02x (1 GPU): careful A/B of programmatic dependent launch on the training step: alternating processes, 20 timed steps
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3; do
  for pdl in 1 2; do
    CT_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-kernel-table > $OUT/${TAG}_bench_pdl${pdl}_run${i}.json 2> $OUT/${TAG}_bench_pdl${pdl}_run${i}.err
    echo "pdl=$pdl run=$i rc=$? $(grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_pdl${pdl}_run${i}.json | head -2 | tr '\n' ' ') $(grep -o '"sm_mhz": [0-9.]*' $OUT/${TAG}_bench_pdl${pdl}_run${i}.json)"
  done
done
date
