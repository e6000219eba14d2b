#!/bin/bash
# r03b (1 GPU): new dropout / decode tests; ncu --set full of the skinny GEMM at HEAD
TAG=${1:-r03b}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_decode.py -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny --launch-skip 12 -c 8 -o $OUT/${TAG}_skinny_ncu -f python tools/skinny_prof.py > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
date
