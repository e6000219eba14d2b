#!/bin/bash
# r02m (1 GPU): dropout kernels + fp16 GradScaler recipe tests, whole suite, skinny GEMM timing + ncu capture
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
echo "== dropout + fp16 tests"; date
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_models.py -m gpu -q -x -k "dropout or fp16" > $OUT/${TAG}_new_tests.log 2>&1; echo "new rc=$?"; tail -30 $OUT/${TAG}_new_tests.log | cut -c1-250
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -8 $OUT/${TAG}_tests.log | cut -c1-250
echo "== skinny timing"; date
timeout 300 python tools/skinny_prof.py $OUT/${TAG}_skinny.json > $OUT/${TAG}_skinny.log 2>&1; echo "skinny rc=$?"; cat $OUT/${TAG}_skinny.log | cut -c1-200
echo "== ncu skinny"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -c 12 -o $OUT/${TAG}_skinny_ncu -f python tools/skinny_prof.py > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
date
