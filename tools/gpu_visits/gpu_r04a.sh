#!/bin/bash
# r04a (1 GPU, the last ~2 GPU-minutes of the round): the two new GPU test files (asynchronous checkpoint / Trainer
# resume; fused gather + LayerNorm) and the preamble A/B
TAG=${1:-r04a}
OUT=gpurun_out
mkdir -p $OUT
timeout 80 python -m pytest tests/test_gpu_z_embed_ln.py tests/test_gpu_z_checkpoint.py -m gpu -q > $OUT/${TAG}_new_tests.log 2>&1; echo "new tests rc=$?"; tail -4 $OUT/${TAG}_new_tests.log | cut -c1-300
timeout 30 python tools/embed_ln_prof.py $TAG > $OUT/${TAG}_embed_ln.log 2>&1; echo "prof rc=$?"; tail -2 $OUT/${TAG}_embed_ln.log | cut -c1-400
date
