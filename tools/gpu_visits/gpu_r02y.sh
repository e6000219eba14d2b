#!/bin/bash
# r02y (1 GPU): observed errors of the model-level tests (to tighten their floors), smoke with the tightened bounds
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q > $OUT/${TAG}_model_tests.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_model_tests.log | cut -c1-200
cat $OUT/r02_model_test_errors.json | head -120
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
