#!/bin/bash
# r03h (1 GPU): last check of the round: whole GPU suite + smoke at HEAD
TAG=${1:-r03h}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log | cut -c1-200
date
