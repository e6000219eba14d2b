#!/bin/bash
# r02zz (1 GPU): final record at HEAD: whole GPU suite, smoke, default bench
TAG=${1:-r02zz}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -4
grep -o '"clocks": {[^}]*}' $OUT/${TAG}_bench.json
date
