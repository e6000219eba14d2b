#!/bin/bash
# r03d (2 GPUs): parameter sync through ct_broadcast (peer memory) in the DDP wrapper: world-2 parity
TAG=${1:-r03d}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "2" > $OUT/${TAG}_multi_tests.log 2>&1; echo "multi rc=$?"; tail -4 $OUT/${TAG}_multi_tests.log | cut -c1-300
cp $OUT/r02_ddp_check_w2.json $OUT/${TAG}_ddp_check_w2.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_ddp_check_w2.json'))
print("failures:", d['failures_all_ranks'], "param_sync:", d['ddp'].get('param_sync'), "p2p_vs_mean:", d['ddp'].get('p2p_vs_mean'))
PY
date
