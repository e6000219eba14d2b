#!/bin/bash
# r02t (1 GPU): ncu --set full of the 16-warp backward, bench with the device-activity accounting, new variant test
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
echo "== variant test"; date
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "variants_agree" > $OUT/${TAG}_variant_test.log 2>&1; echo "test rc=$?"; tail -3 $OUT/${TAG}_variant_test.log | cut -c1-250
echo "== ncu backward (16 warps)"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tc2_kernel<1, 7>|attn_fwd_tc4" -c 2 -o $OUT/${TAG}_attn_ncu -f python tools/kernel_ab.py attn > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
echo "== bench"; date
timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -3
grep -o '"all_device_ms_per_step": [0-9.]*\|"kernel_ms_per_step": [0-9.]*\|"ms_per_step": [0-9.]*' $OUT/${TAG}_bench.json | head -5
date
