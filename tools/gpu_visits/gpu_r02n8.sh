#!/bin/bash
# r02n8 (8 GPUs): HEAD at world 8: DDP / collective parity (tests/test_gpu_multi.py, world 8) and the default bench line
TAG=${1:-r02n8}
OUT=gpurun_out
mkdir -p $OUT
echo "== multi-GPU parity, world 8"; date
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8" > $OUT/${TAG}_multi_tests.log 2>&1; echo "multi rc=$?"; tail -3 $OUT/${TAG}_multi_tests.log | cut -c1-300
cp $OUT/r02_ddp_check_w8.json $OUT/${TAG}_ddp_check_w8.json 2>/dev/null
echo "== bloom_sft at N=8"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_8gpu.json 2> $OUT/${TAG}_bench_8gpu.err; echo "bloom N=8 rc=$?"; tail -2 $OUT/${TAG}_bench_8gpu.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_8gpu.json | head -3
date
