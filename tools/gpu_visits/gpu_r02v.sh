#!/bin/bash
# r02v (2 GPUs): HEAD (PDL, 16-warp attention backward, chunked sparse exchange) through the DDP wrapper: parity + bench
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
echo "== multi-GPU parity tests (world 2)"; date
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $OUT/${TAG}_multi_tests.log 2>&1; echo "multi rc=$?"; tail -4 $OUT/${TAG}_multi_tests.log | cut -c1-300
cp $OUT/r02_ddp_check_w2.json $OUT/${TAG}_ddp_check_w2.json 2>/dev/null
echo "== bloom_sft at N=2"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err; echo "bloom N=2 rc=$?"; tail -2 $OUT/${TAG}_bench_2gpu.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_2gpu.json | head -3
CT_PDL=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench_2gpu_pdl_off.json 2> $OUT/${TAG}_bench_2gpu_pdl_off.err; echo "bloom N=2 (PDL off) rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_2gpu_pdl_off.json | head -2
date
