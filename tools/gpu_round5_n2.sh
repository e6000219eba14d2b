#!/bin/bash
# 2-GPU visit: collective + DDP wrapper check (tied table: early dense all-reduce + sparse exchange), bench at N=2
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== ddp_check"; date
timeout 240 $TR --master-port 29511 tools/ddp_check.py > $OUT/${TAG}_ddp_check.log 2>&1; echo "ddp_check rc=$?"
tail -4 $OUT/${TAG}_ddp_check.log | cut -c1-1500
cp $OUT/ddp_check_rank0.json $OUT/${TAG}_ddp_check_rank0.json 2>/dev/null
echo "== bench N=2 (sparse tied exchange)"; date
timeout 240 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench rc=$?"
tail -c 1800 $OUT/${TAG}_bench_n2.json; tail -3 $OUT/${TAG}_bench_n2.err
echo "== bench N=2 (dense tied bucket)"; date
CT_DDP_SPARSE_TIED=0 timeout 240 $TR --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n2_dense.json 2> $OUT/${TAG}_bench_n2_dense.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_n2_dense.json | head -1; tail -3 $OUT/${TAG}_bench_n2_dense.err
echo "== bench N=2 (nccl baseline collective)"; date
timeout 240 $TR --master-port 29514 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --comm nccl > $OUT/${TAG}_bench_n2_nccl.json 2> $OUT/${TAG}_bench_n2_nccl.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_n2_nccl.json | head -1; tail -3 $OUT/${TAG}_bench_n2_nccl.err
date
