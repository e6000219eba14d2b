#!/bin/bash
# r02i (8 GPUs): parity + bandwidth at world 8 (NVLS), DDP bench variants at N=8 (+ per-kernel table under contention)
TAG=${1:-r02i}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export CT_COMM_TIMEOUT_S=90
nvidia-smi -L | head -8
echo "== ddp_check (full, $N GPUs)"; date
timeout 600 $TR --master-port 29531 tools/ddp_check.py --out $OUT/${TAG}_ddp_check_w$N.json > $OUT/${TAG}_ddp_check.log 2>&1; echo "ddp_check rc=$?"
grep -v "^W1017\|^\[W" $OUT/${TAG}_ddp_check.log | tail -2 | cut -c1-1500
echo "== bench N=$N"; date
port=29540
run() { name=$1; shift; port=$((port+1));
  env "$@" timeout 500 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 $FLAGS > $OUT/${TAG}_bench_n${N}_$name.json 2> $OUT/${TAG}_bench_n${N}_$name.err; echo "bench $name rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $OUT/${TAG}_bench_n${N}_$name.json | head -1; grep -o '"ddp_nvls": [a-z]*' $OUT/${TAG}_bench_n${N}_$name.json | head -1; tail -1 $OUT/${TAG}_bench_n${N}_$name.err | cut -c1-300; }
FLAGS="" run default CT_X=0
FLAGS="--no-kernel-table --no-eager-baseline" run ctas8 CT_DDP_CTAS=8
FLAGS="--no-kernel-table --no-eager-baseline" run ctas32 CT_DDP_CTAS=32
FLAGS="--no-kernel-table --no-eager-baseline" run nvls_off CT_DDP_NVLS=0
FLAGS="--no-kernel-table --no-eager-baseline --comm nccl --no-graph" run nccl CT_X=0
FLAGS="--no-kernel-table --no-eager-baseline" run skipcomm CT_DDP_SKIP_COMM=1
date
