#!/bin/bash
# r02e (2 GPUs): ddp_check (fixed), DDP bench variants; 1-GPU: attention A/B (occupancy), parity, bench with CUPTI table
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export CT_COMM_TIMEOUT_S=60
echo "== ddp_check (full, 2 GPUs)"; date
timeout 600 $TR --master-port 29531 tools/ddp_check.py --out $OUT/${TAG}_ddp_check_w2.json > $OUT/${TAG}_ddp_check.log 2>&1; echo "ddp_check rc=$?"
grep -v "^W1017\|^\[W" $OUT/${TAG}_ddp_check.log | tail -3 | cut -c1-2500
echo "== bench N=2"; date
port=29540
run() { name=$1; shift; port=$((port+1));
  env "$@" timeout 400 $TR --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table --no-eager-baseline $FLAGS > $OUT/${TAG}_bench_n2_$name.json 2> $OUT/${TAG}_bench_n2_$name.err; echo "bench $name rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $OUT/${TAG}_bench_n2_$name.json | head -1; grep -o '"ddp_nvls": [a-z]*' $OUT/${TAG}_bench_n2_$name.json | head -1; }
FLAGS="" run default CT_X=0
FLAGS="" run nvls_off CT_DDP_NVLS=0
FLAGS="" run ctas8 CT_DDP_CTAS=8
FLAGS="" run ctas32 CT_DDP_CTAS=32
FLAGS="--no-graph" run nograph CT_X=0
FLAGS="--comm nccl --no-graph" run nccl CT_X=0
FLAGS="" run skipcomm CT_DDP_SKIP_COMM=1
echo "== 1-GPU"; date
timeout 300 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab rc=$?"
cut -c1-330 $OUT/${TAG}_ab_attn.jsonl | head -4
timeout 900 python -m pytest tests/test_gpu_parity_shapes.py tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_parity_tests.log 2>&1; echo "parity rc=$?"; tail -6 $OUT/${TAG}_parity_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -1; tail -2 $OUT/${TAG}_bench.err
date
