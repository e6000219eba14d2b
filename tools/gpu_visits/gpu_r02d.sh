#!/bin/bash
# r02d (2 GPUs): VMM + NVLS multicast comm, DDP parity, graph-captured DDP step; plus the 1-GPU parity suite
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export CT_COMM_TIMEOUT_S=60
echo "== ddp_check (full, 2 GPUs)"; date
timeout 600 $TR --master-port 29531 tools/ddp_check.py --out $OUT/${TAG}_ddp_check_w2.json > $OUT/${TAG}_ddp_check.log 2>&1; echo "ddp_check rc=$?"
tail -3 $OUT/${TAG}_ddp_check.log | cut -c1-3000
echo "== bench N=2"; date
port=29540
for cfg in "default:" "nograph:--no-graph" "nccl:--comm nccl --no-graph"; do
  name=${cfg%%:*}; flags=${cfg#*:}
  port=$((port+1))
  timeout 400 $TR --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table $flags > $OUT/${TAG}_bench_n2_$name.json 2> $OUT/${TAG}_bench_n2_$name.err; echo "bench $name rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $OUT/${TAG}_bench_n2_$name.json | head -3; tail -2 $OUT/${TAG}_bench_n2_$name.err | cut -c1-400
done
echo "== 1-GPU: attention tests, parity shapes, bench with kernel table"; date
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > $OUT/${TAG}_attn_tests.log 2>&1; echo "attn rc=$?"; tail -3 $OUT/${TAG}_attn_tests.log
timeout 900 python -m pytest tests/test_gpu_parity_shapes.py tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_parity_tests.log 2>&1; echo "parity rc=$?"; tail -8 $OUT/${TAG}_parity_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -1; tail -2 $OUT/${TAG}_bench.err
date
