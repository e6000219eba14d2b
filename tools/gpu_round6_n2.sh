#!/bin/bash
# 2-GPU: where does the N=2 step time go? comm skipped / few / many all-reduce CTAs
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29520
for cfg in "CT_DDP_SKIP_COMM=1" "CT_DDP_CTAS=4" "CT_DDP_CTAS=8" "CT_DDP_CTAS=48" "CT_DDP_CTAS=8 CT_DDP_FINAL_CTAS=8"; do
  port=$((port+1))
  name=$(echo "$cfg" | tr ' =' '__')
  echo "== $cfg"; date
  env $cfg timeout 200 $TR --master-port $port bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err; echo "rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $OUT/${TAG}_$name.json | head -1
  grep -o '"gemm_ms_per_step": [0-9.]*' $OUT/${TAG}_$name.json | head -1
done
date
