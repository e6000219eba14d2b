#!/bin/bash
# r02q (4 GPUs): multi-GPU parity (world 2 and 4, incl. BERT with dropout through the DDP wrapper), configs[4] as
# specified (bert-base 64 x 512 per GPU, 4 x DDP, dropout 0.1), Bloom SFT at N=4
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | head -8
echo "== multi-GPU parity tests"; date
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $OUT/${TAG}_multi_tests.log 2>&1; echo "multi rc=$?"; tail -6 $OUT/${TAG}_multi_tests.log | cut -c1-300
echo "== ddp_check world 4 (full: bandwidth table too)"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py --out $OUT/${TAG}_ddp_check_w4.json > $OUT/${TAG}_ddp_check_w4.log 2>&1; echo "ddp_check rc=$?"; tail -3 $OUT/${TAG}_ddp_check_w4.log | cut -c1-600
echo "== bert_cls at N=4"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload bert_cls --gpus 4 --steps 10 --warmup 3 > $OUT/${TAG}_bench_bert_cls_4gpu.json 2> $OUT/${TAG}_bench_bert_cls_4gpu.err; echo "bert_cls N=4 rc=$?"; tail -2 $OUT/${TAG}_bench_bert_cls_4gpu.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_bert_cls_4gpu.json | head -3
echo "== bloom_sft at N=4"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > $OUT/${TAG}_bench_4gpu.json 2> $OUT/${TAG}_bench_4gpu.err; echo "bloom N=4 rc=$?"; tail -2 $OUT/${TAG}_bench_4gpu.err | cut -c1-300
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_4gpu.json | head -3
date
