for c in 16 48; do
    echo "== CTAS=$c"
    CT_DDP_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
echo "== nccl"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --comm nccl 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
