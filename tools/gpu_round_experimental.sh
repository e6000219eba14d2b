#!/bin/bash
# First GPU visit for the variants written after round 1's GPU budget was spent (compiled, never run):
#   attention backward v5 (persistent, ATTN_BWD_IMPL=6) and v6 (sixteen compute warps, ATTN_BWD_IMPL=7), attention forward v3 (lazy maximum + per-panel P hand-over,
#   ATTN_FWD_IMPL=2), tanh-GELU derivative saved by the forward (CT_SAVE_ACT_GRAD=1; parity already green, timing
#   pending), LM-head GEMM with softmax statistics + one-pass loss (CT_FUSED_LM_STATS=1), CUDA-graph training step (graphs.GraphedTrainStep, bench.py --graph), DDP copy-engine transport (--comm ce; needs 2 GPUs: run with `gpurun --gpus 2` and N2=1).
# Every step runs in its own process under `timeout` (mbarrier waits trap after ~2 s, so a protocol bug shows up as a
# failed test, not a hung box).   gpurun --timeout 900 -- 'bash tools/gpu_round_experimental.sh r02a'
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
export CT_TEST_EXPERIMENTAL=1
echo "== parity: experimental attention variants"; date
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention and v5" > $OUT/${TAG}_attn_v5_tests.log 2>&1; echo "bwd v5 rc=$?"; tail -3 $OUT/${TAG}_attn_v5_tests.log
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention and v6" > $OUT/${TAG}_attn_v6_tests.log 2>&1; echo "bwd v6 rc=$?"; tail -3 $OUT/${TAG}_attn_v6_tests.log
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention and f3" > $OUT/${TAG}_attn_f3_tests.log 2>&1; echo "fwd v3 rc=$?"; tail -3 $OUT/${TAG}_attn_f3_tests.log
echo "== CUDA-graph training step"; date
timeout 180 python -m pytest tests/test_gpu_models.py -m gpu -q -k "graphed" > $OUT/${TAG}_graph_tests.log 2>&1; echo "graph rc=$?"; tail -3 $OUT/${TAG}_graph_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --graph > $OUT/${TAG}_bench_graph.json 2> $OUT/${TAG}_bench_graph.err; echo "bench --graph rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_graph.json | head -1; tail -2 $OUT/${TAG}_bench_graph.err
echo "== fused LM-head statistics"; date
timeout 180 python -m pytest tests -m gpu -q -k "row_stats or fused_lm_head" > $OUT/${TAG}_lmstats_tests.log 2>&1; echo "lm stats rc=$?"; tail -3 $OUT/${TAG}_lmstats_tests.log
echo "== attention A/B (v2 / v3 / v4 / v5 backward, f3 forward)"; date
timeout 200 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab rc=$?"
grep bloom_bench $OUT/${TAG}_ab_attn.jsonl | cut -c1-90,230-420
echo "== bench per knob"; date
for cfg in "CT_X=0" "CT_ATTN_BWD_IMPL=6" "CT_ATTN_BWD_IMPL=7" "CT_ATTN_FWD_IMPL=2" "CT_SAVE_ACT_GRAD=1" "CT_FUSED_LM_STATS=1" "CT_ATTN_BWD_IMPL=5"; do
  name=$(echo "$cfg" | tr '=' '_')
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> /dev/null; echo "$cfg rc=$?"
  grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_$name.json | head -1
  grep -o '"loss": [0-9.]*' $OUT/${TAG}_bench_$name.json | head -1
done
if [ -n "$N2" ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  echo "== 2 GPUs: ddp_check incl. ce, bench p2p / ce / nccl"; date
  timeout 300 $TR --master-port 29531 tools/ddp_check.py > $OUT/${TAG}_ddp_check.log 2>&1; echo "ddp_check rc=$?"
  tail -2 $OUT/${TAG}_ddp_check.log | cut -c1-1800
  port=29540
  for comm in p2p ce nccl; do
    port=$((port+1))
    timeout 240 $TR --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --comm $comm > $OUT/${TAG}_bench_n2_$comm.json 2> $OUT/${TAG}_bench_n2_$comm.err; echo "bench $comm rc=$?"
    grep -o '"ms_per_step": [0-9.]*' $OUT/${TAG}_bench_n2_$comm.json | head -1
  done
fi
date
