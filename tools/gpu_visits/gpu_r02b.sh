#!/bin/bash
# r02b: new real-shape parity tests + whole GPU suite, CUDA-graph step after the autograd-edge fix, multicast probe.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== multicast probe"; date
timeout 120 python tools/probe_multicast.py $TAG > /dev/null 2>&1; echo "probe rc=$?"
grep -i "MULTICAST\|FABRIC" $OUT/${TAG}_multicast_probe.txt | head
echo "== new parity tests"; date
timeout 900 python -m pytest tests/test_gpu_parity_shapes.py -m gpu -q -x > $OUT/${TAG}_parity_tests.log 2>&1; echo "parity rc=$?"; tail -15 $OUT/${TAG}_parity_tests.log
echo "== graph"; date
export TORCH_SHOW_CPP_STACKTRACES=1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --graph > $OUT/${TAG}_bench_graph.json 2> $OUT/${TAG}_bench_graph.err; echo "bench --graph rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_graph.json | head -1; tail -3 $OUT/${TAG}_bench_graph.err
unset TORCH_SHOW_CPP_STACKTRACES
CT_FUSED_LM_STATS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --graph > $OUT/${TAG}_bench_graph_lmstats.json 2> /dev/null; echo "bench --graph lmstats rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_graph_lmstats.json | head -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -1
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_shapes.py > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -5 $OUT/${TAG}_tests.log
date
