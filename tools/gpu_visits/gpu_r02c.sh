#!/bin/bash
# r02c: forward attention generation 4 + fused dQ convert, parity at real shapes, new bench line
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== attention tests"; date
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > $OUT/${TAG}_attn_tests.log 2>&1; echo "attn rc=$?"; tail -4 $OUT/${TAG}_attn_tests.log
echo "== attention A/B"; date
timeout 300 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab rc=$?"
cut -c1-420 $OUT/${TAG}_ab_attn.jsonl
echo "== parity at real shapes"; date
timeout 900 python -m pytest tests/test_gpu_parity_shapes.py -m gpu -q > $OUT/${TAG}_parity_tests.log 2>&1; echo "parity rc=$?"; tail -12 $OUT/${TAG}_parity_tests.log
echo "== bench (graph default, eager baseline, per-kernel table)"; date
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench.json | head -3; tail -3 $OUT/${TAG}_bench.err
CT_ATTN_FWD_IMPL=3 CT_ATTN_DQ_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-kernel-table > $OUT/${TAG}_bench_oldattn.json 2> /dev/null; echo "bench old attention rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_oldattn.json | head -1
echo "== whole GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_shapes.py > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -5 $OUT/${TAG}_tests.log
date
