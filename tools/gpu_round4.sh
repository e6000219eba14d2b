#!/bin/bash
# GPU visit r01g: attention backward v4 (drain warpgroup) + cluster-resident cross entropy: parity, A/B, bench.
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
echo "== parity"; date
timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention and v4" > $OUT/${TAG}_attn_tests.log 2>&1; echo "attn v4 tests rc=$?"
tail -3 $OUT/${TAG}_attn_tests.log
timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cross_entropy" > $OUT/${TAG}_ce_tests.log 2>&1; echo "ce tests rc=$?"
tail -8 $OUT/${TAG}_ce_tests.log
echo "== A/B"; date
timeout 150 python tools/kernel_ab.py attn > $OUT/${TAG}_ab_attn.jsonl 2> $OUT/${TAG}_ab_attn.err; echo "ab attn rc=$?"
grep bloom_bench $OUT/${TAG}_ab_attn.jsonl | cut -c1-80,230-420
timeout 150 python tools/kernel_ab.py ce > $OUT/${TAG}_ab_ce.jsonl 2> $OUT/${TAG}_ab_ce.err; echo "ab ce rc=$?"
cat $OUT/${TAG}_ab_ce.jsonl; tail -2 $OUT/${TAG}_ab_ce.err
for cfg in "4 1" "5 1" "5 2"; do
  set -- $cfg
  echo "== bench ATTN_BWD_IMPL=$1 CE_IMPL=$2"; date
  CT_ATTN_BWD_IMPL=$1 CT_CE_IMPL=$2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$1_$2.json 2> $OUT/${TAG}_bench_$1_$2.err; echo "bench rc=$?"
  grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_$1_$2.json | head -1
  grep -o '"loss": [0-9.]*' $OUT/${TAG}_bench_$1_$2.json | head -1
done
date
