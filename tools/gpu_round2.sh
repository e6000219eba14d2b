#!/bin/bash
# GPU visit for the second-generation kernels: (1) the whole GPU suite on the first-generation variants,
# (2) per-family A/B with parity numbers (separate processes: a trapping kernel only loses its family),
# (3) the whole suite on the defaults, (4) bench on the best known-good configuration + on the old one,
# (5) ncu launch list of one step.   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r01d'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
echo "== safe-config tests"; date
[ -n "$SKIP_SAFE" ] || CT_LN_BWD_IMPL=1 CT_ATTN_FWD_IMPL=1 CT_ATTN_BWD_IMPL=1 CT_GEMM_EPI_IMPL=1 timeout 600 \
  python -m pytest tests -m gpu -q -k "not v2" > $OUT/${TAG}_tests_safe.log 2>&1; echo "safe rc=$?"
tail -3 $OUT/${TAG}_tests_safe.log
echo "== kernel A/B"; date
for fam in ln gemm attn; do
  timeout 300 python tools/kernel_ab.py $fam > $OUT/${TAG}_ab_$fam.jsonl 2> $OUT/${TAG}_ab_$fam.err; echo "ab $fam rc=$?"
  cat $OUT/${TAG}_ab_$fam.jsonl | cut -c1-400
  tail -3 $OUT/${TAG}_ab_$fam.err
done
echo "== default-config tests"; date
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_tests.log | tail -20
echo "== bench"; date
eval "$(python tools/pick_safe_env.py $OUT/${TAG}_ab_ln.jsonl $OUT/${TAG}_ab_gemm.jsonl $OUT/${TAG}_ab_attn.jsonl)"
env | grep "^CT_" > $OUT/${TAG}_bench_env.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 2500 $OUT/${TAG}_bench.json
echo "== launches"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/step_prof.py > $OUT/${TAG}_launches.log 2>&1; echo "launches rc=$?"
echo "== bench, first-generation kernels"; date
[ -n "$SKIP_SAFE" ] || CT_LN_BWD_IMPL=1 CT_ATTN_FWD_IMPL=1 CT_ATTN_BWD_IMPL=1 CT_GEMM_EPI_IMPL=1 timeout 300 \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_v1.json 2> $OUT/${TAG}_bench_v1.err; echo "bench v1 rc=$?"
grep -o '"value": [0-9.]*' $OUT/${TAG}_bench_v1.json | head -1
date
