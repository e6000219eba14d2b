#!/bin/bash
# GPU visit r01j: split-K rule + colsum overlap (A/B, parity), bench per configuration
TAG=${1:-r01j}
OUT=gpurun_out
mkdir -p $OUT
echo "== wgrad A/B"; date
timeout 200 python tools/kernel_ab.py wgrad > $OUT/${TAG}_ab_wgrad.jsonl 2> $OUT/${TAG}_ab_wgrad.err; echo "rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/%s_ab_wgrad.jsonl" % "TAGX".replace("TAGX", __import__("os").environ.get("TAG", "r01j"))):
    r = json.loads(l)
    print(r.get("case"), "rule", r.get("splitk_rule"), "ovl", r.get("overlap_colsum"),
          "us %.1f" % r.get("us_gemm_plus_colsum", -1), "err %.1e %.1e %.1e" % (r.get("err", -1), r.get("err_bias", -1), r.get("err_accumulate", -1)), r.get("error", ""))
PY
tail -2 $OUT/${TAG}_ab_wgrad.err
echo "== gemm / model parity on the new defaults"; date
timeout 300 python -m pytest tests -m gpu -q -x -k "gemm or linear or model or bloom or block" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -3 $OUT/${TAG}_tests.log
for cfg in "1 0" "0 0" "0 1"; do
  set -- $cfg
  echo "== bench GEMM_SPLITK=$1 OVERLAP_COLSUM=$2"; date
  CT_GEMM_SPLITK=$1 CT_OVERLAP_COLSUM=$2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$1_$2.json 2> $OUT/${TAG}_bench_$1_$2.err; echo "bench rc=$?"
  grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_$1_$2.json | head -1
  grep -o '"loss": [0-9.]*' $OUT/${TAG}_bench_$1_$2.json | head -1
  grep -o '"gemm_ms_per_step": [0-9.]*' $OUT/${TAG}_bench_$1_$2.json | head -1
done
date
