#!/bin/bash
# r02u (1 GPU): programmatic dependent launch on the hot kernels: whole suite, then bench / decode bench with PDL on and off
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
echo "== whole GPU suite (PDL on)"; date
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1; echo "suite rc=$?"; tail -4 $OUT/${TAG}_tests.log | cut -c1-250
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log | cut -c1-200
echo "== bench PDL on / off"; date
timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench_pdl_on.json 2> $OUT/${TAG}_bench_pdl_on.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_pdl_on.json | head -2
CT_PDL=2 timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline > $OUT/${TAG}_bench_pdl_off.json 2> $OUT/${TAG}_bench_pdl_off.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_bench_pdl_off.json | head -2
echo "== decode bench PDL on / off"; date
timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 --no-eager-baseline > $OUT/${TAG}_decode_pdl_on.json 2> $OUT/${TAG}_decode_pdl_on.err; echo "rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_decode_pdl_on.json | head -2
CT_PDL=2 timeout 600 python bench.py --workload gpt2_decode --steps 3 --warmup 3 --no-eager-baseline > $OUT/${TAG}_decode_pdl_off.json 2> $OUT/${TAG}_decode_pdl_off.err; echo "rc=$?"
grep -o '"value": [0-9.]*, "unit"' $OUT/${TAG}_decode_pdl_off.json | head -2
date
