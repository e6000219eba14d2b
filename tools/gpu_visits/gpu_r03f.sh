#!/bin/bash
# r03f (1 GPU): host time to issue the decode graph launches vs device time (is the captured loop launch-bound?)
TAG=${1:-r03f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/decode_timing.py $OUT/${TAG}_decode_timing.json 5 > $OUT/${TAG}_decode_timing.log 2>&1; echo "timing rc=$?"; cat $OUT/${TAG}_decode_timing.log | cut -c1-330
CT_PDL=2 timeout 600 python tools/decode_timing.py $OUT/${TAG}_decode_timing_pdl_off.json 3 > $OUT/${TAG}_decode_timing_pdl_off.log 2>&1; echo "timing (PDL off) rc=$?"; cat $OUT/${TAG}_decode_timing_pdl_off.log | cut -c1-330
nproc; grep -m1 "model name" /proc/cpuinfo
date
