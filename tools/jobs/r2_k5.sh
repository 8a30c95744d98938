#!/bin/bash
# session k, job 5: role timeline of the unpacked-FP4 decode kernel at 4 / 8 / 16 tokens on the largest matrix (what sets the unit cadence?)
set -u
O=gpurun_out; mkdir -p $O
export MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_diag.so
for m in 4 8 16; do
  timeout 200 python tools/tc_timeline.py fp4 8192 28672 $m > $O/r2k5_tl_fp4_70bup_m$m.txt 2>&1
done
timeout 200 python tools/tc_timeline.py fp8 4096 14336 8 > $O/r2k5_tl_fp8_gate_m8.txt 2>&1
head -40 $O/r2k5_tl_fp4_70bup_m4.txt | cut -c1-170
