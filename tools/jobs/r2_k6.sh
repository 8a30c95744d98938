#!/bin/bash
# session k, job 6: role timeline of the 16-token decode kernel WITH the activation pre-pass (what sets its unit cadence?)
set -u
O=gpurun_out; mkdir -p $O
export MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_diag.so
export MILAB200_PROF_PRESPLIT=1
timeout 200 python tools/tc_timeline.py fp4 8192 28672 16 > $O/r2k6_tl_fp4_70bup_m16_ps.txt 2>&1
timeout 200 python tools/tc_timeline.py fp8 4096 14336 16 > $O/r2k6_tl_fp8_gate_m16_ps.txt 2>&1
timeout 200 python tools/tc_timeline.py fp4 3840 30720 16 > $O/r2k6_tl_fp4_gemma_gu_m16_ps.txt 2>&1
head -45 $O/r2k6_tl_fp4_70bup_m16_ps.txt | cut -c1-170
head -30 $O/r2k6_tl_fp8_gate_m16_ps.txt | cut -c1-170
