#!/bin/bash
# session k, job 15: ncu --set full of the summed-planes FP4 batched kernel (and the opt-in one-plane FP4 mode)
set -u
O=gpurun_out; mkdir -p $O
python tools/ncu_prefill_case.py fp4 3840 30720 2048 3 > $O/r2k15_plain_pf4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2k15_prof_prefill_fp4_sum_gateup_m2048 \
    python tools/ncu_prefill_case.py fp4 3840 30720 2048 3 > $O/r2k15_ncu_pf4.log 2>&1
echo "fp4 sum capture rc=$?"; tail -1 $O/r2k15_plain_pf4.log
MILAB200_PREFILL_ACT_PLANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2k15_prof_prefill_fp4_a8_gateup_m2048 \
    python tools/ncu_prefill_case.py fp4 3840 30720 2048 3 > $O/r2k15_ncu_pf4a8.log 2>&1
echo "fp4 a8 capture rc=$?"
MILAB200_PREFILL_FP4_SUM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2k15_prof_prefill_fp4_hilo_gateup_m2048 \
    python tools/ncu_prefill_case.py fp4 3840 30720 2048 3 > $O/r2k15_ncu_pf4hl.log 2>&1
echo "fp4 hi|lo capture rc=$?"
ls -la $O/*.ncu-rep | grep r2k15
