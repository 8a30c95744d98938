#!/bin/bash
# session m, job 14: ncu --set full of the 16-token FP8 decode kernel (final epilogue) on the Llama-8B gate shape
O=gpurun_out; mkdir -p $O
python tools/ncu_case.py fp8 4096 14336 16 > $O/r2m14_plain.log 2>&1 &&
timeout 120 ncu --set full --clock-control none --import-source on -k regex:decode_tc -s 3 -c 1 -f -o $O/r2m14_prof_fp8_gate_m16 \
    python tools/ncu_case.py fp8 4096 14336 16 > $O/r2m14_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/r2m14_plain.log
