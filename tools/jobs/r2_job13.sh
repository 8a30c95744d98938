#!/bin/bash
# round 2, GPU job 13: the whole GPU parity suite on the final kernels, goldens, ncu launch list + full captures
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2j13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j13_pytest.log; tail -4 $O/r2j13_pytest.log
timeout 300 python tools/make_golden.py $O/golden > $O/r2j13_golden.log 2>&1; tail -2 $O/r2j13_golden.log
# launch list of the default bench command
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2j13_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2j13_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2j13_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full capture: the chained kernel (4 layers = 12 Linears per launch)
python tools/ncu_chain_case.py 4 1 > $O/r2j13_plain_chain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_chain -s 2 -c 1 -f -o $O/r2j13_prof_chain_fp8_m1 \
    python tools/ncu_chain_case.py 4 1 > $O/r2j13_ncu_chain.log 2>&1
echo "chain capture rc=$?"; tail -2 $O/r2j13_plain_chain.log
# full captures: batched kernel, CTA pairs — FP8 exact, FP8 one-plane, FP4
python tools/ncu_prefill_case.py fp8 4096 14336 2048 3 > $O/r2j13_plain_pf8.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2j13_prof_prefill_fp8_gate_m2048 \
    python tools/ncu_prefill_case.py fp8 4096 14336 2048 3 > $O/r2j13_ncu_pf8.log 2>&1
echo "prefill fp8 capture rc=$?"
MILAB200_PREFILL_ACT_PLANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2j13_prof_prefill_fp8_a8_gate_m2048 \
    python tools/ncu_prefill_case.py fp8 4096 14336 2048 3 > $O/r2j13_ncu_pf8a8.log 2>&1
echo "prefill fp8 a8 capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2j13_prof_prefill_fp4_gateup_m2048 \
    python tools/ncu_prefill_case.py fp4 3840 30720 2048 3 > $O/r2j13_ncu_pf4.log 2>&1
echo "prefill fp4 capture rc=$?"
ls -la $O/*.ncu-rep | grep r2j13
