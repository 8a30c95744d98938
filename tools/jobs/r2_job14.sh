#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for coop in 0 1; do
echo "== cooperative=$coop, metrics-only"
MILAB200_CHAIN_COOPERATIVE=$coop timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_chain -c 2 --csv --log-file $O/r2j14_chain_time_coop$coop.csv \
    python tools/ncu_chain_case.py 2 1 > $O/r2j14_ncu_time_coop$coop.log 2>&1
echo "rc=$?"; tail -3 $O/r2j14_chain_time_coop$coop.csv
done
echo "== cooperative=0 full"
MILAB200_CHAIN_COOPERATIVE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_chain -s 2 -c 1 -f -o $O/r2j14_prof_chain_fp8_m1 \
    python tools/ncu_chain_case.py 4 1 > $O/r2j14_ncu_chain.log 2>&1
echo "rc=$?"; tail -4 $O/r2j14_ncu_chain.log
MILAB200_CHAIN_COOPERATIVE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2j14_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2j14_ncu_launches.log 2>&1
echo "launch list rc=$?"; grep -c decode_chain $O/r2j14_bench_launches.csv
