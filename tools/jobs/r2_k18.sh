#!/bin/bash
# session k, job 17: Llama-70B FP4 chain, M = 1: this session's library vs the library of commit 1482a27 on one box (the default bench line of job 16 had 732 tok/s where earlier boxes measured 946)
set -u
O=gpurun_out; mkdir -p $O
for rep in 1 2; do
for lib in new old; do
    if [ $lib = old ]; then export MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_r2j23.so; else unset MILAB200_LIB; fi
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --workload llama3-70b-mlp-fp4 > $O/r2k18_bench_70b_$lib$rep.json 2>$O/r2k18_bench_70b_$lib$rep.err
    python -c "import json,sys; d=json.load(open('$O/r2k18_bench_70b_$lib$rep.json')); print('$lib $rep', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k18_bench_70b_$lib$rep.err
done; done
unset MILAB200_LIB
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --workload llama3-70b-mlp-fp4 --mode launches > $O/r2k18_bench_70b_launches.json 2>$O/r2k18_bench_70b_launches.err
python -c "import json,sys; d=json.load(open('$O/r2k18_bench_70b_launches.json')); print('launches', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'])"
timeout 300 python tools/chain_timeline.py llama3-70b-mlp-fp4 1 3 > $O/r2k18_timeline_70b.txt 2>&1; head -12 $O/r2k18_timeline_70b.txt | cut -c1-170
