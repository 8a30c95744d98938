#!/bin/bash
# session k, job 19: chain kernel of commit 25417f2 (FP4 scales one unit ahead inside a Linear) vs HEAD (two units ahead across Linears): Llama-70B and Gemma FP4 chains
set -u
O=gpurun_out; mkdir -p $O
for lib in new c254; do
    if [ $lib = c254 ]; then export MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_chain25417f2.so; else unset MILAB200_LIB; fi
    for cfg in "--workload llama3-70b-mlp-fp4 --mode chain" "--workload gemma4-12b-mlp-fp4 --mode chain" "--workload llama3-70b-mlp-fp4 --mode chain --tokens 2"; do
    tag=$(echo $cfg | tr -d ' -')_$lib
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k19_bench_$tag.json 2>$O/r2k19_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k19_bench_$tag.json')); print('$lib $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k19_bench_$tag.err
    done
done
