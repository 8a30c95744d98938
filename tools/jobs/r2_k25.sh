#!/bin/bash
# session k, job 25: chain, E4M3-plane schemes at M <= 2: epilogue reads only the live tokens' columns — A/B on the headline and M = 2
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for sm in 1 0; do
    for cfg in "--mode chain" "--mode chain --tokens 2"; do
    tag=$(echo $cfg | tr -d ' -')_sm$sm$rep
    MILAB200_CHAIN_SMALL_M_EPILOGUE=$sm timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k25_bench_$tag.json 2>$O/r2k25_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k25_bench_$tag.json')); print('small_m=$sm rep$rep $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k25_bench_$tag.err
    done
done; done
