#!/bin/bash
# session k, job 23: FP4 group scales of the units behind pulled into L2 by the epilogue (prefetch.global.L2): chain and launched packed-nibble kernels, A/B
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -2
for ah in 0 2 3; do
    for cfg in "--workload llama3-70b-mlp-fp4 --mode chain" "--workload gemma4-12b-mlp-fp4 --mode chain" "--workload gemma4-12b-mlp-fp4 --mode launches" "--workload llama3-70b-mlp-fp4 --mode launches"; do
    tag=$(echo $cfg | tr -d ' -')_ah$ah
    MILAB200_CHAIN_SCALE_L2_AHEAD=$ah MILAB200_MX4_SCALE_L2_AHEAD=$ah timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k23_bench_$tag.json 2>$O/r2k23_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k23_bench_$tag.json')); print('ahead=$ah $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k23_bench_$tag.err
    done
done
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 > $O/r2k23_timeline_gemma.txt 2>&1; head -11 $O/r2k23_timeline_gemma.txt | cut -c1-135
