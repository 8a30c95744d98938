#!/bin/bash
# session k, job 24: one-token epilogue path of the packed-nibble kernels (x8 TMEM reads, one batch per unit): parity + bench
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_gemv.py tests/test_gpu_glu.py tests/test_rmsnorm.py -x -q -m gpu 2>&1 | tail -2
for cfg in "--workload llama3-70b-mlp-fp4 --mode chain" "--workload gemma4-12b-mlp-fp4 --mode chain" "--workload gemma4-12b-mlp-fp4 --mode launches" "--workload llama3-70b-mlp-fp4 --mode launches" "--workload gemma4-12b-mlp-fp4 --fuse-gate-up --norm-fast"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k24_bench_$tag.json 2>$O/r2k24_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k24_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k24_bench_$tag.err
done
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 > $O/r2k24_timeline_gemma.txt 2>&1; head -11 $O/r2k24_timeline_gemma.txt | cut -c1-135
