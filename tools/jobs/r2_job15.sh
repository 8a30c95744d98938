#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for ll in 1 0; do
echo "== handoff=$ll"
MILAB200_CHAIN_HANDOFF=$ll timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 3 > $O/r2j15_timeline_ll$ll.txt 2>&1; sed -n 1,8p $O/r2j15_timeline_ll$ll.txt | cut -c1-150
for cfg in "--mode chain" "--mode chain --tokens 2"; do
    tag=$(echo $cfg | tr -d ' -')
    MILAB200_CHAIN_HANDOFF=$ll timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2j15_bench_ll${ll}_$tag.json 2>$O/r2j15_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2j15_bench_ll${ll}_$tag.json')); print('ll=$ll $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))" || tail -3 $O/r2j15_bench_$tag.err
done
done
