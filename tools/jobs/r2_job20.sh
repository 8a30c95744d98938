#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -2
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 > $O/r2j20_timeline_gemma.txt 2>&1; head -9 $O/r2j20_timeline_gemma.txt | cut -c1-150
for cfg in "--mode chain --workload gemma4-12b-mlp-fp4" "--mode launches --workload gemma4-12b-mlp-fp4" "--mode chain --workload gemma4-12b-mlp-fp4 --tokens 2" "--mode chain --workload llama3-70b-mlp-fp4" "--mode chain" "--mode chain --tokens 8"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2j20_bench_$tag.json 2>$O/r2j20_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2j20_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'])" || tail -3 $O/r2j20_bench_$tag.err
done
