#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q > $O/r2j6_pytest_chain.log 2>&1; tail -3 $O/r2j6_pytest_chain.log
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 4 > $O/r2j6_timeline.txt 2>&1; head -12 $O/r2j6_timeline.txt
for cfg in "--mode chain" "--mode chain --tokens 8" "--mode chain --workload llama3-70b-mlp-fp4 --tokens 4" "--mode launches --workload llama3-70b-mlp-fp4 --tokens 4"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $cfg > $O/r2j6_bench_$tag.json 2>$O/r2j6_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2j6_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))" || tail -3 $O/r2j6_bench_$tag.err
done
