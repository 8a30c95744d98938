#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q > $O/r2j21_pytest_chain.log 2>&1; tail -6 $O/r2j21_pytest_chain.log
for cfg in "--mode chain --tokens 16" "--mode launches --tokens 16" "--mode chain --tokens 12" "--mode chain --tokens 16 --workload llama3-70b-mlp-fp4" "--mode launches --tokens 16 --workload llama3-70b-mlp-fp4" "--mode chain --tokens 16 --workload gemma4-12b-mlp-fp4" "--mode launches --tokens 16 --workload gemma4-12b-mlp-fp4"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2j21_bench_$tag.json 2>$O/r2j21_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2j21_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['reasons'])" || tail -3 $O/r2j21_bench_$tag.err
done
