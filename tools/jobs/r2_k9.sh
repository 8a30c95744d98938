#!/bin/bash
# session k, job 9: FP4 batched path with summed planes (after the commit-flag fix): parity + bench
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_prefill.py -x -q -m gpu > $O/r2k9_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/r2k9_pytest.txt
for sum in 1 0; do
for cfg in "--workload gemma4-12b-mlp-fp4 --tokens 2048" "--workload llama3-70b-mlp-fp4 --tokens 2048" "--workload gemma4-12b-mlp-fp4 --tokens 512"; do
    tag=$(echo $cfg | tr -d ' -')_sum$sum
    MILAB200_PREFILL_FP4_SUM=$sum timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k9_bench_$tag.json 2>$O/r2k9_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k9_bench_$tag.json')); print('sum=$sum $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k9_bench_$tag.err
done; done
