#!/bin/bash
# session k, job 12: RMSNorm prologue with all loads of the reduction in flight: parity + bench of the fused MLP-block stacks
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_rmsnorm.py -x -q -m gpu > $O/r2k14_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/r2k14_pytest.txt
for cfg in "--workload gemma4-12b-mlp-fp4 --fuse-gate-up" "--fuse-gate-up" "--fuse-gate-up --tokens 16" "--workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 16" \
           "--fuse-gate-up --tokens 8" "--workload llama3-70b-mlp-fp4 --fuse-gate-up"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k14_bench_$tag.json 2>$O/r2k14_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k14_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k14_bench_$tag.err
done
