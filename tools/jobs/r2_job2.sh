#!/bin/bash
# round 2, GPU job 2: chained decode kernel — parity, then bench A/B against per-Linear launches
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q > $O/r2j2_pytest_chain.log 2>&1; echo "pytest rc=$?" >> $O/r2j2_pytest_chain.log
tail -15 $O/r2j2_pytest_chain.log
for cfg in "--mode launches" "--mode chain" "--mode launches --fuse-gate-up" "--mode chain --fuse-gate-up" "--mode chain --tokens 8" "--mode launches --tokens 8" "--mode chain --tokens 16" "--mode launches --tokens 16"; do
  tag=$(echo $cfg | tr -d ' -')
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $cfg > $O/r2j2_bench_$tag.json 2>$O/r2j2_bench_$tag.err
  python -c "import json,sys; d=json.load(open('$O/r2j2_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['roofline']['frac'],4), round(d['e2e']['value'],1), d['launches_per_step'])" || tail -3 $O/r2j2_bench_$tag.err
done
