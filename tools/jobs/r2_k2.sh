#!/bin/bash
# session k, job 2: fused gate|up + GLU stacks, per-Linear launches (the chain loses on them: r2k1 timeline)
set -u
O=gpurun_out; mkdir -p $O
for cfg in "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up" "--mode launches --workload gemma4-12b-mlp-fp4" \
           "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 16" \
           "--mode launches --fuse-gate-up --tokens 16" "--mode chain --fuse-gate-up" \
           "--mode launches --fuse-gate-up" "--mode launches" "--mode chain --fuse-gate-up --tokens 8" "--mode chain --workload llama3-70b-mlp-fp4 --fuse-gate-up" \
           "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 4" "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 8" "--mode launches --workload gemma4-12b-mlp-fp4 --tokens 8"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k2_bench_$tag.json 2>$O/r2k2_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k2_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k2_bench_$tag.err
done
