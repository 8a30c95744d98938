#!/bin/bash
# session k, job 1: gate|up as ONE Linear with the GLU epilogue (Mila's fc_gate_up dataflow) vs the three-Linear stacks, launches and chain
set -u
O=gpurun_out; mkdir -p $O
for cfg in "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up" "--mode chain --workload gemma4-12b-mlp-fp4 --fuse-gate-up" \
           "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 16" "--mode launches --workload gemma4-12b-mlp-fp4 --tokens 16" \
           "--mode launches --fuse-gate-up --tokens 16" "--mode launches --tokens 16" "--mode chain --fuse-gate-up" "--mode chain" \
           "--mode launches --fuse-gate-up" "--mode chain --fuse-gate-up --tokens 8" "--mode chain --workload llama3-70b-mlp-fp4 --fuse-gate-up" \
           "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 4" "--mode launches --workload gemma4-12b-mlp-fp4 --tokens 4"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k1_bench_$tag.json 2>$O/r2k1_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k1_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k1_bench_$tag.err
done
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 --fuse > $O/r2k1_timeline_gemma_fused.txt 2>&1; head -9 $O/r2k1_timeline_gemma_fused.txt | cut -c1-150
