#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for la in 0 4 8; do
  echo "== lookahead $la"
  MILAB200_CHAIN_L2_LOOKAHEAD=$la timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 4 > $O/r2j4_timeline_la$la.txt 2>&1; head -9 $O/r2j4_timeline_la$la.txt
  for cfg in "--mode chain" "--mode chain --fuse-gate-up" "--mode chain --tokens 8"; do
    tag=$(echo $cfg | tr -d ' -')
    MILAB200_CHAIN_L2_LOOKAHEAD=$la timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $cfg > $O/r2j4_bench_la${la}_$tag.json 2>$O/r2j4_bench_la${la}_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2j4_bench_la${la}_$tag.json')); print('la=$la $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))" || tail -3 $O/r2j4_bench_la${la}_$tag.err
  done
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mode launches --fuse-gate-up > $O/r2j4_bench_launches_fuse.json 2>$O/r2j4_bench_launches_fuse.err
python -c "import json,sys; d=json.load(open('$O/r2j4_bench_launches_fuse.json')); print('launches fuse', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))" || tail -3 $O/r2j4_bench_launches_fuse.err
