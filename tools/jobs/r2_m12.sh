#!/bin/bash
# session m, job 12: straight-line GLU evaluation in the decode epilogues (decode_tc at M > 2, chain): parity, then the M = 16 MLP-block line
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_glu.py tests/test_rmsnorm.py tests/test_gpu_chain.py -x -q -m gpu -k "glu or GLU" 2>&1 | tail -3
for cfg in "--tokens 16" "--tokens 8" "--workload gemma4-12b-mlp-fp4 --tokens 16"; do
timeout 200 python bench.py $cfg --fuse-gate-up --norm-fast --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $O/r2m12_bench.json 2>$O/r2m12_bench.err
python -c "
import json; d=json.load(open('$O/r2m12_bench.json')); print('$cfg mlp_block norm_fast', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2m12_bench.err
cat $O/r2m12_bench.json >> $O/r2m12_bench_lines.jsonl
done
