#!/bin/bash
# session m, job 9: host-buffer pass of the stack as one graph (H2D + launches + D2H): parity, then the bench's e2e figure
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "host_buffer or graph_replay" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/r2m9_bench.json 2>$O/r2m9_bench.err; tail -2 $O/r2m9_bench.err
python -c "
import json; d=json.load(open('$O/r2m9_bench.json')); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e'], d['clocks'])"
