#!/bin/bash
# session m, job 4: gate|up + GLU fused into the batched kernel's epilogue (M > 32): parity, then fused vs two-step at M = 2048
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_glu.py -x -q -m gpu -k "batched" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_rmsnorm.py tests/test_gpu_prefill.py tests/test_gpu_glu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/perf_glu.py 2048 > $O/r2m4_perf_glu_m2048.jsonl 2>$O/r2m4_perf_glu.err; tail -2 $O/r2m4_perf_glu.err
python - <<'P'
import json
for l in open('gpurun_out/r2m4_perf_glu_m2048.jsonl'):
    d=json.loads(l); print(d['case'], d['M'], d['fused_us'], d['two_step_us'], d['speedup'], d['kernels'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
P
