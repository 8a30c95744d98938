#!/bin/bash
# session k, job 27: packed FP32 pairs (FFMA2) in the 16-token epilogue: parity + per-shape times (before: r2k3 look-ahead 0 lines)
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemv.py tests/test_gpu_glu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/perf_shapes.py --fmt fp4 --m 16 > $O/r2k27_fp4_m16.jsonl 2>$O/r2k27_err.txt
timeout 300 python tools/perf_shapes.py --fmt fp8 --m 16 >> $O/r2k27_fp4_m16.jsonl 2>>$O/r2k27_err.txt
python -c "
import json
for l in open('$O/r2k27_fp4_m16.jsonl'):
    d=json.loads(l); print(d['shape'],d['M'],d['us'],d['GBps'],d['kernel'])"
for cfg in "--workload gemma4-12b-mlp-fp4 --tokens 16" "--tokens 16"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k27_bench_$tag.json 2>$O/r2k27_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k27_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k27_bench_$tag.err
done
