#!/bin/bash
# session k, job 7: second epilogue set at 9..16 tokens (pre-split path): parity + A/B
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemv.py tests/test_gpu_glu.py tests/test_gpu_fuzz.py tests/test_rmsnorm.py tests/test_gpu_linear.py -x -q -m gpu > $O/r2k7_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/r2k7_pytest.txt
for sets in 1 2; do
  echo "== epilogue sets $sets"
  MILAB200_EPILOGUE_SETS=$sets timeout 300 python tools/perf_shapes.py --fmt fp4 --m 16 > $O/r2k7_fp4_sets$sets.jsonl 2>$O/r2k7_err.txt
  MILAB200_EPILOGUE_SETS=$sets timeout 300 python tools/perf_shapes.py --fmt fp8 --m 16 >> $O/r2k7_fp4_sets$sets.jsonl 2>>$O/r2k7_err.txt
  python -c "
import json
for l in open('$O/r2k7_fp4_sets$sets.jsonl'):
    d=json.loads(l); print(d['shape'],d['M'],d['us'],d['GBps'],d['kernel'])"
done
tail -3 $O/r2k7_err.txt
for sets in 1 2; do
for cfg in "--mode launches --workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 16" "--mode launches --fuse-gate-up --tokens 16" "--mode launches --tokens 16"; do
    tag=$(echo $cfg | tr -d ' -')_sets$sets
    MILAB200_EPILOGUE_SETS=$sets timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k7_bench_$tag.json 2>$O/r2k7_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k7_bench_$tag.json')); print('sets=$sets $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k7_bench_$tag.err
done; done
