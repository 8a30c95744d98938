#!/bin/bash
# session k, job 10: RMSNorm -> gate|up -> GLU as one launch (incl. the packed-nibble FP4 kernel), FP4 one-plane batched mode: parity + bench
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_rmsnorm.py tests/test_gpu_glu.py -x -q -m gpu > $O/r2k11_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/r2k11_pytest.txt
for cfg in "--workload gemma4-12b-mlp-fp4 --fuse-gate-up" "--fuse-gate-up" "--fuse-gate-up --tokens 16" "--workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 16" \
           "--fuse-gate-up --tokens 8" "--workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 2" "--workload llama3-70b-mlp-fp4 --fuse-gate-up"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k11_bench_$tag.json 2>$O/r2k11_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k11_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k11_bench_$tag.err
done
MILAB200_PREFILL_ACT_PLANES=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --workload gemma4-12b-mlp-fp4 --tokens 2048 > $O/r2k11_bench_fp4_a8.json 2>$O/r2k11_bench_fp4_a8.err
python -c "import json,sys; d=json.load(open('$O/r2k11_bench_fp4_a8.json')); print('fp4 one-plane', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['achieved'],1), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k11_bench_fp4_a8.err
