#!/bin/bash
# session k, job 26: four-way k split in the chain (pair over DSMEM + pair over L2 tagged words): parity + bench
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -3
for cfg in "--workload gemma4-12b-mlp-fp4 --mode chain" "--workload llama3-70b-mlp-fp4 --mode chain" "--mode chain" "--workload gemma4-12b-mlp-fp4 --mode chain --tokens 2"; do
    for mp in 0 2; do
    tag=$(echo $cfg | tr -d ' -')_maxp$mp
    MILAB200_CHAIN_MAX_SPLITK=$mp timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k26_bench_$tag.json 2>$O/r2k26_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k26_bench_$tag.json')); print('max_splitk=$mp $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k26_bench_$tag.err
    done
done
MILAB200_CHAIN_MAX_SPLITK=4 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --mode chain > $O/r2k26_bench_fp8_maxp4.json 2>$O/r2k26_bench_fp8_maxp4.err
python -c "import json,sys; d=json.load(open('$O/r2k26_bench_fp8_maxp4.json')); print('max_splitk=4 headline', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 > $O/r2k26_timeline_gemma.txt 2>&1; head -11 $O/r2k26_timeline_gemma.txt | cut -c1-135
