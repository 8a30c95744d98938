#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -2
for sig in 0 1; do
echo "== signal mode $sig"
MILAB200_CHAIN_SIGNAL=$sig timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 3 > $O/r2j8_timeline_sig$sig.txt 2>&1; head -8 $O/r2j8_timeline_sig$sig.txt
for cfg in "--mode chain" "--mode chain --tokens 8"; do
    tag=$(echo $cfg | tr -d ' -')
    MILAB200_CHAIN_SIGNAL=$sig timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline $cfg > $O/r2j8_bench_sig${sig}_$tag.json 2>$O/r2j8_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2j8_bench_sig${sig}_$tag.json')); print('sig=$sig $cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1))" || tail -3 $O/r2j8_bench_$tag.err
done
done
