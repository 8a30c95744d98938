#!/bin/bash
# session m, job 1: (a) stream-K tile schedule of the batched kernel: parity + per-shape / bench A/B (MILAB200_PREFILL_STREAMK=0 vs 1);
# (b) deeper FP4 group-scale ring in the 9..16-token decode kernel: parity + per-shape A/B against the HEAD build
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_prefill.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_gemv.py tests/test_gpu_glu.py -x -q -m gpu 2>&1 | tail -2
for sk in 0 1; do for pl in 2 1; do
    MILAB200_PREFILL_STREAMK=$sk MILAB200_PREFILL_ACT_PLANES=$pl timeout 300 python tools/perf_prefill.py --fmt fp8,fp4 --m 2048 \
        > $O/r2m1_perf_prefill_sk${sk}_pl${pl}.jsonl 2>$O/r2m1_perf_prefill_sk${sk}_pl${pl}.err
    echo "== stream-K $sk, activation planes $pl"; python - <<P
import json
for l in open('$O/r2m1_perf_prefill_sk${sk}_pl${pl}.jsonl'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['shape'], d['us'], d['TFLOPs'], d['frac_bf16_peak'], d['kernel'], d['schedule'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
P
done; done
for sk in 0 1; do
  for cfg in "--tokens 2048" "--workload gemma4-12b-mlp-fp4 --tokens 2048"; do
    tag=$(echo $cfg | tr -d ' -')_sk$sk
    MILAB200_PREFILL_STREAMK=$sk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2m1_bench_$tag.json 2>$O/r2m1_bench_$tag.err
    python -c "import json; d=json.load(open('$O/r2m1_bench_$tag.json')); print('$tag', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2m1_bench_$tag.err
  done
done
MILAB200_PREFILL_ACT_PLANES=1 MILAB200_PREFILL_STREAMK=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --tokens 2048 > $O/r2m1_bench_a8_sk0.json 2>/dev/null
MILAB200_PREFILL_ACT_PLANES=1 MILAB200_PREFILL_STREAMK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --tokens 2048 > $O/r2m1_bench_a8_sk1.json 2>/dev/null
for t in a8_sk0 a8_sk1; do python -c "import json; d=json.load(open('$O/r2m1_bench_$t.json')); print('$t', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
# (b) decode M = 16 FP4: HEAD build vs this build
for lib in libmila_b200_linear_head.so libmila_b200_linear.so; do
    echo "== $lib"
    MILAB200_LIB=$PWD/mila_b200/$lib timeout 300 python tools/perf_shapes.py --fmt fp4 --m 16 > $O/r2m1_fp4_m16_$lib.jsonl 2>$O/r2m1_err.txt
    MILAB200_LIB=$PWD/mila_b200/$lib timeout 300 python tools/perf_shapes.py --fmt fp4 --m 12 >> $O/r2m1_fp4_m16_$lib.jsonl 2>>$O/r2m1_err.txt
    python -c "
import json
for l in open('$O/r2m1_fp4_m16_$lib.jsonl'):
    d=json.loads(l); print(d['shape'],d['M'],d['us'],d['GBps'],d['kernel'])"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --workload gemma4-12b-mlp-fp4 --tokens 16 > $O/r2m1_bench_gemma_m16.json 2>/dev/null
python -c "import json; d=json.load(open('$O/r2m1_bench_gemma_m16.json')); print('gemma fp4 M16', round(d['value'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
