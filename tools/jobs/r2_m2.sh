#!/bin/bash
# session m, job 2: stream-K TAIL (whole-tile waves + the last wave's k blocks over <= 4 pairs per tile): parity + A/B;
# ncu source-level capture of the 16-token FP4 decode kernel (what is the ~1100 clk every fourth unit?)
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_prefill.py -x -q -m gpu 2>&1 | tail -4
for sk in 0 1; do for pl in 2 1; do
    MILAB200_PREFILL_STREAMK=$sk MILAB200_PREFILL_ACT_PLANES=$pl timeout 300 python tools/perf_prefill.py --fmt fp8,fp4 --m 2048 \
        > $O/r2m2_perf_prefill_sk${sk}_pl${pl}.jsonl 2>$O/r2m2_perf_prefill_sk${sk}_pl${pl}.err
    echo "== stream-K $sk, activation planes $pl"; python - <<P
import json
for l in open('$O/r2m2_perf_prefill_sk${sk}_pl${pl}.jsonl'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['shape'], d['us'], d['TFLOPs'], d['frac_bf16_peak'], d['kernel'], d['schedule'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
P
done; done
for sk in 0 1; do
  for cfg in "--tokens 2048" "--workload gemma4-12b-mlp-fp4 --tokens 2048"; do
    tag=$(echo $cfg | tr -d ' -')_sk$sk
    MILAB200_PREFILL_STREAMK=$sk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2m2_bench_$tag.json 2>$O/r2m2_bench_$tag.err
    python -c "import json; d=json.load(open('$O/r2m2_bench_$tag.json')); print('$tag', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2m2_bench_$tag.err
  done
done
python tools/ncu_case.py fp4 3840 30720 16 > $O/r2m2_plain_case.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tc -s 3 -c 1 -f -o $O/r2m2_prof_fp4_gateup_m16 \
    python tools/ncu_case.py fp4 3840 30720 16 > $O/r2m2_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/r2m2_plain_case.log
