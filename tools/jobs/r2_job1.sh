#!/bin/bash
# round 2, GPU job 1: parity suite on the balanced-tile decode kernels + A/B against whole 128-row tiles
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/r2j1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2j1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j1_pytest.log
tail -5 $O/r2j1_pytest.log
for cfg in "MILAB200_TILE_ROWS=128" "MILAB200_TILE_ROWS=0" "MILAB200_COST_UNIT=24" "MILAB200_COST_FIXUP=200"; do
  echo "== $cfg" | tee -a $O/r2j1_ab.txt
  env $cfg timeout 300 python tools/perf_shapes.py --fmt fp8 --only llama8b --m 1,4,8,16 2>>$O/r2j1_ab.err | tee -a $O/r2j1_ab_$cfg.jsonl | python tools/ab_fmt.py | tee -a $O/r2j1_ab.txt
  env $cfg timeout 300 python tools/perf_shapes.py --fmt fp4 --m 1,4,16 2>>$O/r2j1_ab.err | tee -a $O/r2j1_ab_$cfg.jsonl | python tools/ab_fmt.py | tee -a $O/r2j1_ab.txt
done
for cfg in "MILAB200_TILE_ROWS=128" "MILAB200_TILE_ROWS=0"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2j1_bench_$cfg.json 2>$O/r2j1_bench_$cfg.err
  python -c "import json,sys; d=json.load(open('$O/r2j1_bench_$cfg.json')); print('$cfg', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
