#!/bin/bash
# session k, job 3: L2 look-ahead in the decode_tc producer (FP4 unpacked path, 3..16 tokens): A/B per shape
set -u
O=gpurun_out; mkdir -p $O
for la in 0 2 4 8; do
  echo "== FP4 la=$la"
  MILAB200_TC_L2_LOOKAHEAD_FP4=$la timeout 300 python tools/perf_shapes.py --fmt fp4 --m 4,8,16 --only gate_up > $O/r2k3_fp4_gu_la$la.jsonl 2>$O/r2k3_err.txt
  MILAB200_TC_L2_LOOKAHEAD_FP4=$la timeout 300 python tools/perf_shapes.py --fmt fp4 --m 4,16 --only llama70b_up >> $O/r2k3_fp4_gu_la$la.jsonl 2>>$O/r2k3_err.txt
  MILAB200_TC_L2_LOOKAHEAD_FP4=$la timeout 300 python tools/perf_shapes.py --fmt fp4 --m 4,16 --only gemma_down >> $O/r2k3_fp4_gu_la$la.jsonl 2>>$O/r2k3_err.txt
  python -c "
import json
for l in open('$O/r2k3_fp4_gu_la$la.jsonl'):
    d=json.loads(l); print(d['shape'],d['M'],d['us'],d['GBps'],d['kernel'])"
done
for la in 0 2 4; do
  echo "== FP8 la=$la"
  MILAB200_TC_L2_LOOKAHEAD_FP8=$la timeout 300 python tools/perf_shapes.py --fmt fp8 --m 1,8,16 --only llama8b > $O/r2k3_fp8_la$la.jsonl 2>>$O/r2k3_err.txt
  python -c "
import json
for l in open('$O/r2k3_fp8_la$la.jsonl'):
    d=json.loads(l); print(d['shape'],d['M'],d['us'],d['GBps'],d['kernel'])"
done
tail -5 $O/r2k3_err.txt
