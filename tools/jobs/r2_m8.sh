#!/bin/bash
# session m, job 8: chain on the per-rank shard shapes of the 4- / 8-GPU stacks (one GPU); ncu --set full of the batched GLU kernel
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "shard_shapes" 2>&1 | tail -5
python tools/ncu_prefill_case.py fp8 4096 28672 2048 3 glu > $O/r2m8_plain_glu.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_tc -s 1 -c 1 -f -o $O/r2m8_prof_prefill_fp8_glu_m2048 \
    python tools/ncu_prefill_case.py fp8 4096 28672 2048 3 glu > $O/r2m8_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/r2m8_plain_glu.log
