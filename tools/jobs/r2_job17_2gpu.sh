#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/tp_check.py > $O/r2j17_tp_check_w2.log 2>&1
echo "tp_check rc=$?"; grep -E "TP_CHECK_OK|TP_CHAIN_OK|TP_NCCL_ENTRY_OK|AssertionError|Error" $O/r2j17_tp_check_w2.log | cut -c1-200 | head
for mode in auto launches; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 --mode $mode > $O/r2j17_bench_tp2_$mode.json 2>$O/r2j17_bench_tp2_$mode.err
python -c "import json; d=json.load(open('$O/r2j17_bench_tp2_$mode.json')); print('tp2 $mode', round(d['value'],1), d['method']['mode'][:6], d['tp_parity']['ok'], d['tp_parity']['max_rel_err_rowabs'])" || tail -5 $O/r2j17_bench_tp2_$mode.err
done
timeout 300 python -m pytest tests/test_gpu_tp.py -x -q -m gpu 2>&1 | tail -2
