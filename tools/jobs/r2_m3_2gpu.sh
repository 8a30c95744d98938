#!/bin/bash
# session m, job 3 (2 GPUs): tensor-parallel parity (tests/tp_check.py under torchrun) and the bench line at N = 2 with tp_parity + the cfg #5 extra
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -m gpu > $O/r2m3_pytest_tp.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2m3_pytest_tp.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/tp_check.py > $O/r2m3_tp_check_world2.log 2>&1; echo "tp_check rc=$?"; grep -E "TP_CHECK_OK|TP_CHAIN_OK|TP_PARITY|Error|error" $O/r2m3_tp_check_world2.log | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2m3_bench_n2.json 2> $O/r2m3_bench_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2m3_bench_n2.json').read().strip().splitlines()[-1])
print('N=2', round(d['value'],1), d['method']['mode'][:20], d.get('tp_parity'))
for e in d.get('extra',[]): print(e['name'], e.get('error') or (round(e['value'],1), e['mode'], e.get('tp_parity')))
P
