#!/bin/bash
# session m, job 11 (N GPUs, N = $1): the bench line at N GPUs on the final tree (tp_parity + the cfg #5 extra); tp_check when $2 = check
N=$1; O=gpurun_out; mkdir -p $O
if [ "${2:-}" = "check" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 tests/tp_check.py > $O/r2m11_tp_check_world$N.log 2>&1; echo "tp_check rc=$?"
  grep -E "TP_CHECK_OK|TP_CHAIN_OK|Error|error" $O/r2m11_tp_check_world$N.log | head -6
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2m11_bench_n$N.json 2> $O/r2m11_bench_n$N.err; echo "bench rc=$?"; tail -3 $O/r2m11_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2m11_bench_n$N.json').read().strip().splitlines()[-1])
tp=d.get('tp_parity') or {}
print('N=$N', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['method']['mode'][:24], 'tp_parity ok', tp.get('ok'), tp.get('max_rel_err_rowabs'), tp.get('identical_bits_across_ranks'), d['clocks'])
for e in d.get('extra',[]): print(e['name'], e.get('error') or (round(e['value'],1), e['mode'], (e.get('tp_parity') or {}).get('max_rel_err_rowabs'), (e.get('tp_parity') or {}).get('identical_bits_across_ranks')))
P
