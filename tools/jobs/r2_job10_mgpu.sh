#!/bin/bash
# round 2, multi-GPU job: tensor-parallel parity at world 8 / 4 / 2 (tests/tp_check.py under torchrun) and bench lines
set -u
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
run_check() {  # world
  w=$1
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29600 + w)) tests/tp_check.py > $O/r2j10_tp_check_w$w.log 2>&1
  echo "tp_check world=$w rc=$?"; grep -E "TP_CHECK_OK|TP_CHAIN_OK|TP_PARITY|AssertionError|Error" $O/r2j10_tp_check_w$w.log | head -12
}
for w in 8 4 2; do [ $w -le $N ] && run_check $w; done
for w in 2 4 8; do
  [ $w -le $N ] || continue
  for mode in auto launches; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29700 + w)) bench.py --gpus $w --steps 20 --warmup 5 --mode $mode > $O/r2j10_bench_tp${w}_$mode.json 2>$O/r2j10_bench_tp${w}_$mode.err
    python -c "import json; d=json.load(open('$O/r2j10_bench_tp${w}_$mode.json')); print('tp$w $mode', round(d['value'],1), d['method']['mode'][:6], d.get('tp_parity'))" || tail -5 $O/r2j10_bench_tp${w}_$mode.err
  done
done
w=$N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 29790 bench.py --gpus $w --steps 20 --warmup 5 --workload llama3-70b-mlp-fp4 > $O/r2j10_bench_tp${w}_70b_fp4.json 2>$O/r2j10_bench_tp${w}_70b_fp4.err
python -c "import json; d=json.load(open('$O/r2j10_bench_tp${w}_70b_fp4.json')); print('70b fp4 tp$w', round(d['value'],1), d['method']['mode'][:6], d.get('tp_parity'))" || tail -5 $O/r2j10_bench_tp${w}_70b_fp4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 29791 bench.py --gpus $w --steps 20 --warmup 5 --workload llama3-70b-mlp-fp4 --tokens 4 > $O/r2j10_bench_tp${w}_70b_fp4_m4.json 2>$O/r2j10_bench_tp${w}_70b_fp4_m4.err
python -c "import json; d=json.load(open('$O/r2j10_bench_tp${w}_70b_fp4_m4.json')); print('70b fp4 M4 tp$w', round(d['value'],1), d['method']['mode'][:6], d.get('tp_parity'))" || tail -5 $O/r2j10_bench_tp${w}_70b_fp4_m4.err
