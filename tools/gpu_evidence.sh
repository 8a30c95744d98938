#!/bin/bash
# One gpurun call: GPU parity tests, the contract bench lines, the ncu launch list of the bench
# command and one `ncu --set full` capture of the dominant decode kernel (FP8 and FP4).
# usage: tools/gpu_evidence.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_fp8.json 2> $out/${tag}_bench_fp8.err
python bench.py --workload gemma4-12b-mlp-fp4 --no-cpu-baseline > $out/${tag}_bench_fp4.json 2> $out/${tag}_bench_fp4.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
python tools/ncu_case.py fp8 4096 14336 1 6 > $out/${tag}_plain_fp8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_tc -s 3 -c 2 -f -o $out/${tag}_prof_fp8_gate_m1 \
    python tools/ncu_case.py fp8 4096 14336 1 6 > $out/${tag}_ncu_fp8.log 2>&1
python tools/ncu_case.py fp4 3840 30720 1 6 > $out/${tag}_plain_fp4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_mx4 -s 3 -c 2 -f -o $out/${tag}_prof_fp4_gateup_m1 \
    python tools/ncu_case.py fp4 3840 30720 1 6 > $out/${tag}_ncu_fp4.log 2>&1
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench_fp8.json $out/${tag}_bench_fp4.json $out/${tag}_bench_ref.json
