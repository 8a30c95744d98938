#!/bin/bash
# One gpurun call for session 4: full GPU parity suite, contract bench lines (M = 1 and M = 16, FP8 and FP4, the
# reference arm), the per-shape decode sweep, the ncu launch list of the bench command and `ncu --set full` captures
# of the FP8 gate kernel at M = 1 and of the pre-split 16-token kernel.
tag=${1:-r1s4}
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_fp8.json 2> $out/${tag}_bench_fp8.err
python bench.py --tokens 16 --no-cpu-baseline > $out/${tag}_bench_fp8_m16.json 2> $out/${tag}_bench_fp8_m16.err
python bench.py --workload gemma4-12b-mlp-fp4 --no-cpu-baseline > $out/${tag}_bench_fp4.json 2> $out/${tag}_bench_fp4.err
python bench.py --workload gemma4-12b-mlp-fp4 --tokens 16 --no-cpu-baseline > $out/${tag}_bench_fp4_m16.json 2> $out/${tag}_bench_fp4_m16.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 200 python tools/perf_shapes.py > $out/${tag}_perf_shapes_decode.jsonl 2> $out/${tag}_perf_shapes.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
python tools/ncu_case.py fp8 4096 14336 1 6 > $out/${tag}_plain_fp8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_tc -s 3 -c 2 -f -o $out/${tag}_prof_fp8_gate_m1 \
    python tools/ncu_case.py fp8 4096 14336 1 6 > $out/${tag}_ncu_fp8.log 2>&1
python tools/ncu_case.py fp8 4096 14336 16 6 > $out/${tag}_plain_fp8_m16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"decode_tc|act_presplit" -s 6 -c 4 -f -o $out/${tag}_prof_fp8_gate_m16 \
    python tools/ncu_case.py fp8 4096 14336 16 6 > $out/${tag}_ncu_fp8_m16.log 2>&1
tail -3 $out/${tag}_pytest.log
cat $out/${tag}_bench_fp8.json $out/${tag}_bench_fp8_m16.json $out/${tag}_bench_fp4.json $out/${tag}_bench_fp4_m16.json $out/${tag}_bench_ref.json | cut -c1-700
python tools/ab_fmt.py < $out/${tag}_perf_shapes_decode.jsonl
