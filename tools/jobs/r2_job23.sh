#!/bin/bash
# round-2 validation of HEAD on one B200: GPU suite, smoke, default bench (both arms), ncu launch list + one full capture of the 32-Linear... 96-entry chain
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2j23_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2j23_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2j23_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/r2j23_smoke.txt
timeout 900 python bench.py > $O/r2j23_bench.json 2> $O/r2j23_bench.err; echo "bench rc=$?"; tail -c 1500 $O/r2j23_bench.json; tail -3 $O/r2j23_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2j23_bench_ref.json 2> $O/r2j23_bench_ref.err; echo "ref rc=$?"; tail -c 600 $O/r2j23_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2j23_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2j23_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_chain -s 2 -c 1 -f -o $O/r2j23_prof_chain32_fp8_m1 \
    python tools/ncu_chain_case.py 32 1 > $O/r2j23_ncu_chain.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/r2j23_ncu_chain.log
ls -la $O | grep r2j23
