#!/bin/bash
# session k, job 28: validation of the working tree on one B200: full GPU suite, smoke, default bench (both arms), ncu launch list,
# ncu --set full of the Llama-70B FP4 and Gemma FP4 chains (whole stack = one launch)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2k28_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2k28_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k28_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2k28_smoke.txt
( time timeout 900 python bench.py > $O/r2k28_bench.json 2> $O/r2k28_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 $O/r2k28_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2k28_bench.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])
for e in d.get('extra',[]):
    r=e.get('roofline') or {}
    print(e['name'], e.get('error') or (round(e['value'],1), round(r.get('frac',0),4), e.get('mode'), r.get('kernel'), (e.get('clocks') or {}).get('sm_mhz'), (e.get('clocks') or {}).get('reasons')))
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2k28_bench_ref.json 2> $O/r2k28_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2k28_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2k28_ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_chain_case.py 16 1 llama3-70b-mlp-fp4 > $O/r2k28_plain_chain70b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_chain -s 2 -c 1 -f -o $O/r2k28_prof_chain16_fp4_70b_m1 \
    python tools/ncu_chain_case.py 16 1 llama3-70b-mlp-fp4 > $O/r2k28_ncu_chain70b.log 2>&1; echo "ncu 70b rc=$?"; tail -1 $O/r2k28_plain_chain70b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_chain -s 2 -c 1 -f -o $O/r2k28_prof_chain48_fp4_gemma_m1 \
    python tools/ncu_chain_case.py 48 1 gemma4-12b-mlp-fp4 > $O/r2k28_ncu_chaingemma.log 2>&1; echo "ncu gemma rc=$?"
ls -la $O/*.ncu-rep | grep r2k28
