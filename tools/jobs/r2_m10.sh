#!/bin/bash
# session m, job 10: validation of the working tree on one B200: full GPU suite, smoke, default bench (both arms), ncu launch list,

O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2m10_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2m10_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2m10_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2m10_smoke.txt
( time timeout 900 python bench.py > $O/r2m10_bench.json 2> $O/r2m10_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 $O/r2m10_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2m10_bench.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])
for e in d.get('extra',[]):
    r=e.get('roofline') or {}
    print(e['name'], e.get('error') or (round(e['value'],1), round(r.get('frac',0),4), e.get('mode'), r.get('kernel'), (e.get('clocks') or {}).get('sm_mhz'), (e.get('clocks') or {}).get('reasons')))
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2m10_bench_ref.json 2> $O/r2m10_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2m10_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2m10_ncu_launches.log 2>&1; echo "ncu list rc=$?"
