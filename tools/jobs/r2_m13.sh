#!/bin/bash
# session m, job 13: smoke + the default bench line on the final tree
O=gpurun_out; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2m13_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/r2m13_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2m13_bench.json 2> $O/r2m13_bench.err; echo "bench rc=$?"; tail -2 $O/r2m13_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2m13_bench.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])
for e in d.get('extra',[]):
    r=e.get('roofline') or {}
    print(e['name'], e.get('error') or (round(e['value'],1), round(r.get('frac',0),4), e.get('mode'), (e.get('clocks') or {}).get('reasons')))
P
