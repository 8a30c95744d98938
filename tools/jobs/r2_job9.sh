#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 900 python bench.py > $O/r2j9_bench_default.json 2>$O/r2j9_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j9_bench_default.json'))
print('main', round(d['value'],1), d['method']['mode'][:5], round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('sample','')[:80])
print('ctx', {k:(v.get('median_ms') if isinstance(v,dict) else v) for k,v in d.get('cpu_context',{}).items()})
for e in d.get('extra',[]):
    print(' extra', e.get('name'), e.get('mode'), round(e.get('value',0),1), e.get('roofline',{}).get('frac'), e.get('error'), (e.get('clocks') or {}).get('reasons'))
PY
tail -5 $O/r2j9_bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
