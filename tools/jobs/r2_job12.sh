#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_prefill.py -x -q -m gpu > $O/r2j12_pytest_prefill.log 2>&1; tail -4 $O/r2j12_pytest_prefill.log
timeout 300 python tools/make_golden.py $O/golden > $O/r2j12_golden.log 2>&1; tail -3 $O/r2j12_golden.log
( time timeout 900 python bench.py --no-cpu-baseline > $O/r2j12_bench.json 2>$O/r2j12_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j12_bench.json'))
print('main', round(d['value'],1), round(d['roofline']['frac'],4))
for e in d.get('extra',[]):
    print(' extra', e.get('name'), e.get('mode'), round(e.get('value',0),1), e.get('roofline',{}).get('achieved'), e.get('roofline',{}).get('frac'), e.get('roofline',{}).get('peak'), e.get('error'))
PY
tail -3 $O/r2j12_bench.err
