#!/bin/bash
# session k, job 16: fast RMS reduction option: parity + the fused MLP-block stacks; then the full default bench line
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_rmsnorm.py -x -q -m gpu > $O/r2k16_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/r2k16_pytest.txt
for cfg in "--workload gemma4-12b-mlp-fp4 --fuse-gate-up --norm-fast" "--fuse-gate-up --norm-fast" "--fuse-gate-up --tokens 16 --norm-fast" "--workload gemma4-12b-mlp-fp4 --fuse-gate-up --tokens 16 --norm-fast" \
           "--fuse-gate-up --tokens 8 --norm-fast" "--workload llama3-70b-mlp-fp4 --fuse-gate-up --norm-fast"; do
    tag=$(echo $cfg | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $cfg > $O/r2k16_bench_$tag.json 2>$O/r2k16_bench_$tag.err
    python -c "import json,sys; d=json.load(open('$O/r2k16_bench_$tag.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 $O/r2k16_bench_$tag.err
done
( time timeout 900 python bench.py > $O/r2k16_bench_default.json 2> $O/r2k16_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2k16_bench_default.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline']['frac'], d['e2e']['value'])
for e in d.get('extra',[]):
    r=e.get('roofline') or {}
    print(e['name'], e.get('error') or (round(e['value'],1), round(r.get('frac',0),4), r.get('kernel')))
P
tail -3 $O/r2k16_bench_default.err
