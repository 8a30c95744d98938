#!/bin/bash
# usage: tools/ab_perf.sh "<perf_shapes args>" "ENV=.. ENV=.." "ENV=.." ...   — one line of us/launch per configuration
args=$1; shift
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 150 python tools/perf_shapes.py $args 2>/dev/null | python tools/ab_fmt.py
done
