#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 6 > $O/r2j3_timeline_m1.txt 2>&1; cat $O/r2j3_timeline_m1.txt
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 6 --fuse > $O/r2j3_timeline_m1_fuse.txt 2>&1; cat $O/r2j3_timeline_m1_fuse.txt
