#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 3 > $O/r2j7_timeline.txt 2>&1; head -12 $O/r2j7_timeline.txt
