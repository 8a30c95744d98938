#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 6 > $O/r2j16_timeline.txt 2>&1; tail -8 $O/r2j16_timeline.txt
