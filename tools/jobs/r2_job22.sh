#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 16 3 > $O/r2j22_timeline_m16.txt 2>&1; head -12 $O/r2j22_timeline_m16.txt | cut -c1-150; tail -3 $O/r2j22_timeline_m16.txt
