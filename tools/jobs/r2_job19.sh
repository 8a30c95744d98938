#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 > $O/r2j19_timeline_gemma.txt 2>&1; head -12 $O/r2j19_timeline_gemma.txt | cut -c1-170; tail -6 $O/r2j19_timeline_gemma.txt
