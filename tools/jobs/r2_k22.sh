#!/bin/bash
# session k, job 22: role timelines of the (reverted) FP4 chain: where does the epilogue stand relative to the MMAs?
O=gpurun_out; mkdir -p $O
timeout 300 python tools/chain_timeline.py llama3-70b-mlp-fp4 1 3 > $O/r2k22_timeline_70b.txt 2>&1; head -11 $O/r2k22_timeline_70b.txt | cut -c1-135
timeout 300 python tools/chain_timeline.py gemma4-12b-mlp-fp4 1 3 > $O/r2k22_timeline_gemma.txt 2>&1; head -11 $O/r2k22_timeline_gemma.txt | cut -c1-135
timeout 300 python tools/chain_timeline.py llama3.1-8b-mlp-fp8 1 3 > $O/r2k22_timeline_llama8b.txt 2>&1; head -11 $O/r2k22_timeline_llama8b.txt | cut -c1-135
