#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_int4.py tests/test_rmsnorm.py -x -q -m gpu > $O/r2j11_pytest_new.log 2>&1; tail -15 $O/r2j11_pytest_new.log
