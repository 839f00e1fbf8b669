#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
timeout 600 python tools/scan_microbench.py 30 zeros,c2,c2_dense,c2_textured,blocks16,checker2,noise2,noise100,c4 1:1,0:0,1:0,0:1 > gpurun_out/c_micro.log 2>&1
cat gpurun_out/c_micro.log
