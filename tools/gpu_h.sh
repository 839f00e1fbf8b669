#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/pc_bench.py 10 > gpurun_out/h_pc_bench.log 2>&1; cat gpurun_out/h_pc_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pc_ -s 4 -c 2 -f -o gpurun_out/h_pc python tools/pc_bench.py 2 > gpurun_out/h_pc_ncu.log 2>&1; tail -3 gpurun_out/h_pc_ncu.log
python -m pytest tests -m gpu -q -x -k "pointcloud" > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
