#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 0 1 3; do echo "variant $v"; CSPE_PC_VARIANT=$v timeout 300 python tools/pc_bench.py 10 2>&1 | grep both; done > gpurun_out/h2_pc_variants.log 2>&1; cat gpurun_out/h2_pc_variants.log
for v in 1 3; do CSPE_PC_VARIANT=$v python -m pytest tests -m gpu -q -x -k "pointcloud" 2>&1 | tail -1; done
