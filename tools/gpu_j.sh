#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log; tail -30 gpurun_out/j_pytest.log
timeout 300 python tools/pc_bench.py 10 > gpurun_out/j_pc_bench.log 2>&1; cat gpurun_out/j_pc_bench.log
for emit in coco coco_host yolo; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/j_sweep_n1_$emit.json 2> gpurun_out/j_sweep_n1_$emit.err; echo "rc=$?"; tail -3 gpurun_out/j_sweep_n1_$emit.err; cut -c1-600 gpurun_out/j_sweep_n1_$emit.json
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
