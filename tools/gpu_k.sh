#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log; tail -6 gpurun_out/k_pytest.log
for emit in coco; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/k_sweep_n1_$emit.json 2> gpurun_out/k_sweep_n1_$emit.err; echo "rc=$?"; tail -3 gpurun_out/k_sweep_n1_$emit.err; cut -c1-700 gpurun_out/k_sweep_n1_$emit.json
done
timeout 300 python tools/next_rows_bench.py > gpurun_out/k_next_rows.log 2>&1; cat gpurun_out/k_next_rows.log
timeout 300 python tools/pc_bench.py 10 > gpurun_out/k_pc_bench.log 2>&1; cat gpurun_out/k_pc_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "rc=$?"; tail -3 gpurun_out/k_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/k_bench_ref.json 2> gpurun_out/k_bench_ref.err; echo "rc=$?"
