#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -30 gpurun_out/e_pytest.log
timeout 300 python tools/next_rows_bench.py > gpurun_out/e_next_rows.log 2>&1; cat gpurun_out/e_next_rows.log
for emit in coco yolo; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/e_sweep_n1_$emit.json 2> gpurun_out/e_sweep_n1_$emit.err; echo "rc=$?"
  tail -3 gpurun_out/e_sweep_n1_$emit.err; cat gpurun_out/e_sweep_n1_$emit.json
done
mkdir -p /dev/shm/sw && timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit yolo --out /dev/shm/sw > gpurun_out/e_sweep_n1_yolo_files.json 2>&1; cat gpurun_out/e_sweep_n1_yolo_files.json; ls /dev/shm/sw/labels | wc -l; rm -rf /dev/shm/sw
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "rc=$?"; tail -5 gpurun_out/e_bench.err; cat gpurun_out/e_bench.json
