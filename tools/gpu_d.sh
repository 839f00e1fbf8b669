#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -25 gpurun_out/d_pytest.log
for emit in yolo coco none; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/d_sweep_n1_$emit.json 2> gpurun_out/d_sweep_n1_$emit.err; echo "rc=$?"
  tail -3 gpurun_out/d_sweep_n1_$emit.err; cat gpurun_out/d_sweep_n1_$emit.json
done
timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit yolo --eager > gpurun_out/d_sweep_n1_yolo_eager.json 2>&1; cat gpurun_out/d_sweep_n1_yolo_eager.json
mkdir -p /dev/shm/sw && timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit yolo --out /dev/shm/sw > gpurun_out/d_sweep_n1_yolo_files.json 2>&1; cat gpurun_out/d_sweep_n1_yolo_files.json; ls /dev/shm/sw/labels | wc -l; rm -rf /dev/shm/sw
