#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log; tail -6 gpurun_out/l_pytest.log
for emit in coco yolo; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/l_sweep_n1_$emit.json 2> gpurun_out/l_sweep_n1_$emit.err; echo "rc=$?"; tail -3 gpurun_out/l_sweep_n1_$emit.err; cut -c1-500 gpurun_out/l_sweep_n1_$emit.json
done
mkdir -p /dev/shm/sw && timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --out /dev/shm/sw > gpurun_out/l_sweep_n1_coco_file.json 2>&1; cut -c1-400 gpurun_out/l_sweep_n1_coco_file.json; ls -la /dev/shm/sw; python -c "
import json; d=json.load(open('/dev/shm/sw/coco_rank00.json')); print(len(d['images']), len(d['annotations']), d['annotations'][0], d['annotations'][-1]['id'])"; rm -rf /dev/shm/sw
timeout 300 python tools/next_rows_bench.py 2>&1 | grep "f1 " > gpurun_out/l_next_rows.log; cat gpurun_out/l_next_rows.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stress > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "rc=$?"; tail -3 gpurun_out/l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/l_bench.json')); print(d['value'], d['e2e'])"
