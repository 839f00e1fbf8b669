#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "sweep or coco" 2>&1 | tail -2
for t in 2 3; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 4 --io-threads $t > gpurun_out/m4_sweep_coco_t$t.json 2> gpurun_out/m4_sweep_coco_t$t.err
  python -c "
import json; d=json.load(open('gpurun_out/m4_sweep_coco_t$t.json')); print('threads $t', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'])"
done
mkdir -p /dev/shm/sw && timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit yolo --repeat 2 --out /dev/shm/sw > gpurun_out/m4_sweep_yolo_files.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/m4_sweep_yolo_files.json')); print('yolo files', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'])"; ls /dev/shm/sw/labels | wc -l; rm -rf /dev/shm/sw
timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco_host --repeat 2 > gpurun_out/m4_sweep_coco_host.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/m4_sweep_coco_host.json')); print('coco_host', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'])"
