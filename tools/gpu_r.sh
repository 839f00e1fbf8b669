#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
for emit in yolo coco none; do
timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 4 > gpurun_out/r_sweep_$emit.json 2> gpurun_out/r_sweep_$emit.err
python -c "
import json; d=json.load(open('gpurun_out/r_sweep_$emit.json')); print('$emit', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'], d['io_threads'])"
done
