#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "sweep or coco or yolo" 2>&1 | tail -3
timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 4 > gpurun_out/r_sweep_coco.json 2> gpurun_out/r_sweep_coco.err
python -c "
import json; d=json.load(open('gpurun_out/r_sweep_coco.json')); print('coco', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'], d['io_threads'])"
