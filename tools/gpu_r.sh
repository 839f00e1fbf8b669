#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "sweep or coco or yolo" 2>&1 | tail -3
one() { timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $1 --repeat 4 > gpurun_out/r_sweep_$2.json 2> gpurun_out/r_sweep_$2.err
python -c "
import json; d=json.load(open('gpurun_out/r_sweep_$2.json')); print('$2', [round(x) for x in d['frames_per_s_all_ranks_runs']], {k: round(v,4) for k,v in d['host_timers'].items()})"; }
one none none
one coco coco
