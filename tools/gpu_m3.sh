#!/bin/bash
mkdir -p gpurun_out
for t in 1 2 3 4; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 4 --io-threads $t > gpurun_out/m3_sweep_coco_t$t.json 2> gpurun_out/m3_sweep_coco_t$t.err
  python -c "
import json; d=json.load(open('gpurun_out/m3_sweep_coco_t$t.json')); print('threads $t', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'])"
done
