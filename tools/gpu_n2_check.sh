#!/bin/bash
# 2-GPU box: the 2-rank NCCL sweep test and a short COCO sweep on two ranks (start barrier, copy-free COCO path)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "multirank" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 3 > gpurun_out/n2_sweep_coco.json 2> gpurun_out/n2_sweep_coco.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/n2_sweep_coco.json') if l.startswith('{')][-1]); print('coco N=2', [round(x) for x in d['frames_per_s_all_ranks_runs']], d['host_timers'])"
