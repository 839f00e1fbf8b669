#!/bin/bash
# A/B of cspe_pack_rows inside the 100 k-frame COCO sweep: CSPE_PACK_VARIANT = 0 (byte stores) / 1 (five words + funnel shift) / 2 (two aligned chunks)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "sweep or coco or yolo or pack" 2>&1 | tail -3
one() { timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 3 > gpurun_out/r_sweep_$1.json 2> gpurun_out/r_sweep_$1.err
python -c "
import json; d=json.load(open('gpurun_out/r_sweep_$1.json')); print('$1', [round(x) for x in d['frames_per_s_all_ranks_runs']], {k: round(v,4) for k,v in d['host_timers'].items()})"; }
CSPE_PACK_VARIANT=0 one pack0
CSPE_PACK_VARIANT=1 one pack1
CSPE_PACK_VARIANT=2 one pack2
