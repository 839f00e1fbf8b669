#!/bin/bash
# ncu captures of the scan kernel on slow-path masks
set -x
mkdir -p gpurun_out
for spec in "blocks16 0:0 b_blocks16_v00" "blocks16 0:1 b_blocks16_v01" "c2_textured 0:1 b_textured_v01" "c2 0:0 b_c2_v00" "c2 0:1 b_c2_v01"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_scan -s 4 -c 1 -f -o gpurun_out/$3 python tools/scan_microbench.py 3 $1 $2 > gpurun_out/$3.log 2>&1
  tail -2 gpurun_out/$3.log
done
ls -la gpurun_out/*.ncu-rep
