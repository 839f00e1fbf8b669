#!/bin/bash
# round-2 profile captures: launch list of the bench command + full captures of the scan kernel (c2, c4)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline --no-stress > gpurun_out/i_bench_plain.json 2> gpurun_out/i_bench_plain.err; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline --no-stress > gpurun_out/i_ncu_l.log 2>&1; echo "rc=$?"
for cfg in c2 c4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_scan -s 4 -c 1 -f -o gpurun_out/r02_scan_$cfg python tools/scan_microbench.py 3 $cfg > gpurun_out/i_ncu_$cfg.log 2>&1; tail -2 gpurun_out/i_ncu_$cfg.log
done
ls -la gpurun_out/r02_*
