#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/text_race_probe.py 600 > gpurun_out/g_text_probe.log 2>&1; tail -5 gpurun_out/g_text_probe.log
for i in 1 2 3; do python -m pytest tests -m gpu -q -x > gpurun_out/g_pytest_$i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest_$i.log; tail -4 gpurun_out/g_pytest_$i.log; done
grep -B30 "short test summary" gpurun_out/g_pytest_1.log | tail -40
timeout 300 python tools/next_rows_bench.py 2>&1 | grep "f1 " > gpurun_out/g_next_rows.log; cat gpurun_out/g_next_rows.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stress > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "rc=$?"; tail -5 gpurun_out/g_bench.err; cut -c1-1200 gpurun_out/g_bench.json
