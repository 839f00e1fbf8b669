#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "sweep or coco or pointcloud" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log; tail -4 gpurun_out/m_pytest.log
timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 3 > gpurun_out/m_sweep_n1_coco.json 2> gpurun_out/m_sweep_n1_coco.err; echo "rc=$?"; tail -3 gpurun_out/m_sweep_n1_coco.err; cut -c1-500 gpurun_out/m_sweep_n1_coco.json; python -c "
import json; d=json.load(open('gpurun_out/m_sweep_n1_coco.json')); print(d['frames_per_s_all_ranks_runs'])"
timeout 300 python tools/next_rows_bench.py 2>&1 | grep "f1 " > gpurun_out/m_next_rows.log; cat gpurun_out/m_next_rows.log
