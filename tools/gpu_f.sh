#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/text_race_probe.py 400 > gpurun_out/f_text_probe.log 2>&1; tail -30 gpurun_out/f_text_probe.log
python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -30 gpurun_out/f_pytest.log
