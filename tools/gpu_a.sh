#!/bin/bash
# round-2 GPU call A: parity, scan variants on stress masks, bench (graph vs eager)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python tools/scan_microbench.py 30 zeros,c2,c2_dense,c2_textured,blocks16,checker2,noise2,noise100,c4 1:1,0:0,1:0,0:1 > gpurun_out/a_micro.log 2>&1
cat gpurun_out/a_micro.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_graph.json 2> gpurun_out/a_bench_graph.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --eager --no-cpu-baseline --no-stress > gpurun_out/a_bench_eager.json 2> gpurun_out/a_bench_eager.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --config c4 --no-cpu-baseline > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err; echo "rc=$?"
tail -c 1500 gpurun_out/a_bench_graph.err gpurun_out/a_bench_eager.err gpurun_out/a_bench_c4.err
cat gpurun_out/a_bench_graph.json gpurun_out/a_bench_eager.json gpurun_out/a_bench_c4.json
