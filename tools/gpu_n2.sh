#!/bin/bash
# 2-GPU box: multi-rank parity, sweep N=1 vs N=2, bench N=1 / N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_smi.log; nproc >> gpurun_out/n2_smi.log
python -m pytest tests -m gpu -q -x -k "multirank or sweep or step_graph" > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2_pytest.log; tail -5 gpurun_out/n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for emit in yolo coco; do
  timeout 300 python -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/n2_sweep_n1_$emit.json 2> gpurun_out/n2_sweep_n1_$emit.err; echo "rc=$?"
  timeout 300 $TR --nproc-per-node 2 --master-port 29601 -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/n2_sweep_n2_$emit.json 2> gpurun_out/n2_sweep_n2_$emit.err; echo "rc=$?"
  tail -3 gpurun_out/n2_sweep_n2_$emit.err
  python - <<PY
import json
for n in (1, 2):
    d = json.loads([l for l in open("gpurun_out/n2_sweep_n%d_$emit.json" % n) if l.startswith("{")][-1])
    print("$emit", "N=%d" % n, round(d["frames_per_s_all_ranks"]), "frames/s; rank0", round(d["frames_per_s"]), d["host_timers"], d.get("per_rank"))
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stress > gpurun_out/n2_bench_n1.json 2> gpurun_out/n2_bench_n1.err; echo "rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_n2.json 2> gpurun_out/n2_bench_n2.err; echo "rc=$?"; tail -5 gpurun_out/n2_bench_n2.err
python - <<PY
import json
for n in (1, 2):
    d = json.loads([l for l in open("gpurun_out/n2_bench_n%d.json" % n) if l.startswith("{")][-1])
    print("bench N=%d" % n, round(d["value"]), "ms/step", d["ms_per_step"], d["timing"]["window_ms_min"], d["timing"]["window_ms_max"], "e2e", round(d["e2e"]["value"]), d["e2e"]["pcie_ceiling_gbs"], round(d["e2e"]["device_resident"]["value"]))
PY
