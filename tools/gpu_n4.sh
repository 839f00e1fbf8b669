#!/bin/bash
# 8-GPU box: bench c2 / c4 and the 100 k-frame sweep (YOLO, COCO) at N = 1, 2, 4, 8
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n4_smi.log; nproc >> gpurun_out/n4_smi.log; nvidia-smi topo -m >> gpurun_out/n4_smi.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29700
for n in 1 2 4; do
  for cfg in c2 c4; do
    port=$((port+1))
    if [ $n -eq 1 ]; then cmd="python bench.py"; else cmd="$TR --nproc-per-node $n --master-port $port bench.py"; fi
    timeout 600 $cmd --gpus $n --steps 20 --warmup 5 --config $cfg --no-cpu-baseline --no-stress > gpurun_out/n4_bench_${cfg}_n$n.json 2> gpurun_out/n4_bench_${cfg}_n$n.err; echo "bench $cfg n=$n rc=$?"
  done
  for emit in yolo coco; do
    port=$((port+1))
    if [ $n -eq 1 ]; then cmd="python"; else cmd="$TR --nproc-per-node $n --master-port $port"; fi
    timeout 300 $cmd -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit $emit --repeat 2 > gpurun_out/n4_sweep_${emit}_n$n.json 2> gpurun_out/n4_sweep_${emit}_n$n.err; echo "sweep $emit n=$n rc=$?"
  done
done
python - <<PY
import json, glob
def last(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        return None
for cfg in ("c2", "c4"):
    base = None
    for n in (1, 2, 4):
        d = last("gpurun_out/n4_bench_%s_n%d.json" % (cfg, n))
        if d is None: print(cfg, n, "MISSING"); continue
        base = base or d["value"]
        print("bench", cfg, "N=%d" % n, round(d["value"]), "eff %.3f" % (d["value"] / (n * base)), "ms/step %.4f" % d["ms_per_step"],
              "win min/max %.3f/%.3f" % (d["timing"]["window_ms_min"], d["timing"]["window_ms_max"]), "roofline %.3f" % d["roofline"]["frac"],
              "e2e", round(d["e2e"]["value"]), "pcie %.1f" % d["e2e"]["pcie_ceiling_gbs"], "frac %.2f" % d["e2e"]["frac_of_pcie"], "devres", round(d["e2e"]["device_resident"]["value"]))
for emit in ("yolo", "coco"):
    base = None
    for n in (1, 2, 4):
        d = last("gpurun_out/n4_sweep_%s_n%d.json" % (emit, n))
        if d is None: print(emit, n, "MISSING"); continue
        base = base or d["frames_per_s_all_ranks"]
        print("sweep", emit, "N=%d" % n, round(d["frames_per_s_all_ranks"]), "eff %.3f" % (d["frames_per_s_all_ranks"] / (n * base)), "rank0/s", round(d["frames_per_s"]), d["host_timers"])
PY
tail -3 gpurun_out/n4_*.err | tail -60
