#!/bin/bash
# 8-GPU box: the 100 k-frame sweep with COCO annotations formatted on the device, at N = 1, 2, 4, 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
for n in 1 2 4 8; do
  port=$((port+1))
  if [ $n -eq 1 ]; then cmd="python"; else cmd="$TR --nproc-per-node $n --master-port $port"; fi
  timeout 200 $cmd -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit coco --repeat 3 > gpurun_out/n8c_sweep_coco_n$n.json 2> gpurun_out/n8c_sweep_coco_n$n.err; echo "sweep coco n=$n rc=$?"
done
python - <<PY
import json
def last(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        return None
base = None
for n in (1, 2, 4, 8):
    d = last("gpurun_out/n8c_sweep_coco_n%d.json" % n)
    if d is None: print("coco", n, "MISSING"); continue
    base = base or d["frames_per_s_all_ranks"]
    print("sweep coco N=%d" % n, round(d["frames_per_s_all_ranks"]), "eff %.3f" % (d["frames_per_s_all_ranks"] / (n * base)), "rank0/s", round(d["frames_per_s"]), [round(x) for x in d["frames_per_s_all_ranks_runs"]], d["host_timers"], d["io_threads"])
PY
tail -3 gpurun_out/n8c_*.err | tail -30
