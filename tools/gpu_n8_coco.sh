#!/bin/bash
# 8-GPU box: the 100 k-frame sweep with COCO annotations formatted on the device at N = 1, 2, 4, 8; the same at 800 k
# frames (N = 1, 8: a rank's share of 100 k frames is only 24 graph groups = 25 ms at N = 8, so fill / drain of the
# pipeline shows); YOLO at N = 8; the 2-rank NCCL correctness test
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "multirank or shards" 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
run() {  # n frames emit tag repeat
  port=$((port+1))
  if [ $1 -eq 1 ]; then cmd="python"; else cmd="$TR --nproc-per-node $1 --master-port $port"; fi
  timeout 200 $cmd -m constructionsceneposeestimation_b200.sweep --frames $2 --emit $3 --repeat $5 > gpurun_out/n8c_sweep_$4_n$1.json 2> gpurun_out/n8c_sweep_$4_n$1.err; echo "sweep $4 n=$1 rc=$?"
}
for n in 1 2 4 8; do run $n 100000 coco coco 3; done
run 8 800000 coco coco800k 2
python - <<PY
import json
def last(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        return None
for tag, ns in (("coco", (1, 2, 4, 8)), ("coco800k", (8,))):
    base = None
    for n in ns:
        d = last("gpurun_out/n8c_sweep_%s_n%d.json" % (tag, n))
        if d is None: print(tag, n, "MISSING"); continue
        base = base or d["frames_per_s_all_ranks"] / n
        print("sweep %s N=%d" % (tag, n), round(d["frames_per_s_all_ranks"]), "eff %.3f" % (d["frames_per_s_all_ranks"] / (n * base)), "rank0/s", round(d["frames_per_s"]), [round(x) for x in d["frames_per_s_all_ranks_runs"]], {k: round(v, 4) for k, v in d["host_timers"].items()}, d["io_threads"])
PY
for f in gpurun_out/n8c_*.err; do tail -n 2 $f; done | grep -v "^$" | tail -20
