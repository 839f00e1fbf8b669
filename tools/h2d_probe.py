"""H2D bandwidth from pinned memory: one copy vs chunks vs several streams (B200 box PCIe)."""
import torch, time, json
dev = torch.device("cuda")
n = 64 * 1080 * 1920
src = torch.empty(n, dtype=torch.int32, pin_memory=True); src.fill_(3)
dst = torch.empty(n, dtype=torch.int32, device=dev)
def run(chunks, nstreams, reps=5):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    step = n // chunks
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for c in range(chunks):
            with torch.cuda.stream(streams[c % nstreams]):
                dst[c * step:(c + 1) * step].copy_(src[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return round(n * 4 / best / 1e9, 2)
for chunks, ns in ((1, 1), (64, 1), (64, 2), (64, 4), (8, 2), (8, 4), (256, 4)):
    print(json.dumps({"chunks": chunks, "streams": ns, "GB/s": run(chunks, ns)}))
# D2H for reference
back = torch.empty(n // 8, dtype=torch.int32, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); back.copy_(dst[: n // 8], non_blocking=True); torch.cuda.synchronize()
print(json.dumps({"d2h_GB/s": round(n // 8 * 4 / (time.perf_counter() - t0) / 1e9, 2)}))
