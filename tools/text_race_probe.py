"""Stress cspe_format_fixed6 for intermittent mismatches: python tools/text_race_probe.py [iters]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import ops
from oracle import labels as O
from tests.test_gpu_parity import _nasty_doubles

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(1)
bad = 0
for it in range(iters):
    cols = int(rng.choice([1, 6, 7]))
    rows = 3001
    v = _nasty_doubles(rng, rows * cols)
    rng.shuffle(v)
    v = v.reshape(rows, cols)
    live = int(rng.integers(1, rows + 1))
    skew = int(rng.integers(0, 16))
    want = O.savetxt_fixed6(v[:live])
    junk = [torch.empty(int(rng.integers(1, 1 << 20)), dtype=torch.uint8, device="cuda") for _ in range(int(rng.integers(0, 4)))]
    n = torch.tensor([live], dtype=torch.int64, device="cuda")
    buf = torch.zeros((len(want) + 64,), dtype=torch.uint8, device="cuda")
    text, n_bytes, _ = ops.format_fixed6(torch.from_numpy(v).cuda(), n_rows=n, out=buf[skew:])
    del junk
    got = text[: len(want)].cpu().numpy().tobytes()
    if int(n_bytes.item()) != len(want) or got != want:
        bad += 1
        g, w = np.frombuffer(got, dtype=np.uint8), np.frombuffer(want, dtype=np.uint8)
        diff = np.nonzero(g != w)[0]
        # tile boundaries in the text: cumulative byte count every 1024 values
        lens = np.array([len(t) + 1 for t in want.decode().replace("\n", " ").split(" ")[:-1]])
        cum = np.concatenate([[0], np.cumsum(lens)])
        tiles = cum[::1024]
        print(f"iter {it}: cols {cols} live {live} skew {skew} len {len(want)} n_bytes {int(n_bytes.item())} diffs {len(diff)} at {diff[:12].tolist()}"
              f" got {bytes(g[diff[0]-8:diff[0]+12])!r} want {bytes(w[diff[0]-8:diff[0]+12])!r}; tile starts near: "
              f"{[int(t) for t in tiles if abs(int(t) - int(diff[0])) < 64]} (mod16 of text+first: {[(int(t)+skew) % 16 for t in tiles if abs(int(t) - int(diff[0])) < 64]})")
print(f"{bad} bad of {iters}")
