"""Top stalled / most executed SASS lines of an .ncu-rep source page: python tools/ncu_hot.py rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel only
start = 1
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[start + 1:]:
    if len(r) < 10 or r[0] == "Kernel Name": break
    body.append(r)
S, E, T = ix["Warp Stall Sampling (All Samples)"], ix["Instructions Executed"], ix["Avg. Threads Executed"]
tot_s = sum(int(r[S]) for r in body); tot_e = sum(int(r[E]) for r in body)
print(f"instructions {len(body)}, samples {tot_s}, warp-instr executed {tot_e}")
print("--- by stall samples")
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][S]))[:n]:
    print(f"{i:5d} {100*int(r[S])/tot_s:5.1f}%  exec {int(r[E]):9d} thr {r[T]:>5s}  {r[ix['Source']].strip()[:80]}")
print("--- by executed count")
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][E]))[:n]:
    print(f"{i:5d} exec {int(r[E]):9d} ({100*int(r[E])/tot_e:4.1f}%) thr {r[T]:>5s}  {r[ix['Source']].strip()[:80]}")
