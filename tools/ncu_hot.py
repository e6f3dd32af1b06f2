"""Program-order stall view of one kernel from an ncu report's source page (run here, no GPU).
usage: python tools/ncu_hot.py report.ncu-rep [min_samples=100]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = rows[2:]
ia, isrc, iall, iex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
names = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
cols = {k: h.index(k) for k in names}
tot = sum(int(r[iall] or 0) for r in data)
agg = {k: sum(int(r[c] or 0) for r in data) for k, c in cols.items()}
print("total samples", tot, {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
run = runs = 0
for r in data:
    s = int(r[iall] or 0); src = r[isrc].strip()
    if src.startswith(("FFMA", "CS2R")):
        run += s; runs += 1; continue
    if runs:
        print(f"      ... {runs} FFMA/FFMA2/CS2R, samples {run}"); run = runs = 0
    if s >= thr:
        st = {k[6:]: int(r[c] or 0) for k, c in cols.items() if int(r[c] or 0) >= max(20, s // 10)}
        print(r[ia][-4:], str(s).rjust(6), r[iex].rjust(9), src[:64].ljust(64), st)
