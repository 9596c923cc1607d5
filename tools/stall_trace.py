"""Per-instruction stall samples of the hot loop in program order (ncu source page): stall_trace.py rep.ncu-rep [min_samples]"""
import csv, io, re, subprocess, sys
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
isrc, iex, ismp, ia = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
reasons = {h[6:]: hdr.index(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
ex = [int(r[iex]) for r in data]; main = max(set(ex), key=lambda v: v * ex.count(v))
base = int(data[0][ia], 16)
tot = sum(int(r[ismp]) for r in data)
for r in data:
    if int(r[iex]) < main * 0.9: continue
    n = int(r[ismp])
    if n < mins: continue
    rs = sorted(((int(r[i] or 0), k) for k, i in reasons.items()), reverse=True)[:3]
    print(f"{int(r[ia],16)-base:6x} {n:5d} {n/tot*100:5.2f}%  {r[isrc].strip()[:70]:70s} " + " ".join(f"{k}:{v}" for v, k in rs if v))
