"""Aggregate an ncu source page by address regions.  Usage: ncu_regions.py rep name:lo:hi ...   (offsets relative to base)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
regs = [(a.split(":")[0], int(a.split(":")[1], 16), int(a.split(":")[2], 16)) for a in sys.argv[2:]]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
base = int(data[0][hdr.index("Address")], 16)
ia, iex, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
cols = {n[6:]: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
tot_s = sum(int(r[ismp]) for r in data); tot_e = sum(int(r[iex]) for r in data)
print("total samples", tot_s, "total warp-instructions", tot_e)
for name, lo, hi in regs:
    s = e = n = 0; st = {k: 0 for k in cols}
    for r in data:
        a = int(r[ia], 16) - base
        if lo <= a < hi:
            s += int(r[ismp]); e += int(r[iex]); n += 1
            for k, v in cols.items(): st[k] += int(r[v])
    top = ", ".join(f"{k} {v/max(s,1)*100:.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:6])
    print(f"{name:14s} {n:5d} instrs  exec {e/tot_e*100:5.1f}%  samples {s/tot_s*100:5.1f}%  samples/exec {s/max(e,1)*1e3:7.2f}  | {top}")
