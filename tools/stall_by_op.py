"""Warp-stall samples of the streaming kernel grouped by opcode and reason (ncu source page): stall_by_op.py rep.ncu-rep"""
import csv, io, re, subprocess, sys, collections
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); ismp = hdr.index("# Samples")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {r: hdr.index(r) for r in reasons}
tot = collections.Counter(); by = collections.defaultdict(collections.Counter)
for r in data:
    m = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z0-9_]+)", r[isrc]); op = m.group(1) if m else "?"
    for k, i in idx.items():
        v = int(r[i] or 0); by[op][k[6:]] += v; tot[k[6:]] += v
T = sum(tot.values())
print("total samples", T, " ".join(f"{k}:{v / T * 100:.1f}%" for k, v in tot.most_common(10)))
for op, c in sorted(by.items(), key=lambda kv: -sum(kv[1].values()))[:22]:
    s = sum(c.values())
    print(f"{op:10s} {s / T * 100:5.1f}%  " + " ".join(f"{k}:{v / T * 100:.1f}" for k, v in c.most_common(5) if v / T > 0.002))
