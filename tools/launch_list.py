"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares of ONE steady-state step:
the launches between the last two launches of the streaming kernel.  Usage: launch_list.py launches.csv [kernel-substring]"""
import csv, sys, collections
path = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else "vcb_stream"
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]; data = rows[1:]
iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
def us(r):
    v = float(r[ival].replace(",", "")); u = r[iunit]
    return v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
hits = [i for i, r in enumerate(data) if key in r[iname]]
assert len(hits) >= 2, "need two launches of the key kernel"
step = data[hits[-2]:hits[-1]]
agg = collections.OrderedDict()
for r in step:
    a = agg.setdefault(r[iname], [0, 0.0]); a[0] += 1; a[1] += us(r)
tot = sum(a[1] for a in agg.values())
print(f"# launches in the step: {len(step)}; sum of durations {tot/1e3:.3f} ms")
print("kernel,launches,total_us,share_pct")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{n},{t:.1f},{t/tot*100:.2f}")
