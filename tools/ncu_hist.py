"""Per-stage opcode histogram (weighted by execution count) and stall breakdown of the streaming kernel from an .ncu-rep."""
import csv, collections, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; d = dict(zip(hdr, rows[2]))
print(f"ncu: {float(d['gpu__time_duration.sum']):.4f} ms  inst {float(d['smsp__inst_executed.sum'])/1e6:.1f}M  issue {float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']):.1f}%  "
      f"fma {float(d['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']):.1f}% xu {float(d['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']):.1f}% "
      f"alu {float(d['sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']):.1f}% lsu {float(d['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']):.1f}% regs {d['launch__registers_per_thread']}")
st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr if "issue_stalled" in h and "per_issue_active" in h}
tot = sum(st.values())
print("stalls:", ", ".join(f"{k} {v/tot*100:.1f}" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v / tot > 0.015))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex = collections.Counter(int(r[iex]) for r in data if int(r[iex]) > 0)
main = max(ex.items(), key=lambda kv: kv[0] * kv[1])[0]
h = collections.Counter()
for r in data:
    e = int(r[iex])
    if e == 0: continue
    m = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.MOV)?)", r[isrc])
    op = m.group(1) if m else r[isrc][:10]
    if op == "IMAD.MOV": op = "MOV"
    h[op] += e / main
print(f"per-stage instrs {sum(h.values()):.0f} (main count {main}):", " ".join(f"{k}:{v:.0f}" for k, v in h.most_common(34)))
