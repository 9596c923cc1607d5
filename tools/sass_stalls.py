"""Sum of the stall-count fields (bits 105..108 of each 128-bit SASS instruction) over the executed instructions of the
streaming kernel: the single-warp issue time of one stage if no scoreboard ever waited.
Usage: sass_stalls.py lib.so mangled_kernel_name rep.ncu-rep"""
import csv, io, re, subprocess, sys, collections
lib, fn, rep = sys.argv[1:4]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout.splitlines()
enc = {}
i = 0
cur = None
for l in out:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", l)
    if m:
        cur = int(m.group(1), 16); lo = int(m.group(3), 16); txt = m.group(2)
        enc[cur] = [txt, lo, None]
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", l)
    if m and cur is not None and enc[cur][2] is None:
        enc[cur][2] = int(m.group(1), 16)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ia, iex = hdr.index("Address"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
ex = collections.Counter(int(r[iex]) for r in data if int(r[iex]) > 0)
main = max(ex.items(), key=lambda kv: kv[0] * kv[1])[0]
tot_stall = 0.0; n = 0.0; hist = collections.Counter(); yld = 0
byop = collections.defaultdict(lambda: [0.0, 0.0])
for r in data:
    e = int(r[iex])
    if e == 0: continue
    a = int(r[ia], 16) - base
    if a not in enc or enc[a][2] is None: continue
    hi = enc[a][2]
    stall = (hi >> 41) & 0xf
    w = e / main
    tot_stall += w * stall; n += w; hist[stall] += w
    op = re.sub(r"^@!?U?P\d\s+", "", enc[a][0]).split()[0].split(".")[0]
    byop[op][0] += w; byop[op][1] += w * stall
print(f"instrs/stage {n:.0f}  sum of stall fields {tot_stall:.0f} cycles  (avg {tot_stall / n:.2f} per instr)")
print("stall histogram:", " ".join(f"{k}:{v:.0f}" for k, v in sorted(hist.items())))
print("by opcode (count, avg stall):", " ".join(f"{k}:{v[0]:.0f}/{v[1] / v[0]:.1f}" for k, v in sorted(byop.items(), key=lambda kv: -kv[1][1])[:24]))
