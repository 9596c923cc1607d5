"""Static opcode histogram of the streaming kernel's hot region (the two unrolled 8-cell groups of the un-masked stage path) from
cuobjdump -sass of an object file: static_count.py obj.o [kernel-substring].  A CPU-side proxy for the per-stage dynamic
instruction count that ncu reports (tools/ncu_hist.py)."""
import collections, re, subprocess, sys
obj = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "vcb_stream2_kernelILi3ELb1ELb1E"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); fn[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
    if m and cur: fn[cur].append((int(m.group(1), 16), m.group(2)))
name = [k for k in fn if want in k][0]
ins = fn[name]
def op(t):
    m = re.match(r"(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.MOV)?)", t)
    o = m.group(1) if m else t[:8]
    return "MOV" if o == "IMAD.MOV" else o
hm = [i for i, (a, t) in enumerate(ins) if op(t) == "HMMA"]
# clusters of HMMAs not interrupted by a backward branch
clusters, cur = [], [hm[0]]
for a, b in zip(hm, hm[1:]):
    back = any(op(ins[j][1]) == "BRA" and (lambda m: m and int(m.group(1), 16) < ins[j][0])(re.search(r"0x([0-9a-f]+)", ins[j][1])) for j in range(a, b))
    if back: clusters.append(cur); cur = []
    cur.append(b)
clusters.append(cur)
best = max(clusters, key=len)
lo, hi = best[0], best[-1]
while lo > 0 and op(ins[lo][1]) not in ("DEPBAR", "BRA"): lo -= 1
while hi < len(ins) - 1 and op(ins[hi][1]) != "LDGDEPBAR": hi += 1
# second LDGDEPBAR may follow the last HMMA closely; extend to the last LDGDEPBAR before the next BRA / LDL block
h = collections.Counter(op(t) for a, t in ins[lo:hi + 1])
print(f"{name[:60]} total {len(ins)} | hot region {hi - lo + 1} instrs, {len(best)} HMMA:", " ".join(f"{k}:{v}" for k, v in h.most_common(30)))
