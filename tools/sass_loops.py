"""List the loops (backward branches) of a kernel's SASS listing with their instruction counts and opcode groups.
Usage: sass_loops.py k.sass [min_instrs]"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ins = []
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = [a for a, _ in ins]
def hist(lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            t2 = re.sub(r"^@!?U?P\d\s+", "", t)
            op = t2.split()[0].split(".")[0]
            if t2.startswith("IMAD.MOV"): op = "MOV"
            c[op] += 1
    return c
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a:
            n = sum(1 for x in addr if tgt <= x <= a)
            if n >= mn:
                h = hist(tgt, a)
                print(f"loop {tgt:#x}..{a:#x}: {n} instrs  HMMA {h['HMMA']} MUFU {h['MUFU']} LDL {h['LDL']} STL {h['STL']}")
                print("   ", " ".join(f"{k}:{v}" for k, v in h.most_common(40)))
