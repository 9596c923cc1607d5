"""Opcode histogram of an address range of a kernel's SASS (cuobjdump -sass listing filtered to instruction lines).
Usage: sass_hist.py k.sass 0x82b0 0x9540"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
cnt = collections.Counter(); n = 0
groups = {"MUFU": "mufu", "HMMA": "mma", "FFMA2": "fp32x2", "FMUL2": "fp32x2", "FADD2": "fp32x2", "FFMA": "fp32", "FMUL": "fp32",
          "FADD": "fp32", "FMNMX": "alu", "FSEL": "alu", "FSET": "alu", "FSETP": "alu", "LOP3": "alu", "MOV": "mov", "IMAD.MOV": "mov",
          "LDS": "lds", "STS": "sts", "LDGSTS": "ldgsts", "SHFL": "shfl", "LDL": "spill", "STL": "spill"}
gc = collections.Counter()
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if not m: continue
    a = int(m.group(1), 16)
    if lo <= a < hi:
        op = m.group(2); n += 1
        base = op.split(".")[0]
        if op.startswith("IMAD.MOV"): base = "IMAD.MOV"
        cnt[base] += 1
        gc[groups.get(base, "other")] += 1
print("instructions:", n)
print(" ".join(f"{k}:{v}" for k, v in gc.most_common()))
print(" ".join(f"{k}:{v}" for k, v in cnt.most_common()))
