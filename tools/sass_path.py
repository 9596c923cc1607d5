"""Walk the hot path of the streaming kernel's main loop in a SASS listing: start at the loop head and follow the
fall-through / taken edges given on the command line.  Simpler: print per-region histograms between marker lines.
Usage: sass_path.py k.sass  -> prints the skeleton (branches, barriers, spills) with line numbers and addresses."""
import re, sys
for i, l in enumerate(open(sys.argv[1]).read().splitlines(), 1):
    if re.search(r"BRA|DEPBAR|WARPSYNC.ALL|SYNCS|LDL|STL|EXIT|BAR\.", l) and "COLLECTIVE" not in l:
        print(i, " ".join(l.split())[:90])
