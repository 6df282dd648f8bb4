"""Opcode histogram with modifiers (e.g. IMAD.MOV.U32) from an .ncu-rep source page."""
import csv, re, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; n_events = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hh = src[1]; ix = {n: i for i, n in enumerate(hh)}; data = src[2:]
by = defaultdict(float)
for r in data:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']].strip())
    op = s.split()[0] if s else '?'
    try: by[op] += float(r[ix['Instructions Executed']])
    except Exception: pass
tot = sum(by.values())
for op, x in sorted(by.items(), key=lambda x: -x[1])[:45]:
    print(f"{op:28s} {x/n_events:7.2f} {x/tot*100:5.1f}%")
