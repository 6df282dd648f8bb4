"""SASS listing in address order with executed counts / samples from an .ncu-rep (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; n_events = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hh = src[1]; ix = {n: i for i, n in enumerate(hh)}
for k, r in enumerate(src[2:]):
    try:
        e = float(r[ix['Instructions Executed']]); s = float(r[ix['# Samples']])
    except Exception:
        continue
    print(f"{k:5d} {e/n_events:6.3f} {int(s):6d}  {r[ix['Source']]}")
