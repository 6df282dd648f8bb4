"""Per-CUDA-source-line instruction counts from `ncu --page source --print-source cuda,sass --csv`."""
import csv, subprocess, sys
rep = sys.argv[1]; n_events = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; addr_i = r.index("Address"); continue
    if hdr is None or len(r) < 10: continue
    if r[addr_i] == "-" and r[0].isdigit():
        try:
            agg.append((cur_file, int(r[0]), r[1].strip()[:90], float(r[hdr["Instructions Executed"]]), float(r[hdr["# Samples"]])))
        except ValueError:
            pass
tot = sum(a[3] for a in agg); ts = sum(a[4] for a in agg)
print(f"total inst {tot:.0f} per event {tot/n_events:.1f}")
for f, ln, src, inst, smp in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{inst/n_events:7.2f} {smp/ts*100:5.1f}%  {f}:{ln:<4d} {src}")
