"""Summarise an .ncu-rep: headline raw metrics, opcode mix, stall reasons, top stalled SASS lines."""
import csv, re, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
n_events = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, v = raw[0], raw[-1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    for i, n in enumerate(h):
        if n == k:
            print(f"{k:75s} {v[i]}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hh = src[1]; ix = {n: i for i, n in enumerate(hh)}; data = src[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
byop = defaultdict(float); st = defaultdict(float)
for r in data:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']].strip())
    op = s.split()[0].split('.')[0] if s else '?'
    byop[op] += f(r, 'Instructions Executed'); st[op] += f(r, '# Samples')
tot = sum(byop.values()); ts = sum(st.values())
print(f"\ntotal warp instructions {tot:.0f}  per event {tot/n_events:.1f}")
for op, x in sorted(byop.items(), key=lambda x: -x[1])[:22]:
    print(f"  {op:10s} {x/tot*100:6.2f}% inst {st[op]/ts*100:6.2f}% samples  per-event {x/n_events:7.2f}")
cols = [c for c in hh if c.startswith('stall_') and 'Not Issued' not in c]
T = sum(sum(f(r, c) for r in data) for c in cols)
print("\nstall reasons")
for c, x in sorted({c: sum(f(r, c) for r in data) for c in cols}.items(), key=lambda x: -x[1])[:8]:
    print(f"  {c:26s}{x/T*100:6.2f}%")
print("\ntop stalled instructions (samples, executed)")
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:25]:
    print("  ", r[ix['Source']][:70].ljust(72), int(f(r, '# Samples')), int(f(r, 'Instructions Executed')))
