"""Developer timing of each public-API call of the end-to-end C3 pass (host wall clock, synchronous calls)."""
import time, sys
import numpy as np, torch
from bourse_b200 import abi, core, workloads
n_envs, n_steps = 4096, 1000
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=65536, max_trades=65536, max_steps=n_steps, max_queue=128)
groups = workloads.c3_groups()
hist = torch.empty((n_envs, n_steps, 9), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
hist_pageable = np.empty((n_envs, n_steps, 9), dtype=np.uint32)
def T(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); print(f"{name:28s} {(time.perf_counter()-t0)*1e3:8.2f} ms"); return r
for it in range(3):
    print("iter", it)
    T("reset", env.reset); T("set_agents", lambda: env.set_agents(groups)); T("run_agents", lambda: env.run_agents(n_steps, 101))
    T("history_all pinned", lambda: env.history_all(n_steps, hist)); T("history_all pageable", lambda: env.history_all(n_steps, hist_pageable))
    T("history_all alloc", lambda: env.history_all(n_steps))
    T("stats", env.stats); T("level_1_data", env.level_1_data)
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=uuid,serial,clocks.mem,clocks.sm,temperature.gpu,power.draw", "--format=csv"], capture_output=True, text=True).stdout)
s = torch.cuda.current_stream(); env.set_stream(s.cuda_stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for mode in ("noflush", "flush"):
    for it in range(3):
        a, m, b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(s); env.reset(); m.record(s); env.run_agents(n_steps, 101, sync=False); b.record(s)
        if mode == "flush": flush.fill_(1)
        torch.cuda.synchronize(); print(mode, "events: reset %.3f ms  k_sim %.3f ms" % (a.elapsed_time(m), m.elapsed_time(b)))
