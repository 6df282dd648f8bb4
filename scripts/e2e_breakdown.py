"""Developer timing of each public-API call of the end-to-end C3 pass (host wall clock, synchronous calls)."""
import time, sys
import numpy as np, torch
from bourse_b200 import abi, core, workloads
n_envs, n_steps = 4096, 1000
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=65536, max_trades=65536, max_steps=n_steps, max_queue=128,
                      price_window=(20, 180), live_cap=128)
groups = workloads.c3_groups()
hist = torch.empty((n_envs, n_steps, 9), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
def T(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); print(f"{name:28s} {(time.perf_counter()-t0)*1e3:8.3f} ms"); return r
for it in range(3):
    print("iter", it)
    T("reset", env.reset); T("set_agents", lambda: env.set_agents(groups)); T("run_agents", lambda: env.run_agents(n_steps, 101))
    T("history_all pinned", lambda: env.history_all(n_steps, hist)); T("stats", env.stats)
    for ch in (0, 64, 128, 256, 504, 1000):
        env.reset(); env.set_agents(groups)
        T(f"run_agents_to_host chunk={ch}", lambda: env.run_agents_to_host(n_steps, 101, hist, chunk_steps=ch))
