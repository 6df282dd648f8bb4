"""Developer run of a C5-style shard (deep books, paged engine) for profiling: preload + a short phase 2."""
import sys
import numpy as np, torch
from bourse_b200 import abi, core, workloads
n_envs, n_rest, n_steps, per_step = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 10_000
s = workloads.c5_stream(n_rest, n_steps, per_step, seed=100)
dev = torch.device("cuda", 0)
d1 = torch.from_numpy(s[:n_rest].view(np.uint8)).to(dev).repeat(n_envs, 1).contiguous()
d2 = torch.from_numpy(s[n_rest:].view(np.uint8)).to(dev).repeat(n_envs, 1).contiguous()
o1 = torch.arange(0, n_envs + 1, dtype=torch.int64, device=dev) * n_rest
o2 = torch.arange(0, n_envs + 1, dtype=torch.int64, device=dev) * (n_steps * per_step)
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=n_rest + n_steps * per_step, max_trades=1 << 20,
                      max_steps=n_steps, max_queue=32, pages_smem=int(sys.argv[4]) if len(sys.argv) > 4 else 192, pages_total=192)
torch.cuda.synchronize()
env.replay_device(d1.data_ptr(), o1.data_ptr()); env.synchronize()
import time; t0 = time.time()
env.replay_device(d2.data_ptr(), o2.data_ptr()); env.synchronize()
print("phase 2", time.time() - t0, env.stats(), np.unique(env.env_errors()))
