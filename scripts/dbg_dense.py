import sys
import numpy as np
from bourse_b200 import abi, core, workloads
n_envs, n_steps = int(sys.argv[1]), int(sys.argv[2])
kw = dict(price_window=(20, 180), live_cap=128) if (len(sys.argv) < 4 or sys.argv[3] == "dense") else {}
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=8192, max_trades=16384, max_steps=n_steps, max_queue=128, **kw)
env.set_agents(workloads.c3_groups())
env.run_agents(n_steps, 101)
print(env.stats(), np.unique(env.env_errors()))
