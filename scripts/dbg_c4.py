"""Developer run of the C4 population (RandomAgents + MomentumAgent, level-2 records) on the general engine:
python scripts/dbg_c4.py n_envs n_steps"""
import sys
import numpy as np
from bourse_b200 import abi, core, workloads
n_envs, n_steps = int(sys.argv[1]), int(sys.argv[2])
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=8192, max_trades=16384, max_steps=n_steps, max_queue=128)
env.set_agents(workloads.c4_groups())
env.run_agents(n_steps, 7)
print(env.stats(), np.unique(env.env_errors()))
