"""Developer run of the multi-asset example population (k_sim<.., MKT>): python scripts/dbg_market.py n_books n_steps [dense|paged] [max_queue]"""
import sys
import numpy as np
from bourse_b200 import abi, core, workloads
n_envs, n_steps = int(sys.argv[1]), int(sys.argv[2])
kw = dict(price_window=(20, 180), live_cap=128) if (len(sys.argv) < 4 or sys.argv[3] == "dense") else {}
mq = int(sys.argv[4]) if len(sys.argv) > 4 else 80
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=8192, max_trades=16384, max_steps=n_steps, max_queue=mq, assets=2, **kw)
g, a = workloads.market_example_groups()
env.set_agents(g, assets=a)
env.run_agents(n_steps, 101)
print(env.stats(), np.unique(env.env_errors()))
