"""How many envs of the MomentumAgent / NoiseAgent configurations are bit-identical to the oracle?"""
import numpy as np
from bourse_b200 import abi, core, workloads
from oracle import oracle as orc
orc.build()
def run(groups, ogroups, n_envs, n_steps, seed, **kw):
    env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=65536, max_trades=65536, max_steps=n_steps, max_queue=256, **kw)
    env.set_agents(groups); env.run_agents(n_steps, seed)
    hist = env.history_all(n_steps); same = 0
    for e in range(n_envs):
        ce = orc.StepEnvNumpy(0, 0, 1, 1_000_000); ce.set_groups(ogroups); ce.run_agents(n_steps, seed, env_id=e, keyed=True)
        same += int(np.array_equal(hist[e], ce._history()) and env.get_orders(e) == ce.get_orders() and env.get_trades(e) == ce.get_trades())
    return same, n_envs, np.unique(env.env_errors())
print("momentum c4 x1000 steps", run(workloads.c4_groups(), workloads.c4_groups(), 96, 1000, 7))
print("momentum c4 dense_L", run(workloads.c4_groups(), workloads.c4_groups(), 96, 1000, 7, price_window=(0, 1024), live_cap=254))
g = [core.random_group(40, (40, 60), (10, 20), 2, 0.8), core.noise_group(100, 30, 2, 0.2, 0.2, 0.1, 15, 0.0, 1.0)]
og = [orc.random_group(40, (40, 60), (10, 20), 2, 0.8), orc.noise_group(100, 30, 2, 0.2, 0.2, 0.1, 15, 0.0, 1.0)]
print("noise x600 steps", run(g, og, 96, 600, 3))
