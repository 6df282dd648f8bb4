import sys, time
import numpy as np
from bourse_b200 import abi, core, workloads
n_envs, n_steps = int(sys.argv[1]), int(sys.argv[2])
mo = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=mo, max_trades=2*mo, max_steps=n_steps, max_queue=128)
env.set_agents(workloads.c3_groups())
t0 = time.time()
env.run_agents(n_steps, 101)
dt = time.time() - t0
st = env.stats()
print(n_envs, n_steps, "ok", dt, st, "orders/s", st["instructions"]/dt, "errs", np.unique(env.env_errors()), "max orders", max(env.n_orders(e) for e in range(0, n_envs, max(1, n_envs//16))))
