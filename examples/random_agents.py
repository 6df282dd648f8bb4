"""crates/step_sim/examples/random_agents/main.rs, batched: the same agent set (50 + 50 RandomAgents) driving 4096
independent Envs for 1000 steps inside one persistent kernel (config C3).

    python examples/random_agents.py [n_envs]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bourse_b200 import abi, core  # noqa: E402

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_steps=1000, max_queue=128,
                      price_window=(20, 180))                              # Env::new(0, 1, 1_000_000, true), n_envs times
env.set_agents([core.random_group(50, (40, 60), (10, 20), 2, 0.8),          # RandomAgents::new(50, (40, 60), (10, 20), 2, 0.8)
                core.random_group(50, (10, 90), (50, 70), 2, 0.2)])         # RandomAgents::new(50, (10, 90), (50, 70), 2, 0.2)
t0 = time.perf_counter()
hist = env.run_agents_to_host(1000, seed=101)                              # sim_runner(&mut env, &mut agents, 101, 1000, ..)
dt = time.perf_counter() - t0
st = env.stats()
print(f"{st['trades']} trades, {st['instructions']} instructions in {dt * 1e3:.1f} ms "
      f"({st['instructions'] / dt:.3e} orders/s); env 0 bid-ask now {tuple(int(x) for x in hist[0, -1, 1:3])}")
