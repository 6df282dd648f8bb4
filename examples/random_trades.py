"""The reference's examples/random_trades.py (config C1) on the B200 back end: the only change is the import.

    python examples/random_trades.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bourse_b200 as bourse  # noqa: E402  (`import bourse` in the reference)
from bourse_b200.step_sim.agents import RandomAgent  # noqa: E402

TICK_SIZE = 2


def run(seed: int, n_steps: int, n_agents: int):
    agents = [RandomAgent(i, 0.5, (10, 100), (20, 50), TICK_SIZE) for i in range(n_agents)]
    env = bourse.core.StepEnv(seed, 0, TICK_SIZE, 100_000)
    market_data = bourse.step_sim.run(env, agents, n_steps, seed)
    return market_data


if __name__ == "__main__":
    data = run(101, 200, 100)
    print({k: v[-1] for k, v in data.items() if k in ("bid_price", "ask_price", "bid_vol", "ask_vol", "trade_vol")})
