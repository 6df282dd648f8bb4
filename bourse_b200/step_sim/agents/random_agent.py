"""Uniform random agents; same decision rules and draw order as the reference's Python agents
(/root/reference/src/bourse/step_sim/agents/random_agent.py:59-91 and :129-166)."""
import numpy as np

from .base_agent import BaseAgent, BaseNumpyAgent, InstructionArrays


class RandomAgent(BaseAgent):
    def __init__(self, i, activity_rate, tick_range, vol_range, tick_size):
        self.i, self.activity_rate = i, activity_rate
        self.tick_range, self.vol_range, self.tick_size = tick_range, vol_range, tick_size
        self.order_id = None

    def update(self, rng: np.random.Generator, env) -> None:
        if rng.random() >= self.activity_rate:
            return
        if self.order_id is not None and env.order_status(self.order_id) == 1:
            env.cancel_order(self.order_id)
            self.order_id = None
            return
        tick = rng.integers(*self.tick_range)
        vol = rng.integers(*self.vol_range)
        side = bool(rng.choice([True, False]))
        self.order_id = env.place_order(side, int(vol), self.i, price=int(tick) * self.tick_size)


class NumpyRandomAgents(BaseNumpyAgent):
    def __init__(self, n_agents, tick_range, vol_range, tick_size):
        self.n_agents, self.tick_range, self.vol_range, self.tick_size = n_agents, tick_range, vol_range, tick_size

    def update(self, rng: np.random.Generator, level_2_data) -> InstructionArrays:
        n = self.n_agents
        sides = rng.choice([True, False], size=n).astype(bool)
        # NB the reference samples volumes from tick_range (random_agent.py:149); kept for parity
        vols = rng.integers(*self.tick_range, size=n, dtype=np.uint32)
        prices = rng.integers(*self.tick_range, size=n, dtype=np.uint32) * self.tick_size
        return (np.ones(n, dtype=np.uint32), sides, vols, np.arange(n, dtype=np.uint32), prices.astype(np.uint32),
                np.zeros(n, dtype=np.uint64))
