"""Python-side simulation loop, same contract as the reference's ``bourse.step_sim.run``
(/root/reference/src/bourse/step_sim/runner.py:12-120): agents act, then ``env.step()``."""
from __future__ import annotations

import typing

import numpy as np

from .. import core
from .agents import BaseAgent, BaseNumpyAgent


def run(env, agents: typing.Iterable, n_steps: int, seed: int, show_progress: bool = False,
        use_numpy: bool = False) -> typing.Dict[str, np.ndarray]:
    agents = list(agents)
    if use_numpy:
        assert isinstance(env, core.StepEnvNumpy)
        assert all(isinstance(a, BaseNumpyAgent) for a in agents), "Agents should implement BaseNumpyAgent"
    else:
        assert isinstance(env, core.StepEnv)
        assert all(isinstance(a, BaseAgent) for a in agents), "Agents should implement BaseAgent"
    rng = np.random.default_rng(seed)
    it = range(n_steps)
    if show_progress:
        try:
            import tqdm
            it = tqdm.trange(n_steps)
        except ImportError:
            pass
    for _ in it:
        if use_numpy:
            level_2_data = env.level_2_data()
            for agent in agents:
                env.submit_instructions(agent.update(rng, level_2_data))
        else:
            for agent in agents:
                agent.update(rng, env)
        env.step()
    return env.get_market_data()
