"""Agent base classes (contract of /root/reference/src/bourse/step_sim/agents/base_agent.py)."""
import typing

import numpy as np

# (action u32, side bool, vol u32, trader u32, price u32, order_id u64) — rust/src/types.rs:33-40
InstructionArrays = typing.Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]


class BaseAgent:
    """Agents that talk to a ``StepEnv`` call by call."""

    def update(self, rng: np.random.Generator, env) -> None:
        raise NotImplementedError


class BaseNumpyAgent:
    """Agents that return instruction arrays for ``StepEnvNumpy.submit_instructions``."""

    def update(self, rng: np.random.Generator, level_2_data: np.ndarray) -> InstructionArrays:
        raise NotImplementedError
