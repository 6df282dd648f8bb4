"""Python agents with the reference's interfaces (/root/reference/src/bourse/step_sim/agents/)."""
from .base_agent import BaseAgent, BaseNumpyAgent, InstructionArrays  # noqa: F401
from .random_agent import NumpyRandomAgents, RandomAgent  # noqa: F401
