"""Discrete event market simulation (mirror of /root/reference/src/bourse/step_sim/__init__.py)."""
from . import agents, runner  # noqa: F401
from .runner import run  # noqa: F401
