"""bourse_b200 — B200-native batched limit-order-book simulator.

Mirrors the reference package layout (`/root/reference/src/bourse/__init__.py:1-3`):
``bourse_b200.core`` stands where ``bourse.core`` (the PyO3 extension) does, ``step_sim`` holds the
Python runner and agents.  Importing the package does not need a GPU; creating any book does.
"""
from . import abi, core, data_processing, step_sim  # noqa: F401

MAX_PRICE = 2**32 - 1
__all__ = ["abi", "core", "data_processing", "step_sim", "MAX_PRICE"]
