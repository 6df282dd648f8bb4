"""The reference's OWN pytest files, vendored verbatim (zero edits; see README.md here), run against the CUDA library.

`import bourse` in those files resolves to this package's mirror: `bourse.core` is `bourse_b200.core` (the classes behind the
C ABI), `bourse.step_sim` / `bourse.data_processing` are the mirrors of the reference's pure-Python layer, `bourse.MAX_PRICE`
as in /root/reference/src/bourse/__init__.py:1-3.  Every test collected from this directory needs the GPU: the `gpu` marker
is attached by tests/conftest.py (the files themselves are not touched).
"""
import sys
import types

import bourse_b200
from bourse_b200 import core, data_processing, step_sim
from bourse_b200.step_sim import agents, runner

bourse = types.ModuleType("bourse")
bourse.__doc__ = "alias of bourse_b200 for the reference's unmodified tests"
bourse.__path__ = []  # a package: `from bourse.step_sim.agents import ...` must resolve through sys.modules
bourse.core, bourse.step_sim, bourse.data_processing, bourse.MAX_PRICE = core, step_sim, data_processing, bourse_b200.MAX_PRICE
sys.modules.setdefault("bourse", bourse)
sys.modules.setdefault("bourse.core", core)
sys.modules.setdefault("bourse.step_sim", step_sim)
sys.modules.setdefault("bourse.step_sim.agents", agents)
sys.modules.setdefault("bourse.step_sim.runner", runner)
sys.modules.setdefault("bourse.data_processing", data_processing)
