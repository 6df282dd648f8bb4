"""CPU suite: the vendored copies of the reference's test files are still byte-identical to the reference's own
(checked wherever /root/reference is mounted, i.e. in the build container; the GPU box runs the copies)."""
import filecmp
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tests"
PAIRS = {"test_order_book.py": "test_order_book.py", "test_env.py": "test_step_sim/test_env.py",
         "test_numpy_api.py": "test_step_sim/test_numpy_api.py", "test_agents.py": "test_step_sim/test_agents.py"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
@pytest.mark.parametrize("ours,theirs", sorted(PAIRS.items()))
def test_vendored_reference_tests_are_unmodified(ours, theirs):
    assert filecmp.cmp(os.path.join(HERE, "golden", "ref_tests", ours), os.path.join(REF, theirs), shallow=False)
