"""The examples run as written (the reference's examples/random_trades.py and the two Rust examples, on this back end)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name, *args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", name), *args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    return out.stdout


def test_random_trades_example():
    assert "bid_price" in _run("random_trades.py")


def test_random_agents_example():
    out = _run("random_agents.py", "256")
    assert "trades" in out and "orders/s" in out


def test_multi_asset_example():
    out = _run("multi_asset.py", "8")
    assert "trades of asset 0" in out and "trades of asset 1" in out and "'error_envs': 0" in out


def test_vector_env_torch_example():
    out = _run("vector_env_torch.py")
    assert "mean traded volume" in out and "'error_envs': 0" in out
