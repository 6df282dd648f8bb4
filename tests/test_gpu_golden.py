"""GPU suite: the reference's known-answer tests through the CUDA-backed `bourse.core` mirror."""
import numpy as np
import pytest

from . import scenarios

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scenario", scenarios.ALL_BOOK + scenarios.ALL_ENV + scenarios.ALL_NUMPY,
                         ids=lambda f: f.__name__)
def test_reference_known_answers_cuda(core, scenario):
    scenario(core)


def test_bad_order_id_raises(core):
    """orderbook.rs:642: the reference panics on an unknown id; the mirror raises."""
    ob = core.OrderBook(0, 1)
    ob.place_order(True, 1, 0, price=5)
    with pytest.raises(core.PanicException):
        ob.cancel_order(7)
    env = core.StepEnv(1, 0, 1, 100)
    env.cancel_order(3)
    with pytest.raises(core.PanicException):
        env.step()


def test_book_stays_usable_after_a_refused_instruction(core, oracle):
    """The reference mutates nothing before it panics on an unknown id (orderbook.rs:642), so a caller that catches the
    panic keeps a working book.  Here: the error is reported by the call that raised it and by no later call, the env's
    sticky word (bb_env_errors) remembers it until bb_clear_errors, and results stay identical to an oracle book that
    never saw the bad instruction."""
    ob, ref = core.OrderBook(0, 1), oracle.OrderBook(0, 1)
    for b in (ob, ref):
        b.set_time(1); b.place_order(True, 10, 0, price=50)
    with pytest.raises(core.PanicException):
        ob.cancel_order(99)
    for b in (ob, ref):   # every later call works and reports nothing
        b.set_time(2); b.place_order(False, 4, 1, price=50)
        b.set_time(3); b.cancel_order(0)
        b.set_time(4); b.place_order(False, 6, 2, price=60)
    assert ob.get_orders() == ref.get_orders() and ob.get_trades() == ref.get_trades()
    assert ob.bid_ask() == ref.bid_ask()
    assert int(ob._env.env_errors()[0]) & 0x10          # sticky record of the refused id ...
    ob._env.clear_errors()
    assert int(ob._env.env_errors()[0]) == 0            # ... until cleared
    # Env mode: the step that carried the bad cancel raises, the next steps do not
    env, renv = core.StepEnv(1, 0, 1, 100), oracle.StepEnv(1, 0, 1, 100)
    env.cancel_order(3)
    with pytest.raises(core.PanicException):
        env.step()
    renv.step()
    for e in (env, renv):
        e.place_order(True, 5, 0, price=10); e.step()
        e.place_order(False, 5, 1, price=10); e.step()
    assert env.get_trades() == renv.get_trades() and env.get_orders() == renv.get_orders()


def test_python_runner_and_agents(core):
    """tests/test_step_sim/test_env.py:118-145, test_numpy_api.py:85-120, test_agents.py:6-31"""
    import bourse_b200
    from bourse_b200.step_sim.agents import BaseAgent, BaseNumpyAgent, NumpyRandomAgents, RandomAgent

    class A(BaseAgent):
        def __init__(self, side, start):
            self.side, self.start, self.k = side, start, 0

        def update(self, _rng, env):
            p = self.start + self.k if self.side else self.start - self.k
            env.place_order(self.side, 10, 101, price=p)
            self.k += 1

    env = core.StepEnv(101, 0, 1, 100_000)
    data = bourse_b200.step_sim.run(env, [A(True, 10), A(False, 50)], 10, 101)
    assert np.array_equal(data["bid_price"], 10 + np.arange(10))
    assert np.array_equal(data["ask_price"], 50 - np.arange(10))
    assert np.array_equal(data["bid_vol"], 10 * np.arange(1, 11))
    assert np.array_equal(data["ask_vol"], 10 * np.arange(1, 11))
    assert np.array_equal(data["bid_vol_0"], 10 * np.ones(10))
    assert np.array_equal(data["trade_vol"], np.zeros(10))

    class B(BaseNumpyAgent):
        def __init__(self, side, start):
            self.side, self.start, self.k = side, start, 0

        def update(self, _rng, _l2):
            p = self.start + self.k if self.side else self.start - self.k
            self.k += 1
            return (np.array([1], np.uint32), np.array([self.side]), np.array([10], np.uint32), np.array([101], np.uint32),
                    np.array([p], np.uint32), np.array([0], np.uint64))

    env = core.StepEnvNumpy(101, 0, 1, 100_000)
    data = bourse_b200.step_sim.run(env, [B(True, 10), B(False, 50)], 10, 101, use_numpy=True)
    assert np.array_equal(data["bid_price"], 10 + np.arange(10))
    assert np.array_equal(data["ask_vol_0"], 10 * np.ones(10))

    env = core.StepEnv(101, 0, 1, 100_000)
    agent = RandomAgent(0, 1.0, (10, 20), (20, 30), 2)
    rng = np.random.default_rng(101)
    agent.update(rng, env)
    assert agent.order_id == 0
    env.step()
    assert env.order_status(0) == 1
    agent.update(rng, env)
    assert agent.order_id is None
    env.step()
    assert env.order_status(0) == 3
    agent.update(rng, env)
    assert agent.order_id == 1

    env = core.StepEnvNumpy(101, 0, 1, 100_000)
    ins = NumpyRandomAgents(20, (10, 60), (10, 20), 2).update(np.random.default_rng(101), env.level_2_data())
    ids = env.submit_instructions(ins)
    assert list(ids) == list(range(20))


def test_random_trades_example_c1(core, oracle):
    """BASELINE config C1 (examples/random_trades.py:4-20) at reduced length: the Python agents drive
    the CUDA env and the oracle env with the same numpy rng; everything but the shuffle is shared,
    and the shuffle follows the same Xoroshiro stream, so the market data must be identical."""
    import bourse_b200
    from bourse_b200.step_sim.agents import RandomAgent

    def sim(mod, runner_env_cls):
        agents = [RandomAgent(i, 0.5, (10, 100), (20, 50), 2) for i in range(100)]
        env = runner_env_cls(101, 0, 2, 100_000)
        rng = np.random.default_rng(101)
        for _ in range(40):
            for a in agents:
                a.update(rng, env)
            env.step()
        return env.get_market_data(), env.get_trades(), env.get_orders()

    d_gpu, t_gpu, o_gpu = sim(core, core.StepEnv)
    d_cpu, t_cpu, o_cpu = sim(oracle, oracle.StepEnv)
    assert set(d_gpu) == set(d_cpu) and len(d_gpu) == 45
    for k in d_cpu:
        assert d_gpu[k].dtype == np.uint32 and np.array_equal(d_gpu[k], d_cpu[k]), k
    assert t_gpu == t_cpu and len(t_gpu) > 50
    assert o_gpu == o_cpu
