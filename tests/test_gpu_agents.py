"""GPU suite: the persistent agents+step kernel against the oracle driven by the same Philox contract.
RandomAgents involve only integer and f32-compare arithmetic, so the whole market history, order
table and trade log must match bit for bit.  MomentumAgent / NoiseAgent go through f64 tanh/exp/log/cos
whose last-ulp behaviour may differ between CUDA and glibc, but a differing ulp only matters when it
flips a probability compare or moves a price across a tick boundary (~1e-14 per draw): measured on
B200, 96 of 96 envs x 1000 env-steps of the C4 population and 96 of 96 x 600 steps of the noise
population are bit-identical (scripts/dbg_mom_parity.py), so these tests assert identity on EVERY env."""
import numpy as np
import pytest

from bourse_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def oracle_run(oracle, groups, seed, env_id, n_steps, tick=1, step_size=1_000_000):
    env = oracle.StepEnvNumpy(0, 0, tick, step_size)
    env.set_groups(groups)
    env.run_agents(n_steps, seed, env_id=env_id, keyed=True)
    return env


@pytest.mark.parametrize("obs_words", [abi.OBS_L1, abi.OBS_L2])
def test_random_agents_bit_exact(core, oracle, obs_words):
    n_envs, n_steps, seed = 48, 64, 101
    groups = workloads.c3_groups()
    env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=obs_words, env_id_base=1000, max_orders=8192,
                          max_trades=16384, max_steps=n_steps, max_queue=128)
    env.set_agents(groups)
    env.run_agents(n_steps, seed)
    assert not env.env_errors().any()
    hist = env.history_all(n_steps)
    total_instr = 0
    for e in range(n_envs):
        ce = oracle_run(oracle, groups, seed, 1000 + e, n_steps)
        h = ce._history()
        assert np.array_equal(hist[e], h[:, :obs_words]), e
        assert env.get_trades(e) == ce.get_trades(), e
        assert env.get_orders(e) == ce.get_orders(), e
        total_instr += ce.n_instructions()
    st = env.stats()
    assert st["instructions"] == total_instr and st["env_steps"] == n_envs * n_steps
    assert st["trades"] > 1000


def test_run_in_pieces_equals_one_run(core):
    """Idempotence of the launch boundary: 3 launches of 8/5/19 steps == one launch of 32 (state fully
    round-trips through HBM, unaligned history offsets take the direct-store path)."""
    groups = workloads.c3_groups()
    def make():
        e = core.BatchedEnv(16, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=4096, max_trades=8192, max_steps=32,
                            max_queue=128)
        e.set_agents(groups)
        return e
    a, b = make(), make()
    a.run_agents(32, 9)
    for k in (8, 5, 19):
        b.run_agents(k, 9)
    assert np.array_equal(a.history_all(32), b.history_all(32))
    for e in range(16):
        assert a.get_trades(e) == b.get_trades(e) and a.get_orders(e) == b.get_orders(e)
    assert a.stats() == b.stats()


def test_sharding_invariance(core):
    """Envs keyed by GLOBAL id: 2 shards of 8 envs == 1 handle of 16."""
    groups = workloads.c3_groups()
    def make(n, base):
        e = core.BatchedEnv(n, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, env_id_base=base, max_orders=2048, max_trades=4096,
                            max_steps=16, max_queue=128)
        e.set_agents(groups)
        e.run_agents(16, 4)
        return e.history_all(16)
    whole = make(16, 0)
    assert np.array_equal(whole[:8], make(8, 0)) and np.array_equal(whole[8:], make(8, 8))


def test_momentum_agents_match_oracle(core, oracle):
    n_envs, n_steps, seed = 40, 80, 7
    groups = workloads.c4_groups()
    env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768,
                          max_steps=n_steps, max_queue=256)
    env.set_agents(groups)
    env.run_agents(n_steps, seed)
    assert not env.env_errors().any()
    hist = env.history_all(n_steps)
    same = 0
    tv_gpu, tv_cpu, n_mom = 0, 0, 0
    for e in range(n_envs):
        ce = oracle_run(oracle, groups, seed, e, n_steps)
        h = ce._history()
        same += int(np.array_equal(hist[e], h) and env.get_orders(e) == ce.get_orders())
        tv_gpu += int(hist[e][:, 0].sum()); tv_cpu += int(h[:, 0].sum())
        n_mom += sum(1 for o in ce.get_orders() if o[7] >= 80)
    assert n_mom > 100, "momentum traders must actually trade in this config"
    assert same == n_envs, f"only {same}/{n_envs} envs identical"   # bit-exact: histories and order tables of every env
    assert tv_gpu == tv_cpu


def test_noise_agents_match_oracle(core, oracle):
    """NoiseAgent (noise_agent.rs:126-177) mixed with RandomAgents; bit-exact like the
    MomentumAgent test (the log-normal price goes through f64 exp/log/cos)."""
    n_envs, n_steps, seed = 32, 60, 3
    groups = [core.random_group(40, (40, 60), (10, 20), 2, 0.8), core.noise_group(100, 30, 2, 0.2, 0.2, 0.1, 15, 0.0, 1.0)]
    ogroups = [oracle.random_group(40, (40, 60), (10, 20), 2, 0.8), oracle.noise_group(100, 30, 2, 0.2, 0.2, 0.1, 15, 0.0, 1.0)]
    env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768,
                          max_steps=n_steps, max_queue=256)
    env.set_agents(groups)
    env.run_agents(n_steps, seed)
    assert not env.env_errors().any()
    hist = env.history_all(n_steps)
    same, n_noise = 0, 0
    for e in range(n_envs):
        ce = oracle_run(oracle, ogroups, seed, e, n_steps)
        same += int(np.array_equal(hist[e], ce._history()) and env.get_orders(e) == ce.get_orders()
                    and env.get_trades(e) == ce.get_trades())
        n_noise += sum(1 for o in ce.get_orders() if o[7] >= 100)
    assert n_noise > 1000
    assert same == n_envs, f"only {same}/{n_envs} envs identical"


def test_noise_agents_on_empty_book_far_prices(core, oracle):
    """From an empty book the mid price is 2^31 - 0.5: orders land around 2^31, far from any other page."""
    g = [core.noise_group(10, 10, 2, 1.0, 0.0, 1.0, 100, 0.0, 10.0)]
    og = [oracle.noise_group(10, 10, 2, 1.0, 0.0, 1.0, 100, 0.0, 10.0)]
    env = core.BatchedEnv(4, 0, 0, 1, 1_000_000, max_orders=1024, max_trades=1024, max_steps=8, max_queue=64, pages_total=32)
    env.set_agents(g)
    env.run_agents(2, 101)
    assert not env.env_errors().any()
    ok = 0
    for e in range(4):
        ce = oracle_run(oracle, og, 101, e, 2)
        ok += int(env.get_orders(e) == ce.get_orders() and np.array_equal(env.history(e), ce._history()))
        st = [o[1] for o in env.get_orders(e)]
        assert st[:10] == [3] * 10 and st[10:] == [1] * 10
    assert ok == 4


def test_live_list_overflow_is_flagged_on_every_env(core, oracle):
    """A MomentumAgent whose limit orders never fill (one-sided books) outgrows the 254-entry live-order list.  The
    overflow is detected per trader lane; every lane's bits must reach the env's error word (they used to be taken from
    lane 0 only), so results never differ from the oracle silently.  Found by scripts/soak.py, seed 14177."""
    M = (1000, 17, 1, 0.02, 10, 0.41, 6.35, 0.75, 1.82, 0.0, 0.98)
    N = (2000, 15, 1, 0.385, 0.198, 0.315, 7, 0.66, 1.07)
    n_envs, n_steps = 8, 60
    e = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768, max_steps=64,
                        max_queue=512, pages_smem=64, pages_total=64)
    e.set_agents([core.momentum_group(*M), core.noise_group(*N)])
    e.run_agents(n_steps, 5, sync=False)
    with pytest.raises(MemoryError, match="0x80"):
        e.synchronize(); e.stats(); e.level_2_data(); e.step(1)   # the next synchronous launch-check reports the flag
    err = e.env_errors()
    n_bad = 0
    for env in range(n_envs):
        o = oracle.StepEnvNumpy(0, 0, 1, 1_000_000)
        o.set_groups([oracle.momentum_group(*M), oracle.noise_group(*N)])
        o.run_agents(n_steps, 5, env_id=env, keyed=True)
        same = np.array_equal(e.history(env)[:n_steps], o._history())
        assert same or (err[env] & 0x80), env       # differs => flagged
        n_bad += not same
    assert n_bad > 0 and (err & 0x80).any()


def test_live_list_capacity_is_a_parameter_of_the_group(core, oracle):
    """The reference's live-order list is a Vec (momentum_agent.rs:99-102).  The population of the overflow test above, with
    room for 4096 entries asked for on one group (it sizes every group's list), runs the same 60 steps without a flag and equals
    the oracle on every env; a second population on the same handle falls back to the default size and still overflows."""
    M = (1000, 17, 1, 0.02, 10, 0.41, 6.35, 0.75, 1.82, 0.0, 0.98)
    N = (2000, 15, 1, 0.385, 0.198, 0.315, 7, 0.66, 1.07)
    n_envs, n_steps = 8, 60
    e = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768, max_steps=64,
                        max_queue=2048, pages_smem=64, pages_total=64)
    e.set_agents([core.momentum_group(*M, live_cap=4096), core.noise_group(*N)])
    e.run_agents(n_steps, 5)
    assert not e.env_errors().any()
    longest = 0
    for env in range(n_envs):
        o = oracle.StepEnvNumpy(0, 0, 1, 1_000_000)
        o.set_groups([oracle.momentum_group(*M), oracle.noise_group(*N)])
        o.run_agents(n_steps, 5, env_id=env, keyed=True)
        assert np.array_equal(e.history(env)[:n_steps], o._history()), env
        assert e.get_orders(env) == o.get_orders() and e.get_trades(env) == o.get_trades(), env
        longest = max(longest, sum(1 for x in o.get_orders() if x[1] == 1 and 1000 <= x[7] < 1017))
    assert longest > 254    # (Active orders of the momentum traders at the end: the list did outgrow the default)
    e.reset()
    e.set_agents([core.momentum_group(*M), core.noise_group(*N)])
    e.run_agents(n_steps, 5, sync=False)
    with pytest.raises(MemoryError, match="0x80"):
        e.synchronize(); e.stats(); e.level_2_data(); e.step(1)
