"""GPU suite for the dense-window engine (csrc/dense.cuh; bb_config.win_levels > 0): same parity bar as the paged
engine — bit-exact against the oracle — plus dense == paged on identical input, and the engine's three preconditions
(window, resting-order slots, time order) must flag an env error instead of changing results."""
import types

import numpy as np
import pytest

from bourse_b200 import abi, workloads

from . import scenarios

pytestmark = pytest.mark.gpu

WINDOW = (0, 512)


@pytest.fixture(scope="module")
def dense_core(core):
    """`bourse.core`-shaped namespace whose classes run on the dense engine."""
    kw = dict(price_window=WINDOW, live_cap=254)

    def wrap(cls):
        return lambda *a, **k: cls(*a, **{**kw, **k})

    return types.SimpleNamespace(OrderBook=wrap(core.OrderBook), StepEnv=wrap(core.StepEnv), StepEnvNumpy=wrap(core.StepEnvNumpy),
                                 PanicException=core.PanicException)


# scenarios that place several orders at one price without advancing time: the reference's equal-(price, time) key
# collisions are outside the dense engine's domain and must be refused, not mis-simulated
COLLISION_SCENARIOS = {"book_level_data"}


@pytest.mark.parametrize("scenario", scenarios.ALL_BOOK + scenarios.ALL_ENV + scenarios.ALL_NUMPY, ids=lambda f: f.__name__)
def test_reference_known_answers_dense(dense_core, scenario):
    if scenario.__name__ in COLLISION_SCENARIOS:
        with pytest.raises(MemoryError, match="0x100"):
            scenario(dense_core)
    else:
        scenario(dense_core)


def _oracle_env(oracle, groups, seed, env_id, n_steps):
    env = oracle.StepEnvNumpy(0, 0, 1, 1_000_000)
    env.set_groups(groups)
    env.run_agents(n_steps, seed, env_id=env_id, keyed=True)
    return env


@pytest.mark.parametrize("obs_words", [abi.OBS_L1, abi.OBS_L2])
def test_random_agents_bit_exact_dense(core, oracle, obs_words):
    n_envs, n_steps, seed = 48, 64, 101
    groups = workloads.c3_groups()
    env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=obs_words, env_id_base=1000, max_orders=8192, max_trades=16384,
                          max_steps=n_steps, max_queue=128, price_window=(0, 192), live_cap=128)
    env.set_agents(groups)
    env.run_agents(n_steps, seed)
    assert not env.env_errors().any()
    hist = env.history_all(n_steps)
    total = 0
    for e in range(n_envs):
        ce = _oracle_env(oracle, groups, seed, 1000 + e, n_steps)
        assert np.array_equal(hist[e], ce._history()[:, :obs_words]), e
        assert env.get_trades(e) == ce.get_trades(), e
        assert env.get_orders(e) == ce.get_orders(), e
        total += ce.n_instructions()
    st = env.stats()
    assert st["instructions"] == total and st["env_steps"] == n_envs * n_steps and st["trades"] > 1000
    assert np.array_equal(hist[:, -1, :], env.level_1_data() if obs_words == abi.OBS_L1 else env.level_2_data())


def test_dense_equals_paged_and_launch_splitting(core):
    """Same population, same seed: the two engines must produce identical histories, order tables and trade logs;
    and three dense launches of 8/5/19 steps equal one of 32 (the whole slot table round-trips through HBM)."""
    groups = workloads.c3_groups()

    def make(**kw):
        e = core.BatchedEnv(16, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=4096, max_trades=8192, max_steps=32,
                            max_queue=128, **kw)
        e.set_agents(groups)
        return e
    a, b, c = make(), make(price_window=(0, 192)), make(price_window=(16, 208), live_cap=100)
    a.run_agents(32, 9)
    b.run_agents(32, 9)
    for k in (8, 5, 19):
        c.run_agents(k, 9)
    for x in (b, c):
        assert not x.env_errors().any()
        assert np.array_equal(a.history_all(32), x.history_all(32))
        for e in range(16):
            assert a.get_trades(e) == x.get_trades(e) and a.get_orders(e) == x.get_orders(e)
        assert a.stats() == x.stats()


@pytest.mark.parametrize("tick_size,seed", [(1, 0), (2, 1), (1, 2)])
def test_shallow_replay_bit_exact_dense(core, oracle, tick_size, seed):
    """Replayed place / cancel / modify / market mix incl. trading-off windows on ONE dense book vs the oracle."""
    s = workloads.shallow_replay_stream(30_000, seed, tick_size)
    ob = oracle.OrderBook(0, tick_size)
    obs_cpu = ob.replay(s, obs_cap=len(s))
    lo = (1000 - 32) * tick_size
    g = core.OrderBook(0, tick_size, max_orders=1 << 15, max_trades=1 << 16, max_steps=1024, price_window=(lo, lo + 64 * tick_size),
                       live_cap=254)
    obs_gpu = g.replay(s)
    assert np.array_equal(obs_gpu, obs_cpu)
    co, go = ob.orders_arrays(), g._env.orders_arrays(0)
    for k in co:
        assert np.array_equal(co[k], go[k]), k
    ct, gt = ob.trades_arrays(), g._env.trades_arrays(0)
    assert len(ct["vol"]) > 5000
    for k in ct:
        assert np.array_equal(ct[k], gt[k]), k
    assert np.array_equal(g.level_2_data(), ob.level_2_data())
    assert g.bid_ask() == ob.bid_ask() and g.best_bid_vol_and_orders() == ob.best_bid_vol_and_orders()
    assert g.best_ask_vol_and_orders() == ob.best_ask_vol_and_orders()


def test_many_books_replay_dense_equals_paged(core):
    n_books = 64
    streams = [workloads.shallow_replay_stream(3000, 50 + i) for i in range(n_books)]
    offs = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.uint64)
    allin = np.concatenate(streams)
    envs = []
    for kw in (dict(pages_smem=10, pages_total=64), dict(price_window=(960, 1056), live_cap=200)):
        e = core.BatchedEnv(n_books, 0, 0, 1, 1, max_orders=4096, max_trades=8192, max_steps=64, max_queue=32, **kw)
        e.replay(allin, offs)
        assert not e.env_errors().any()
        envs.append(e)
    a, b = envs
    assert np.array_equal(a.level_2_data(), b.level_2_data())
    for e in range(n_books):
        assert np.array_equal(a.history(e), b.history(e)), e
        assert a.get_trades(e) == b.get_trades(e) and a.get_orders(e) == b.get_orders(e), e


def test_env_mode_batched_bit_exact_dense(core, oracle):
    """Host-queued instructions + Env::step (Xoroshiro shuffle) over 32 envs x 30 steps with cancels, modifies and market
    orders on the dense engine vs the oracle env."""
    n_envs, n_steps = 32, 30
    rng = np.random.default_rng(5)
    genv = core.BatchedEnv(n_envs, 77, 0, 1, 1000, max_orders=4096, max_trades=8192, max_steps=64, max_queue=128,
                           price_window=(64, 128), live_cap=254)
    cenvs = [oracle.StepEnv(77 + e, 0, 1, 1000) for e in range(n_envs)]
    issued = np.zeros(n_envs, np.int64)
    for step in range(n_steps):
        m = 40
        env_idx = rng.integers(0, n_envs, size=m).astype(np.uint32)
        u = rng.random(m)
        action = np.where(u < 0.6, 1, np.where(u < 0.85, 2, 3)).astype(np.uint32)
        side = rng.random(m) < 0.5
        vol = rng.integers(1, 50, size=m, dtype=np.uint32)
        price = rng.integers(90, 111, size=m, dtype=np.uint32)
        flags = np.zeros(m, np.uint32)
        kinds = rng.integers(0, 4, size=m)
        oid = np.zeros(m, np.uint64)
        for r in range(m):
            e = env_idx[r]
            if action[r] != 1:
                if issued[e] == 0:
                    action[r] = 0
                else:
                    oid[r] = rng.integers(0, issued[e])
            if action[r] == 1:
                if kinds[r] == 0:
                    flags[r] = abi.F_MARKET
                issued[e] += 1
            elif action[r] == 3:
                flags[r] = [abi.F_HAS_VOL, abi.F_HAS_PRICE, abi.F_HAS_VOL | abi.F_HAS_PRICE, abi.F_HAS_VOL][kinds[r]]
        ids = genv.submit(action, side, vol, np.arange(m, dtype=np.uint32), price, oid, env_idx, flags)
        for r in range(m):
            ce = cenvs[env_idx[r]]
            if action[r] == 1:
                cid = ce.place_order(bool(side[r]), int(vol[r]), r, None if flags[r] & abi.F_MARKET else int(price[r]))
                assert cid == ids[r]
            elif action[r] == 2:
                ce.cancel_order(int(oid[r]))
            elif action[r] == 3:
                ce.modify_order(int(oid[r]), int(price[r]) if flags[r] & abi.F_HAS_PRICE else None,
                                int(vol[r]) if flags[r] & abi.F_HAS_VOL else None)
        genv.step()
        for ce in cenvs:
            ce.step()
    assert not genv.env_errors().any()
    l2 = genv.level_2_data()
    for e, ce in enumerate(cenvs):
        assert genv.get_trades(e) == ce.get_trades(), e
        assert genv.get_orders(e) == ce.get_orders(), e
        assert np.array_equal(genv.history(e), ce._history()), e
        assert np.array_equal(l2[e], ce._l2()), e
    assert sum(len(ce.get_trades()) for ce in cenvs) > 100


def test_dense_preconditions_are_flagged(core):
    # (1) resting price outside the window
    ob = core.OrderBook(0, 1, price_window=(100, 164))
    ob.place_order(True, 5, 0, price=120)
    with pytest.raises(MemoryError, match="0x4"):
        ob.place_order(False, 5, 0, price=200)
    # a crossing (non-resting) order outside the window is fine: it only matches
    ob = core.OrderBook(0, 1, price_window=(100, 164))
    ob.place_order(True, 5, 0, price=120)
    ob.set_time(1)
    ob.place_order(False, 5, 1, price=3)
    assert ob.get_trades() == [(1, True, 120, 5, 1, 0)] and ob.bid_ask() == (0, 2**32 - 1)
    # (2) more resting orders than slots
    ob = core.OrderBook(0, 1, price_window=(100, 164), live_cap=8)
    for i in range(8):
        ob.set_time(i + 1)
        ob.place_order(True, 1, 0, price=110)
    ob.set_time(20)
    with pytest.raises(MemoryError, match="0x80"):
        ob.place_order(True, 1, 0, price=111)
    # (3) a resting order at the same (price, time) as the level's tail: key collision of the reference (N1)
    ob = core.OrderBook(0, 1, price_window=(100, 164))
    ob.place_order(True, 1, 0, price=110)
    ob.place_order(True, 1, 0, price=111)      # same time, other level: fine
    with pytest.raises(MemoryError, match="0x100"):
        ob.place_order(True, 1, 0, price=110)
    # slots are recycled: 8 slots carry any number of orders as long as at most 8 rest at once
    ob = core.OrderBook(0, 1, price_window=(100, 164), live_cap=8)
    for i in range(100):
        ob.set_time(2 * i + 1)
        a = ob.place_order(True, 3, 0, price=110 + (i % 5))
        ob.set_time(2 * i + 2)
        if i % 2:
            ob.cancel_order(a)
        else:
            ob.place_order(False, 3, 1, price=100)
    assert ob.bid_vol() == 0 and len(ob.get_trades()) == 50


def test_snapshot_round_trip_dense(core, oracle, tmp_path):
    """JSON snapshot taken mid-stream from a dense book, loaded into a fresh dense book AND a paged book: all three
    continue identically (queue order inside each level is rebuilt from the stored key times)."""
    from bourse_b200 import snapshot
    s = workloads.shallow_replay_stream(6000, 11)
    kw = dict(max_orders=1 << 14, max_trades=1 << 15, max_steps=256)
    lo = 1000 - 32
    a = core.OrderBook(0, 1, price_window=(lo, lo + 64), **kw)
    a.replay(s[:3000])
    path = str(tmp_path / "book.json")
    a.save_json_snapshot(path)
    b = snapshot.order_book_from_json(path, price_window=(lo, lo + 64), **kw)
    c = snapshot.order_book_from_json(path, **kw)
    for x in (a, b, c):
        x.replay(s[3000:])
    for x in (b, c):
        assert x.get_orders() == a.get_orders() and x.get_trades() == a.get_trades()
        assert np.array_equal(x.level_2_data(), a.level_2_data())
    ob = oracle.OrderBook(0, 1)
    ob.replay(s, obs_cap=len(s))
    assert a.get_orders() == ob.get_orders() and a.get_trades() == ob.get_trades()


@pytest.mark.parametrize("kw", [dict(), dict(price_window=(0, 192))], ids=["paged", "dense"])
def test_run_agents_to_host_streams_the_same_history(core, kw):
    """bb_run_agents_to_host (chunked launches, records copied out while the next chunk runs) == bb_run_agents followed
    by bb_history_all, including a chunk size that does not divide the run."""
    groups = workloads.c3_groups()

    def make():
        e = core.BatchedEnv(24, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=4096, max_trades=8192, max_steps=64,
                            max_queue=128, **kw)
        e.set_agents(groups)
        return e
    a, b = make(), make()
    a.run_agents(44, 5)
    ref = a.history_all(44)
    out = b.run_agents_to_host(44, 5, chunk_steps=16)
    assert np.array_equal(out, ref) and np.array_equal(b.history_all(44), ref)
    assert a.stats() == b.stats()
    # a second streamed run continues the same history
    a.run_agents(20, 5)
    out2 = b.run_agents_to_host(20, 5)
    assert np.array_equal(out2, a.history_all(64)[:, 44:])


def test_host_instructions_between_agent_launches_dense(core):
    """The agents' on-chip slot table is rebuilt at every launch: orders cancelled / replaced from the host between two
    agent launches (which moves them to other slots or off the book) must be seen exactly as on the paged engine."""
    groups = workloads.c3_groups()

    def run(**kw):
        e = core.BatchedEnv(4, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=4096, max_trades=8192, max_steps=64,
                            max_queue=128, **kw)
        e.set_agents(groups)
        e.run_agents(12, 3)
        for env in range(4):
            o = e.orders_arrays(env)
            act = np.flatnonzero(o["status"] == 1)
            assert len(act) >= 6
            ids = act[:: max(1, len(act) // 6)][:6]
            e.submit([abi.ACT_CANCEL] * 2, order_id=ids[:2], env=[env] * 2)
            e.submit([abi.ACT_MODIFY] * 2, order_id=ids[2:4], price=o["price"][ids[2:4]], vol=[7, 9], env=[env] * 2)
            e.submit([abi.ACT_MODIFY] * 2, order_id=ids[4:6], vol=[1, 2], env=[env] * 2, flags=[abi.F_HAS_VOL] * 2)
        e.step(1)
        e.run_agents(20, 3)
        assert not e.env_errors().any()
        return e
    a, b = run(), run(price_window=(0, 192))
    assert np.array_equal(a.history_all(33), b.history_all(33))
    for env in range(4):
        assert a.get_trades(env) == b.get_trades(env) and a.get_orders(env) == b.get_orders(env)


def test_slot_overflow_stays_memory_safe(core):
    """More resting orders than slots is flagged (0x80) and the overflowing orders are left off the book; the image must
    stay consistent, so that hundreds of further matches, cancels and inserts on the flagged book cannot touch memory
    outside it (a reused live slot used to hand the matching loop a stale order id -> out-of-bounds record write)."""
    n = 600
    rng = np.random.default_rng(9)
    ins = np.zeros(n, abi.INSTR_DTYPE)
    ins["t"] = np.arange(1, n + 1)
    new = rng.random(n) < 0.8
    new[0] = True   # cancels always name an id that exists
    bid = rng.random(n) < 0.5
    ins["op_flags"] = np.where(new, abi.OP_NEW | np.where(bid, abi.F_BID, 0), abi.OP_CANCEL).astype(np.uint32)
    ins["price"] = np.where(bid, rng.integers(100, 132, n), rng.integers(124, 160, n))   # mostly resting, some crossing
    ins["vol"] = rng.integers(1, 9, n)
    ins["order_id"] = (rng.random(n) * np.maximum(1, np.cumsum(new) - 1)).astype(np.uint32)
    env = core.BatchedEnv(64, 0, 0, 1, 1000, max_orders=1024, max_trades=2048, max_steps=8, max_queue=16,
                          price_window=(96, 160), live_cap=16)
    with pytest.raises(MemoryError, match="0x80"):
        env.replay(np.tile(ins, 64), np.arange(65, dtype=np.uint64) * n)
    assert (env.env_errors() == 0x80).all()
    env.synchronize()
    # a fresh handle on the same device still works: nothing was corrupted
    ob = core.OrderBook(0, 1, price_window=(100, 164))
    ob.place_order(True, 5, 0, price=120)
    assert ob.bid_ask() == (120, 2**32 - 1)
