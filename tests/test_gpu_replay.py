"""GPU suite: bit-exact parity of the CUDA matching engine with the oracle on replayed instruction
streams (BASELINE config C2), including the reference's edge semantics (N1-N6)."""
import numpy as np
import pytest

from bourse_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def compare_book(gpu_env, env_idx, ob, obs_gpu=None, obs_cpu=None):
    assert gpu_env.get_trades(env_idx) == ob.get_trades()
    assert gpu_env.get_orders(env_idx) == ob.get_orders()
    assert list(gpu_env.book_level_1(env_idx)) == ob._l1()
    assert np.array_equal(gpu_env.book_level_2(env_idx), ob.level_2_data())
    if obs_cpu is not None:
        assert np.array_equal(obs_gpu, obs_cpu)


@pytest.mark.parametrize("time_mode", ["strict", "flat", "jitter"])
@pytest.mark.parametrize("tick_size,seed", [(1, 0), (2, 1), (1, 2)])
def test_single_book_replay_bit_exact(core, oracle, time_mode, tick_size, seed):
    n = 20000
    s = workloads.replay_stream(n, seed, tick_size=tick_size, time_mode=time_mode, min_vol=0 if seed == 2 else 1)
    ob = oracle.OrderBook(0, tick_size)
    obs_cpu = ob.replay(s, obs_cap=n)
    g = core.OrderBook(0, tick_size, max_orders=n, max_trades=4 * n, max_steps=n // 64 + 8)
    obs_gpu = g.replay(s)
    compare_book(g._env, 0, ob, obs_gpu, obs_cpu)
    assert len(ob.get_trades()) > 100
    assert not g._env.env_errors().any()


def test_wide_price_range_uses_hbm_pages(core, oracle):
    """Prices spread over far more than the shared-memory resident pages: overflow pages in HBM."""
    n = 6000
    s = workloads.replay_stream(n, 11, tick_size=1, half_width=900, trading_windows=False)
    ob = oracle.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=n)
    g = core.OrderBook(0, 1, max_orders=n, max_trades=2 * n, pages_smem=4, pages_total=128)
    obs_gpu = g.replay(s)
    compare_book(g._env, 0, ob, obs_gpu, obs_cpu)


@pytest.mark.parametrize("pages", [64, 40, 200], ids=["64_pages", "40_pages_not_x32", "200_pages_2_books_per_cta"])
def test_all_pages_resident_geometry(core, oracle, pages):
    """pages_smem == pages_total selects k_apply<.., ENG_PAGED_RES> (no HBM page variants); the image of a 200-page book
    (103 KB) makes the launch use two books per CTA.  Wide price range, modifies, market orders, trading windows."""
    n = 12000
    streams = [workloads.replay_stream(n, 20 + i, tick_size=1, half_width=500) for i in range(3)]
    env = core.BatchedEnv(3, 5, 0, 1, 1000, max_orders=n + 64, max_trades=4 * n, max_steps=n // 64 + 16, max_queue=32,
                          pages_smem=pages, pages_total=pages)
    env.replay(np.concatenate(streams), np.arange(4, dtype=np.uint64) * n)
    assert not env.env_errors().any()
    books = []
    for e in range(3):
        ob = oracle.OrderBook(0, 1)
        obs_cpu = ob.replay(streams[e], obs_cap=n)
        compare_book(env, e, ob, env.history(e), obs_cpu)
        books.append(ob)
    assert sum(len(b.get_trades()) for b in books) > 1000


def test_extreme_prices(core, oracle):
    """Orders at the ends of the u32 price range, market sentinels as limit prices (N3), wrapping L2 levels."""
    M = 2**32 - 1
    def run(ob):
        ob.place_order(False, 5, 0, price=M)        # limit ask at u32::MAX
        ob.set_time(1); ob.place_order(True, 4, 0, price=0)   # bid at 0
        ob.set_time(2); ob.place_order(True, 3, 0, price=M)   # limit bid at MAX == market bid (N3)
        ob.set_time(3); ob.place_order(False, 9, 0, price=M - 1)
        ob.set_time(4); ob.place_order(False, 2, 0, price=0)  # limit ask at 0 == market ask
        ob.set_time(5); ob.modify_order(3, new_price=1)
        return ob.get_trades(), ob.get_orders(), list(ob.level_2_data())
    assert run(core.OrderBook(0, 1, pages_total=64)) == run(oracle.OrderBook(0, 1))


def test_many_books_replay(core, oracle):
    """64 books, each with its own stream, one launch; every book bit-exact."""
    n_envs, n = 64, 3000
    streams = [workloads.replay_stream(n + 17 * e, 100 + e, tick_size=1 + (e % 2), half_width=16 + e % 7) for e in range(n_envs)]
    # one handle has one tick size: split by tick
    for tick in (1, 2):
        idx = [e for e in range(n_envs) if 1 + (e % 2) == tick]
        offs = np.zeros(len(idx) + 1, np.uint64)
        offs[1:] = np.cumsum([len(streams[e]) for e in idx])
        env = core.BatchedEnv(len(idx), 0, 0, tick, 1, max_orders=4096, max_trades=16384, max_steps=256, pages_total=32)
        env.replay(np.concatenate([streams[e] for e in idx]), offs)
        for k, e in enumerate(idx):
            ob = oracle.OrderBook(0, tick)
            obs_cpu = ob.replay(streams[e], obs_cap=len(streams[e]))
            compare_book(env, k, ob, env.history(k), obs_cpu)
        assert not env.env_errors().any()
        st = env.stats()
        assert st["error_envs"] == 0 and st["trades"] > 0


@pytest.mark.parametrize("kw", [dict(), dict(pages_smem=48, pages_total=48), dict(pages_smem=4, pages_total=64)],
                         ids=["fast", "all_resident", "hbm_pages"])
def test_env_mode_batched_bit_exact(core, oracle, kw):
    """Host-queued instructions + Env::step over 32 envs x 30 steps with cancels and modifies: the CUDA
    env and the oracle env share the Xoroshiro shuffle stream, so everything must match exactly (on the single-directory
    geometry, the all-resident one — k_apply<ENV, ENG_PAGED_RES> — and the generic one with HBM pages)."""
    n_envs, n_steps = 32, 30
    rng = np.random.default_rng(5)
    genv = core.BatchedEnv(n_envs, 77, 0, 1, 1000, max_orders=4096, max_trades=8192, max_steps=64, max_queue=128, **kw)
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
    l2 = genv.level_2_data()
    for e, ce in enumerate(cenvs):
        assert genv.get_trades(e) == ce.get_trades(), e
        assert genv.get_orders(e) == ce.get_orders(), e
        assert np.array_equal(genv.history(e), ce._history()), e
        assert np.array_equal(l2[e], ce._l2()), e
        assert genv.time(e) == n_steps * 1000
    assert sum(len(ce.get_trades()) for ce in cenvs) > 100


@pytest.mark.parametrize("pages", [(8, 32), (32, 32), (64, 64)], ids=["hbm_pages", "single_directory", "all_resident"])
@pytest.mark.parametrize("tick_size", [1, 2])
def test_adversarial_fuzz_many_books(core, oracle, tick_size, pages):
    """768 books, each replaying its own short ADVERSARIAL stream in one launch (tiny price / id / time domains:
    equal-key collisions N1, zero volumes N5, market sentinels N3, trading toggles N6, every modify variant N4, prices at
    both ends of the u32 range); every book must equal the oracle — trade log, order table, every emitted record."""
    from .test_oracle_hypothesis import random_adversarial_stream

    rng = np.random.default_rng(20 + tick_size)
    n_books = 768
    streams = [random_adversarial_stream(rng, int(rng.integers(1, 120)), tick_size) for _ in range(n_books)]
    off = np.zeros(n_books + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in streams])
    # price_granule=1: modify_order applies no tick check (N4), so with tick 2 an order may come to rest on an odd price
    env = core.BatchedEnv(n_books, 0, 0, tick_size, 1000, obs_words=abi.OBS_L2, max_orders=128, max_trades=1024, max_steps=128,
                          max_queue=16, pages_smem=pages[0], pages_total=pages[1], price_granule=1)
    env.replay(np.concatenate(streams), off)
    assert not env.env_errors().any()
    n_tr = 0
    for e in range(n_books):
        ob = oracle.OrderBook(0, tick_size)
        obs = ob.replay(streams[e], obs_cap=len(streams[e]))
        assert np.array_equal(env.history(e), obs), e
        assert env.get_orders(e) == ob.get_orders(), e
        assert env.get_trades(e) == ob.get_trades(), e
        n_tr += len(ob.get_trades())
    assert n_tr > 2000
