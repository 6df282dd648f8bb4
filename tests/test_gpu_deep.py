"""GPU suite: the deep-book engine (csrc/deep.cuh — one CTA per book, fetch / match / retire warps, chunked array queues,
warp prefix-sum sweeps) against the oracle, bit for bit, on replayed instruction streams.

The engine's preconditions (prices inside the window, strictly increasing time, one side per level) hold on every stream
here except where a test checks that a violation is FLAGGED."""
import numpy as np
import pytest

from bourse_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def compare_book(gpu_env, env_idx, ob, obs_gpu=None, obs_cpu=None):
    go, co = gpu_env.orders_arrays(env_idx), ob.orders_arrays()
    for k in co:
        assert np.array_equal(co[k], go[k]), k
    gt, ct = gpu_env.trades_arrays(env_idx), ob.trades_arrays()
    for k in ct:
        assert np.array_equal(ct[k], gt[k]), k
    assert list(gpu_env.book_level_1(env_idx)) == ob._l1()
    assert np.array_equal(gpu_env.book_level_2(env_idx), ob.level_2_data())
    if obs_cpu is not None:
        assert np.array_equal(obs_gpu, obs_cpu)


def deep_env(core, n_envs, window, n, tick=1, chunks=None, obs_words=abi.OBS_L2, **kw):
    return core.BatchedEnv(n_envs, 5, 0, tick, 1000, obs_words=obs_words, max_orders=n + 64, max_trades=4 * n + 64, max_steps=n // 32 + 64,
                           max_queue=32, price_window=window, deep_chunks=chunks or (n // 8 + 2 * (window[1] - window[0]) + 64), **kw)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c2_stream_bit_exact(core, oracle, seed):
    """The C2 mix (55% limit, 5% market, 25% cancel, 15% modify; targets uniform over all issued ids, so many are dead)
    with strictly increasing time.  (Zero-volume orders, SURVEY N5, rest and get swept in the sweep test below; a zero-volume
    order that CROSSES rests inside the other side on the reference — that is the flagged `locked level` precondition.)"""
    n = 20000
    s = workloads.replay_stream(n, seed, tick_size=1, trading_windows=False)
    ob = oracle.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=n)
    env = deep_env(core, 1, (896, 1152), n)
    env.replay(s)
    assert not env.env_errors().any()
    compare_book(env, 0, ob, env.history(0), obs_cpu)
    assert len(ob.get_trades()) > 1000


def sweep_stream(n_rest, n_sweeps, seed, levels=6, mid=500, big=4000):
    """Many small resting orders on a few levels, a third of them cancelled again (tombstones anywhere in the queues,
    including at their heads), then large aggressive orders that each sweep many orders, chunks and levels."""
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n_rest):
        bid = rng.random() < 0.5
        off = int(rng.integers(1, levels + 1))
        rows.append((abi.OP_NEW | (abi.F_BID if bid else 0), 0, mid - off if bid else mid + off, int(rng.integers(0, 9))))
        if rng.random() < 0.33:
            rows.append((abi.OP_CANCEL, int(rng.integers(0, i + 1)), 0, 0))
        if rng.random() < 0.05:
            rows.append((abi.OP_MODIFY | abi.F_HAS_VOL, int(rng.integers(0, i + 1)), 0, int(rng.integers(0, 12))))
    for _ in range(n_sweeps):
        bid = rng.random() < 0.5
        kind = rng.random()
        vol = int(rng.integers(1, big))
        if kind < 0.4:      # market order
            rows.append((abi.OP_NEW | abi.F_MARKET | (abi.F_BID if bid else 0), 0, 0, vol))
        elif kind < 0.8:    # limit order crossing some of the levels, remainder rests on the other side
            rows.append((abi.OP_NEW | (abi.F_BID if bid else 0), 0, mid + int(rng.integers(-levels, levels + 1)), vol))
        else:               # refill
            for _k in range(40):
                b2 = rng.random() < 0.5
                off = int(rng.integers(1, levels + 1))
                rows.append((abi.OP_NEW | (abi.F_BID if b2 else 0), 0, mid - off if b2 else mid + off, int(rng.integers(1, 9))))
    out = np.zeros(len(rows), dtype=abi.INSTR_DTYPE)
    for i, (of, oid, price, vol) in enumerate(rows):
        out[i] = (i + 1, of | (abi.F_EMIT if i % 50 == 49 else 0), oid, price, vol, i % 97, 0)
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_prefix_sum_sweeps_over_tombstoned_queues(core, oracle, seed):
    s = sweep_stream(6000, 300, seed)
    n = len(s)
    ob = oracle.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=n)
    env = deep_env(core, 1, (448, 576), n)
    env.replay(s)
    assert not env.env_errors().any()
    compare_book(env, 0, ob, env.history(0), obs_cpu)
    tr = ob.trades_arrays()
    # the sweeps really are multi-order: the busiest aggressor traded against more than a chunk's worth of passive orders
    assert np.bincount(tr["active"].astype(np.int64)).max() > 62


def test_c5_shaped_books_and_launch_splitting(core, oracle):
    """Small C5-shaped books (pre-loaded depth, then the 15/15/60/10 mix), 6 books in one launch; a second launch continues
    from the persisted book image and chunk pool."""
    n_rest, n_steps, per_step = 20000, 8, 1500
    streams = [workloads.c5_stream(n_rest, n_steps, per_step, seed=40 + i, mid_ticks=3000, depth_ticks=512) for i in range(6)]
    n = len(streams[0])
    env = deep_env(core, 6, (2432, 3584), n)
    cut = n_rest + 3 * per_step + 17
    env.replay(np.concatenate([x[:cut] for x in streams]), np.arange(7, dtype=np.uint64) * cut)
    env.replay(np.concatenate([x[cut:] for x in streams]), np.arange(7, dtype=np.uint64) * (n - cut))
    assert not env.env_errors().any()
    for e in range(6):
        ob = oracle.OrderBook(0, 1)
        obs_cpu = ob.replay(streams[e], obs_cap=n_steps)
        compare_book(env, e, ob, env.history(e), obs_cpu)
        assert len(ob.get_trades()) > 2000
    st = env.stats()
    assert st["instructions"] == 6 * n and st["error_envs"] == 0


def test_level1_history_and_tick_2(core, oracle):
    n = 8000
    s = workloads.replay_stream(n, 7, tick_size=2, trading_windows=False)
    ob = oracle.OrderBook(0, 2)
    obs_cpu = ob.replay(s, obs_cap=n)
    env = deep_env(core, 1, (1792, 2304), n, tick=2, obs_words=abi.OBS_L1)
    env.replay(s)
    assert not env.env_errors().any()
    compare_book(env, 0, ob)
    assert np.array_equal(env.history(0), obs_cpu[:, :9])


def test_preconditions_are_flagged_not_silent(core):
    def stream(rows):
        out = np.zeros(len(rows), dtype=abi.INSTR_DTYPE)
        for i, (t, of, oid, price, vol) in enumerate(rows):
            out[i] = (t, of, oid, price, vol, 0, 0)
        return out

    # a resting price outside the window
    env = deep_env(core, 1, (100, 164), 64)
    with pytest.raises(MemoryError):
        env.replay(stream([(1, abi.OP_NEW | abi.F_BID, 0, 90, 5)]))
    assert int(env.env_errors()[0]) & 0x04
    # time standing still between two resting inserts at one level (the reference's equal-key collision, N1)
    env = deep_env(core, 1, (100, 164), 64)
    with pytest.raises(MemoryError):
        env.replay(stream([(5, abi.OP_NEW | abi.F_BID, 0, 120, 5), (5, abi.OP_NEW | abi.F_BID, 0, 120, 6)]))
    assert int(env.env_errors()[0]) & 0x100
    # both sides resting at one price while trading is disabled
    env = deep_env(core, 1, (100, 164), 64)
    with pytest.raises(MemoryError):
        env.replay(stream([(1, abi.OP_SET_TRADING, 0, 0, 0), (2, abi.OP_NEW | abi.F_BID, 0, 120, 5), (3, abi.OP_NEW, 0, 120, 6)]))
    assert int(env.env_errors()[0]) & 0x800
    # ... or a zero-volume order whose price crosses: it never matches and rests inside the other side (N5)
    env = deep_env(core, 1, (100, 164), 64)
    with pytest.raises(MemoryError):
        env.replay(stream([(1, abi.OP_NEW, 0, 120, 6), (2, abi.OP_NEW | abi.F_BID, 0, 120, 0)]))
    assert int(env.env_errors()[0]) & 0x800
    # unknown id: the reference panics; later instructions still run
    env = deep_env(core, 1, (100, 164), 64)
    with pytest.raises(core.PanicException):
        env.replay(stream([(1, abi.OP_NEW | abi.F_BID, 0, 120, 5), (2, abi.OP_CANCEL, 9, 0, 0), (3, abi.OP_NEW, 0, 130, 4)]))
    assert list(env.book_level_1(0)[:2]) == [120, 130]
    # Env mode and agents are refused
    with pytest.raises(ValueError):
        env.step()


def test_trading_disabled_rests_and_rejects(core, oracle):
    """N6 without locked levels: limit orders rest unmatched while trading is off, market orders are rejected; when trading
    resumes a crossing order sweeps what rested."""
    rows = [(1, abi.OP_SET_TRADING, 0, 0, 0), (2, abi.OP_NEW | abi.F_BID, 0, 130, 5), (3, abi.OP_NEW, 0, 120, 6),
            (4, abi.OP_NEW | abi.F_MARKET, 0, 0, 3), (5, abi.OP_SET_TRADING, 0, 0, 1), (6, abi.OP_NEW, 0, 110, 20),
            (7, abi.OP_NEW | abi.F_BID | abi.F_MARKET, 0, 0, 4), (8, abi.OP_MODIFY | abi.F_HAS_PRICE, 2, 140, 0)]
    s = np.zeros(len(rows), dtype=abi.INSTR_DTYPE)
    for i, (t, of, oid, price, vol) in enumerate(rows):
        s[i] = (t, of | abi.F_EMIT, oid, price, vol, 3, 0)
    ob = oracle.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=len(s))
    env = deep_env(core, 1, (100, 164), 64)
    env.replay(s)
    assert not env.env_errors().any()
    compare_book(env, 0, ob, env.history(0), obs_cpu)


def test_recently_written_orders_are_not_read_stale(core, oracle):
    """Cancels / modifies that name an order created, filled or modified a handful of events earlier: the record the fetch
    warp prefetched is older than the write still travelling through the retire ring, and must not be used."""
    rng = np.random.default_rng(3)
    rows, issued = [], 0
    for i in range(12000):
        u = rng.random()
        if u < 0.5 or issued < 4:
            bid = rng.random() < 0.5
            rows.append((abi.OP_NEW | (abi.F_BID if bid else 0), 0, 200 + int(rng.integers(-6, 7)), int(rng.integers(1, 30))))
            issued += 1
        elif u < 0.75:
            rows.append((abi.OP_CANCEL, issued - 1 - int(rng.integers(0, 4)), 0, 0))
        else:
            k = rng.integers(0, 3)
            of = abi.OP_MODIFY | (abi.F_HAS_VOL if k != 1 else 0) | (abi.F_HAS_PRICE if k != 0 else 0)
            rows.append((of, issued - 1 - int(rng.integers(0, 4)), 200 + int(rng.integers(-6, 7)), int(rng.integers(1, 30))))
    s = np.zeros(len(rows), dtype=abi.INSTR_DTYPE)
    for i, (of, oid, price, vol) in enumerate(rows):
        s[i] = (i + 1, of | (abi.F_EMIT if i % 40 == 39 else 0), oid, price, vol, 1, 0)
    ob = oracle.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=len(s))
    env = deep_env(core, 1, (160, 256), len(s))
    env.replay(s)
    assert not env.env_errors().any()
    compare_book(env, 0, ob, env.history(0), obs_cpu)


def test_one_order_sweeping_more_levels_than_the_chain_list_holds(core, oracle):
    """An aggressive order whose volume the fast path takes (<= 192) but which crosses ~150 one-lot levels: the chain's output
    list fills up in the middle of the sweep, is flushed and the sweep goes on (deepw.cuh, `in_place`); the same for a market
    order, for a modify that crosses, and with cancels of the levels' orders in between."""
    rows, t = [], 0
    def add(of, oid, price, vol):
        nonlocal t
        t += 1
        rows.append((t, of, oid, price, vol))
    for rep in range(3):
        base = len([r for r in rows if (r[1] & 0xFF) == abi.OP_NEW])
        for k in range(150):
            add(abi.OP_NEW, 0, 5001 + k, 1)                      # asks, one lot per level
        for k in range(150):
            add(abi.OP_NEW | abi.F_BID, 0, 4999 - k, 1)          # bids
        for k in range(0, 150, 7):
            add(abi.OP_CANCEL, base + k, 0, 0)                   # holes in the ask ladder
        add(abi.OP_NEW | abi.F_BID, 0, 5150, 140)                # sweeps ~128 ask levels, fills completely or rests
        add(abi.OP_NEW | abi.F_MARKET, 0, 0, 150)                # market sell through the bid ladder
        add(abi.OP_NEW | abi.F_BID, 0, 4000, 30)                 # a deep resting bid ...
        add(abi.OP_MODIFY | abi.F_HAS_PRICE | abi.F_HAS_VOL, len([r for r in rows if (r[1] & 0xFF) == abi.OP_NEW]) - 1, 5150, 100)  # ... re-priced through what is left
        add(abi.OP_NEW | abi.F_MARKET | abi.F_BID, 0, 0, 190)    # and a market buy that exhausts the side
    s = np.zeros(len(rows), dtype=abi.INSTR_DTYPE)
    for i, (tt, of, oid, price, vol) in enumerate(rows):
        s[i] = (tt, of | (abi.F_EMIT if i % 97 == 96 else 0), oid, price, vol, i % 5, 0)
    ob = oracle.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=len(s))
    env = deep_env(core, 1, (3968, 5184), len(s))
    env.replay(s)
    assert not env.env_errors().any()
    compare_book(env, 0, ob, env.history(0), obs_cpu)
    tr = ob.trades_arrays()
    assert np.bincount(tr["active"].astype(np.int64)).max() > 100   # one aggressor, more than a hundred one-lot levels


def test_chunk_pool_exhaustion_is_flagged(core):
    """Fewer queue chunks than the stream needs: BB_ERR_CAP_PAGES, never a silent difference or an out-of-bounds access."""
    n = 4000
    s = np.zeros(n, dtype=abi.INSTR_DTYPE)
    for i in range(n):
        s[i] = (i + 1, abi.OP_NEW | abi.F_BID, 0, 4000 + (i % 64), 3, 0, 0)   # 64 levels x 62 orders: 2 chunks a level, plus one at birth
    env = deep_env(core, 1, (3968, 4096), n, chunks=40)
    with pytest.raises(MemoryError):
        env.replay(s)
    assert int(env.env_errors()[0]) & 0x04


def test_trade_log_overflow_is_flagged(core, oracle):
    """More trades than max_trades: BB_ERR_CAP_TRADES; the book and the order table are still the reference's (only the log is
    cut), and nothing is written past the log's end (the next env's log stays intact)."""
    n = 3000
    s = workloads.replay_stream(n, 4, tick_size=1, trading_windows=False)
    env = core.BatchedEnv(2, 5, 0, 1, 1000, obs_words=abi.OBS_L2, max_orders=n + 64, max_trades=64, max_steps=n // 32 + 64, max_queue=32,
                          price_window=(896, 1152), deep_chunks=4096)
    short = s[:40]   # env 1 gets a short stream: its log must not be touched by env 0's overflow
    with pytest.raises(MemoryError):
        env.replay(np.concatenate([s, short]), np.array([0, n, n + len(short)], dtype=np.uint64))
    errs = env.env_errors()
    assert int(errs[0]) & 0x02 and not int(errs[1])
    ob = oracle.OrderBook(0, 1)
    ob.replay(s, obs_cap=n)
    go, co = env.orders_arrays(0), ob.orders_arrays()
    for k in co:
        assert np.array_equal(co[k], go[k]), k
    ct, gt = ob.trades_arrays(), env.trades_arrays(0)
    assert len(gt["t"]) == 64 and all(np.array_equal(ct[k][:64], gt[k]) for k in ct)
    ob1 = oracle.OrderBook(0, 1)
    ob1.replay(short, obs_cap=len(short))
    assert env.get_trades(1) == ob1.get_trades() and env.get_orders(1) == ob1.get_orders()


# ---- the reference's OrderBook known answers through core.OrderBook on the deep-book engine ---------------------------------
# A scenario outside the engine's stated domain must be REFUSED with a flagged error, not mis-simulated: equal-(price, time)
# key collisions (N1: several orders at one price without advancing time).
from . import scenarios  # noqa: E402

DEEP_REFUSED = {"book_level_data": "0x100"}


@pytest.fixture(scope="module")
def deep_core(core):
    import types

    def book(*a, **k):
        return core.OrderBook(*a, **{**dict(price_window=(0, 512), deep_chunks=1024), **k})

    return types.SimpleNamespace(OrderBook=book, PanicException=core.PanicException)


@pytest.mark.parametrize("scenario", scenarios.ALL_BOOK, ids=lambda f: f.__name__)
def test_reference_known_answers_deep(deep_core, scenario):
    if scenario.__name__ in DEEP_REFUSED:
        with pytest.raises(MemoryError, match=DEEP_REFUSED[scenario.__name__]):
            scenario(deep_core)
    else:
        scenario(deep_core)


def test_snapshot_midstream_on_the_deep_engine(core, oracle, tmp_path):
    """save_json_snapshot / order_book_from_json (f1, orderbook.rs:898-905) with the book restored onto the deep-book engine:
    save after half of a C2 stream (paged engine), reload into a deep book, apply the second half to both: identical to the
    oracle that never saved — order table, trade log, level data, records."""
    s = workloads.replay_stream(6000, 21, tick_size=1, time_mode="strict", half_width=24, trading_windows=False)
    a = core.OrderBook(0, 1)
    a.replay(s[:3000])
    path = str(tmp_path / "mid.json")
    a.save_json_snapshot(path)
    b = core.order_book_from_json(path, price_window=(896, 1152), deep_chunks=4096)
    assert a.get_orders() == b.get_orders() and a.get_trades() == b.get_trades()
    assert np.array_equal(a.level_2_data(), b.level_2_data()) and a._l1() == b._l1()
    oa, ob_ = a.replay(s[3000:]), b.replay(s[3000:])
    assert np.array_equal(oa, ob_)
    ref = oracle.OrderBook(0, 1)
    ref.replay(s)
    assert b.get_trades() == ref.get_trades() and b.get_orders() == ref.get_orders()
    assert np.array_equal(b.level_2_data(), ref.level_2_data())
    # and a deep book saves a snapshot the general engine continues from
    b.save_json_snapshot(str(tmp_path / "deep.json"))
    c = core.order_book_from_json(str(tmp_path / "deep.json"))
    assert c.get_orders() == ref.get_orders() and c.get_trades() == ref.get_trades() and np.array_equal(c.level_2_data(), ref.level_2_data())


@pytest.mark.parametrize("n_books", [150, 300, 600])
def test_every_kernel_variant_against_the_oracle(core, oracle, n_books):
    """k_deepw comes in two instantiations and three filter sizes, chosen by how many books share an SM (one, two, more):
    150 books run the roomy kernel with x4 filters, 300 and 600 the compact one (generic placement code, compact filters;
    600 also goes beyond one wave of 4 x 148 CTAs).  C5-shaped books, 8 distinct streams, the books that share a stream must be
    identical and the distinct ones equal the oracle bit for bit."""
    n_rest, n_steps, per_step, n_distinct = 6000, 4, 900, 8
    streams = [workloads.c5_stream(n_rest, n_steps, per_step, seed=70 + i, mid_ticks=3000, depth_ticks=96) for i in range(n_distinct)]
    n = len(streams[0])
    allb = np.concatenate([streams[b % n_distinct] for b in range(n_books)])
    env = core.BatchedEnv(n_books, 5, 0, 1, 1000, obs_words=abi.OBS_L2, max_orders=n + 64, max_trades=2 * n + 64, max_steps=n_steps + 8,
                          max_queue=32, price_window=(2880, 3136), deep_chunks=n // 8 + 1024)
    cut = n_rest + per_step + 333
    off = np.arange(n_books + 1, dtype=np.uint64)
    env.replay(np.concatenate([streams[b % n_distinct][:cut] for b in range(n_books)]), off * cut)
    env.replay(np.concatenate([streams[b % n_distinct][cut:] for b in range(n_books)]), off * (n - cut))
    del allb
    assert not env.env_errors().any()
    hist = env.history_all(n_steps)
    l2 = env.level_2_data()
    for b in range(n_distinct, n_books):
        assert np.array_equal(hist[b], hist[b % n_distinct]) and np.array_equal(l2[b], l2[b % n_distinct]), b
        assert env.n_orders(b) == env.n_orders(b % n_distinct) and env.n_trades(b) == env.n_trades(b % n_distinct), b
    for i in list(range(n_distinct)) + [n_books - 1]:
        ob = oracle.OrderBook(0, 1)
        obs = ob.replay(streams[i % n_distinct], obs_cap=n_steps)
        compare_book(env, i, ob, hist[i], obs)
        assert len(ob.get_trades()) > 500
