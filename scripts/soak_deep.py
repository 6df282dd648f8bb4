"""Randomised soak of the deep-book engine (k_deepw) against the oracle: every round draws a book shape (depth, spread zone,
volumes incl. ones beyond the fast path's limit, cancel / modify / market mix, market-data record spacing), a number of books and
a launch split, replays the streams on the GPU and compares every book bit for bit (order table, trade log, level-1, level-2,
history).  python scripts/soak_deep.py <seconds> [first_seed] [seeds...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bourse_b200 import abi, core  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def deep_stream(rng, n_rest, n_ev, mid, depth, zone, max_vol, p_cancel, p_modify, p_market, emit_every, big_every):
    """Pre-load `n_rest` non-crossing orders over `depth` ticks per side, then `n_ev` events: cancels / modifies of random ids
    issued so far, limit orders within +-zone ticks of mid (either side), market orders; every `big_every`-th order is LARGE
    (sweeps many levels / takes the serial path).  Strictly increasing time."""
    n = n_rest + n_ev
    out = np.zeros(n, dtype=abi.INSTR_DTYPE)
    side1 = rng.random(n_rest) < 0.5
    off = rng.integers(1, depth + 1, size=n_rest)
    out["op_flags"][:n_rest] = abi.OP_NEW | np.where(side1, abi.F_BID, 0).astype(np.uint32)
    out["price"][:n_rest] = np.where(side1, mid - off, mid + off).astype(np.uint32)
    out["vol"][:n_rest] = rng.integers(1, max_vol + 1, size=n_rest, dtype=np.uint32)
    u = rng.random(n_ev)
    op = np.full(n_ev, abi.OP_NEW, dtype=np.uint32)
    op[u < p_cancel + p_modify] = abi.OP_MODIFY
    op[u < p_cancel] = abi.OP_CANCEL
    if n_rest == 0:
        op[:4] = abi.OP_NEW   # (a cancel / modify needs an order to name: unknown ids make the reference panic)
    is_new = op == abi.OP_NEW
    market = is_new & (rng.random(n_ev) < p_market)
    side2 = rng.random(n_ev) < 0.5
    flags = np.zeros(n_ev, dtype=np.uint32)
    flags[is_new & side2] |= abi.F_BID
    flags[market] |= abi.F_MARKET
    kind = rng.integers(0, 3, size=n_ev)
    is_mod = op == abi.OP_MODIFY
    flags[is_mod & (kind != 1)] |= abi.F_HAS_VOL
    flags[is_mod & (kind != 0)] |= abi.F_HAS_PRICE
    issued_before = n_rest + np.cumsum(is_new) - is_new
    # targets: mostly uniform over everything issued, sometimes one of the last few orders (records still in flight)
    recent = rng.random(n_ev) < 0.15
    target = np.where(recent, np.maximum(issued_before - 1 - rng.integers(0, 6, size=n_ev), 0), np.floor(rng.random(n_ev) * np.maximum(issued_before, 1)))
    target = target.astype(np.uint32)
    price2 = rng.integers(mid - zone, mid + zone + 1, size=n_ev).astype(np.uint32)
    vol2 = rng.integers(1, max_vol + 1, size=n_ev, dtype=np.uint32)
    if big_every:
        k = np.arange(n_ev)
        big = (k % big_every) == big_every - 1
        vol2[big] = rng.integers(200, 200 + 40 * max_vol * max(depth, 4), size=int(big.sum()), dtype=np.uint32)
    k = np.arange(n_ev)
    if emit_every:
        flags[(k % emit_every) == emit_every - 1] |= abi.F_EMIT
    out["op_flags"][n_rest:] = op | flags
    out["order_id"][n_rest:] = np.where(is_new, 0, target)
    out["price"][n_rest:] = np.where(market, 0, price2)
    out["vol"][n_rest:] = vol2
    out["trader"] = (np.arange(n) % 1000).astype(np.uint32)
    out["t"] = np.arange(n, dtype=np.uint64) * np.uint64(int(rng.integers(1, 4))) + np.uint64(1)
    return out


budget, seed0 = float(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 0
only = [int(x) for x in sys.argv[3:]]
t_end = time.time() + budget
rounds = fails = flagged = 0
seed = seed0
while time.time() < t_end and (not only or rounds < len(only)):
    if only:
        seed = only[rounds]
    rng = np.random.default_rng(seed)
    rounds += 1
    depth = int(rng.choice([3, 8, 40, 200, 1000]))
    zone = int(rng.choice([1, 2, 6, 32, min(depth, 100)]))
    mid = 5000
    n_rest = int(rng.choice([0, 50, 2000, 20000]))
    n_ev = int(rng.integers(500, 12000))
    max_vol = int(rng.choice([1, 5, 50, 150, 400]))
    p_cancel, p_modify = float(rng.uniform(0, 0.35)), float(rng.uniform(0, 0.3))
    p_market = float(rng.uniform(0, 0.3))
    emit_every = int(rng.choice([0, 1, 7, 64, 1000]))
    big_every = int(rng.choice([0, 0, 5, 37, 500]))
    n_books = int(rng.integers(1, 7))
    n_cuts = int(rng.integers(0, 4))
    streams = [deep_stream(rng, n_rest, n_ev, mid, depth, zone, max_vol, p_cancel, p_modify, p_market, emit_every, big_every) for _ in range(n_books)]
    n = n_rest + n_ev
    lo = (mid - max(depth, zone) - 2) // 32 * 32
    hi = -(-(mid + max(depth, zone) + 3) // 32) * 32
    desc = (f"seed {seed} depth {depth} zone {zone} rest {n_rest} ev {n_ev} maxvol {max_vol} cancel {p_cancel:.2f} modify {p_modify:.2f} "
            f"market {p_market:.2f} emit {emit_every} big {big_every} books {n_books} cuts {n_cuts}")
    seed += 1
    try:
        env = core.BatchedEnv(n_books, 5, 0, 1, 1000, obs_words=abi.OBS_L2, max_orders=n + 64, max_trades=40 * n + 64, max_steps=n + 8, max_queue=32,
                              price_window=(lo, hi), deep_chunks=n // 4 + 2 * (hi - lo) + 256)
        cuts = sorted(set([0, n] + [int(c) for c in rng.integers(1, n, size=n_cuts)]))
        raised = None
        for a, b in zip(cuts[:-1], cuts[1:]):
            try:
                env.replay(np.concatenate([s[a:b] for s in streams]), np.arange(n_books + 1, dtype=np.uint64) * (b - a))
            except core.PanicException as e:   # (unknown ids: the reference panics; the stream goes on)
                raised = e
        errs = env.env_errors()
        if (errs & ~np.uint32(0x10)).any():
            flagged += 1
            print("FLAGGED", [hex(int(x)) for x in errs], desc, flush=True)
            env.close()
            continue
        bad = None
        for bk in range(n_books):
            ob = orc.OrderBook(0, 1)
            obs = ob.replay(streams[bk], obs_cap=n)
            go, co = env.orders_arrays(bk), ob.orders_arrays()
            for kk in co:
                if not np.array_equal(co[kk], go[kk]):
                    bad = f"book {bk} orders.{kk} first at id {np.nonzero(co[kk] != go[kk])[0][0] if len(co[kk]) == len(go[kk]) else 'len'}"
            gt, ct = env.trades_arrays(bk), ob.trades_arrays()
            for kk in ct:
                if not np.array_equal(ct[kk], gt[kk]):
                    bad = f"book {bk} trades.{kk}"
            if list(env.book_level_1(bk)) != ob._l1() or not np.array_equal(env.book_level_2(bk), ob.level_2_data()):
                bad = f"book {bk} level data"
            if not np.array_equal(env.history(bk), obs):
                bad = f"book {bk} history"
            if bad:
                break
        env.close()
        if bad:
            fails += 1
            print("MISMATCH", bad, desc, flush=True)
    except Exception as e:  # noqa: BLE001
        fails += 1
        print("EXCEPTION", repr(e)[:300], desc, flush=True)
print(f"deep soak: {rounds} rounds, {fails} parity failures, {flagged} rounds stopped by a flagged precondition / capacity")
