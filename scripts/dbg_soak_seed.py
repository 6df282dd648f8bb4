"""Reproduce one market-mode round of scripts/soak.py and localise the first difference against the oracle, on several engines.
python scripts/dbg_soak_seed.py <seed>"""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.argv = [sys.argv[0], "0", sys.argv[1]] if len(sys.argv) > 1 else [sys.argv[0], "0", "14177"]
seed = int(sys.argv[2])
from bourse_b200 import abi, core  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import importlib.util
spec = importlib.util.spec_from_file_location("soak", "scripts/soak.py")
src = open("scripts/soak.py").read().split("rounds = fails = capacity = 0")[0]
ns = {}
exec(compile(src, "soak_head", "exec"), ns)
rand_groups, build = ns["rand_groups"], ns["build"]

rng = np.random.default_rng(seed)
mode = ["single", "market", "ext"][int(rng.integers(3))]
dense = rng.random() < 0.4
kw = dict(price_window=(0, 256), live_cap=254) if dense else (dict() if rng.random() < 0.6 else dict(pages_smem=int(rng.integers(2, 12)), pages_total=64))
gs = rand_groups(rng, dense)
n_steps = int(rng.integers(5, 60)); split = int(rng.integers(1, n_steps + 1)); rs = int(rng.integers(1 << 30))
print("mode", mode, "kw", kw, "steps", n_steps, split, "seed", rs)
for k, a in gs:
    print("  group", k, a)
assert mode == "market"
A = int(rng.integers(2, 5)); n_m = int(rng.integers(1, 60)); assets = [int(rng.integers(A)) for _ in gs]
print("A", A, "markets", n_m, "assets", assets)
for name, ekw, sp in (("soak config", kw, split), ("soak config, one launch", kw, n_steps), ("all pages resident", dict(pages_smem=64, pages_total=64), split),
                      ("fast (10 pages)", dict(), split)):
    e = core.BatchedEnv(n_m * A, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768, max_steps=64, max_queue=512, assets=A, **ekw)
    e.set_agents(build(core, gs), assets=assets); e.run_agents(sp, rs)
    if n_steps > sp:
        e.run_agents(n_steps - sp, rs)
    errs = e.env_errors()
    bad = []
    for m in range(n_m):
        o = orc.MarketEnv(0, 0, [1] * A, 1_000_000); o.set_groups(build(orc, gs), assets); o.run_agents(n_steps, rs, market_id=m, keyed=True)
        for a in range(A):
            hg, ho = e.history(m * A + a), o.history(a)
            if not np.array_equal(hg, ho):
                step = int(np.flatnonzero((hg != ho).any(axis=1))[0])
                bad.append((m, a, step))
                if len(bad) == 1:
                    go, oo = e.get_orders(m * A + a), o.get_orders(a)
                    k = next((i for i in range(min(len(go), len(oo))) if go[i] != oo[i]), None)
                    print("   first differing order", k, "gpu", go[k] if k is not None else None, "oracle", oo[k] if k is not None else None, "n", len(go), len(oo))
                    print("   hist gpu", hg[step][:9], "oracle", ho[step][:9])
    print(f"{name:28s} error bits 0x{int(np.bitwise_or.reduce(errs)):x} mismatching (market, asset, first step): {bad[:6]}")

# ---- event-level diff of one market at its first differing step (soak config)
e = core.BatchedEnv(n_m * A, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768, max_steps=64, max_queue=512, assets=A, **kw)
e.set_agents(build(core, gs), assets=assets); e.run_agents(n_steps, rs)
m = 2
o = orc.MarketEnv(0, 0, [1] * A, 1_000_000); o.set_groups(build(orc, gs), assets); o.run_agents(n_steps, rs, market_id=m, keyed=True)
def events(orders, a):
    ev = {}
    for od in orders:
        side, status, arr, end, vol, sv, price, trader, oid = od
        if status != 0: ev[arr] = ("arr", a, oid, trader)
        if end != 2**64 - 1 and status == 3: ev[end] = ("cancel", a, oid, trader)
    return ev
for step in range(n_steps):
    g_ev, o_ev = {}, {}
    for a in range(A):
        g_ev.update({t: v for t, v in events(e.get_orders(m * A + a), a).items() if step * 1_000_000 <= t < (step + 1) * 1_000_000})
        o_ev.update({t: v for t, v in events(o.get_orders(a), a).items() if step * 1_000_000 <= t < (step + 1) * 1_000_000})
    if g_ev != o_ev:
        print("market", m, "step", step, "events gpu", len(g_ev), "oracle", len(o_ev))
        ts = sorted(set(g_ev) | set(o_ev))
        shown = 0
        for t in ts:
            if g_ev.get(t) != o_ev.get(t) and shown < 14:
                print("   t", t % 1_000_000, "gpu", g_ev.get(t), "oracle", o_ev.get(t)); shown += 1
        gs_, os__ = set(g_ev.values()), set(o_ev.values())
        print("   only on gpu:", sorted(gs_ - os__)[:10])
        print("   only on oracle:", sorted(os__ - gs_)[:10])
        for a in range(A):
            go, oo = e.get_orders(m * A + a), o.get_orders(a)
            for (kind, aa, oid, tr) in sorted(os__ - gs_) + sorted(gs_ - os__):
                if aa == a:
                    print("   order", oid, "gpu", go[oid], "oracle", oo[oid])
        break
