#!/usr/bin/env python
"""Developer check: one C5-shaped book on the deep engine against the oracle; prints the first differences.
    python scripts/check_deep.py [n_rest] [n_steps] [per_step] [depth_ticks]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bourse_b200 import abi, core, workloads
from oracle import oracle as orc

if len(sys.argv) > 1 and sys.argv[1] == "c2":
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n = 20000
    s = workloads.replay_stream(n, seed, tick_size=1, trading_windows=False)
    ob = orc.OrderBook(0, 1)
    obs_cpu = ob.replay(s, obs_cap=n)
    env = core.BatchedEnv(1, 5, 0, 1, 1000, obs_words=abi.OBS_L2, max_orders=n + 64, max_trades=4 * n + 64, max_steps=n // 32 + 64, max_queue=32,
                          price_window=(896, 1152), deep_chunks=n // 8 + 2 * 256 + 64)
    try:
        env.replay(s)
    except Exception as e:
        print("replay raised", repr(e))
    go, co = env.orders_arrays(0), ob.orders_arrays()
    new_rows = np.nonzero((s["op_flags"] & 0xFF) == 1)[0]
    for k in co:
        d = np.nonzero(co[k] != go[k])[0]
        for i in d[:3]:
            print("orders", k, "id", i, "cpu", co[k][i], "gpu", go[k][i], "created at row", new_rows[i])
            rows = np.nonzero(((s["op_flags"] & 0xFF) != 1) & (s["order_id"] == i))[0]
            print("   rows naming it:", [(int(r), hex(int(s["op_flags"][r])), int(s["price"][r]), int(s["vol"][r])) for r in rows[:8]])
            print("   its row:", hex(int(s["op_flags"][new_rows[i]])), int(s["price"][new_rows[i]]), int(s["vol"][new_rows[i]]))
            print("   cpu", {kk: co[kk][i] for kk in co}, "\n   gpu", {kk: go[kk][i] for kk in go})
    gt, ct = env.trades_arrays(0), ob.trades_arrays()
    for k in ct:
        m = min(len(ct[k]), len(gt[k]))
        d = np.nonzero(ct[k][:m] != gt[k][:m])[0]
        print("trades", k, len(ct[k]), len(gt[k]), "first diff", (d[0], ct[k][d[0]], gt[k][d[0]]) if len(d) else None)
    print("hist equal", np.array_equal(env.history(0), obs_cpu))
    sys.exit(0)
n_rest = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
s = workloads.c5_stream(n_rest, n_steps, per_step, seed=100, depth_ticks=depth)
n = len(s)
ob = orc.OrderBook(0, 1)
obs_cpu = ob.replay(s, obs_cap=n_steps)
env = core.BatchedEnv(1, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=n + 64, max_trades=4 * n, max_steps=n_steps + 8, max_queue=32,
                      price_window=(10000 - depth - 64, 10000 + depth + 64), deep_chunks=int(os.environ.get("DBG_CHUNKS", n // 8 + 4 * depth + 4096)))
try:
    env.replay(s[:n_rest]); print("preload ok", env.stats())
    env.replay(s[n_rest:])
except Exception as e:
    print("replay raised", repr(e))
print("errors", [hex(int(x)) for x in env.env_errors()], env.stats())
go, co = env.orders_arrays(0), ob.orders_arrays()
for k in co:
    if len(co[k]) != len(go[k]):
        print("orders len", k, len(co[k]), len(go[k])); continue
    d = np.nonzero(co[k] != go[k])[0]
    if len(d):
        print("orders", k, "first diff at id", d[0], "cpu", co[k][d[0]], "gpu", go[k][d[0]], "n diffs", len(d), "(event idx of id:", np.nonzero((s["op_flags"] & 0xFF) == 1)[0][d[0]], ")")
gt, ct = env.trades_arrays(0), ob.trades_arrays()
for k in ct:
    m = min(len(ct[k]), len(gt[k]))
    d = np.nonzero(ct[k][:m] != gt[k][:m])[0]
    print("trades", k, len(ct[k]), len(gt[k]), "first diff", (d[0], ct[k][d[0]], gt[k][d[0]]) if len(d) else None)
print("l1", list(env.book_level_1(0)), ob._l1())
print("hist equal", np.array_equal(env.history(0), obs_cpu))
