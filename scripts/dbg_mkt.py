"""Debug: price range of the mixed market-agent population on the oracle, and dense-engine errors on the GPU."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as orc
from tests import test_gpu_market_agents as T
from bourse_b200 import market

n_assets = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
agents = T.mixed_agents(n_assets, noise=False)
for mk in range(5):
    o = T._oracle_market(orc, agents, n_assets, 100_000, [30], 2024, mk)
    for a in range(n_assets):
        pr = np.array([x[6] for x in o.get_orders(a)])
        lim = pr[(pr != 0) & (pr != 2**32 - 1)]
        print("market", mk, "asset", a, "orders", len(pr), "limit price range", lim.min(), lim.max(), "n>=1024:", (lim >= 1024).sum())
if "--gpu" in sys.argv:
    g = market.MarketEnv(0, 0, [1] * n_assets, 100_000, n_markets=5, max_orders=8192, max_trades=8192, max_queue=128,
                         price_window=(0, 1024), live_cap=254)
    g.set_agents(agents)
    for s in range(30):
        try:
            g.run_agents(1, 2024)
        except Exception as ex:
            print("step", s, ex)
        e = g.env_errors()
        if e.any():
            print("step", s, e)
            break
