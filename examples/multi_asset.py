"""crates/step_sim/examples/multi_asset/main.rs on the GPU: MarketEnv::<2> with four RandomMarketAgents groups, run by
market_sim_runner — here for 2048 markets in lockstep.

    python examples/multi_asset.py [n_markets]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bourse_b200.market import MarketEnv, RandomMarketAgents, market_sim_runner  # noqa: E402

n_markets = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
env = MarketEnv(0, 0, [1, 1], 1_000_000, True, n_markets=n_markets, price_window=(20, 180), max_queue=80)
agents = [RandomMarketAgents(0, 50, (40, 60), (10, 20), 2, 0.8), RandomMarketAgents(0, 50, (10, 90), (50, 70), 2, 0.2),
          RandomMarketAgents(1, 50, (40, 60), (10, 20), 2, 0.8), RandomMarketAgents(1, 50, (10, 90), (50, 70), 2, 0.2)]
market_sim_runner(env, agents, 101, 100, True)
print(f"{len(env.get_trades(0))} trades of asset 0")
print(f"{len(env.get_trades(1))} trades of asset 1")
print("all markets:", env.stats())
