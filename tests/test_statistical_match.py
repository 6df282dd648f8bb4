"""Statistical match of agent-driven runs (BASELINE.json north_star: "statistically matched on agent-driven runs").

The reference threads ONE Xoroshiro128** stream through every agent and the shuffle (crates/step_sim/src/runner.rs:46-69);
the batched path keys Philox per (env, step, agent) instead, so individual runs differ draw for draw and only the
distributions can agree.  Each test runs the same population both ways over many independent envs and compares the
across-env means of per-env summary statistics with a two-sample z-test (tolerance: 5 standard errors, stated at the
assertion) and the DISTRIBUTIONS of per-step observables with two-sample Kolmogorov-Smirnov tests (one observation per env
and step, alpha = 0.001; `_assert_ks_match`).  Seeds are fixed, so the outcome is deterministic."""
import numpy as np
import pytest

from bourse_b200 import workloads

N_STEPS = 300


def _env_summaries(hist):
    """hist [n_envs, n_steps, >=9] (numpy level-1 layout) -> [n_envs, k] per-env summary statistics."""
    h = hist[:, 50:, :].astype(np.float64)   # skip the warm-up from the empty book
    both = (h[:, :, 1] > 0) & (h[:, :, 2] < 2**32 - 1)
    spread = np.where(both, h[:, :, 2] - h[:, :, 1], np.nan)
    mid = np.where(both, 0.5 * (h[:, :, 2] + h[:, :, 1]), np.nan)
    return np.stack([h[:, :, 0].mean(1),            # traded volume per step
                     np.nanmean(spread, 1), np.nanmean(mid, 1),
                     h[:, :, 3].mean(1), h[:, :, 4].mean(1),   # resting ask / bid volume
                     h[:, :, 5].mean(1), h[:, :, 7].mean(1),   # touch volumes
                     h[:, :, 6].mean(1), h[:, :, 8].mean(1)], axis=1)


NAMES = ["trade_vol/step", "spread", "mid", "ask_vol", "bid_vol", "bid_touch_vol", "ask_touch_vol", "bid_touch_n", "ask_touch_n"]


def _assert_match(a, b, n_sigma=5.0):
    ma, mb = a.mean(0), b.mean(0)
    se = np.sqrt(a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b))
    z = np.abs(ma - mb) / np.maximum(se, 1e-12)
    for name, x, y, zz in zip(NAMES, ma, mb, z):
        assert zz <= n_sigma, f"{name}: {x:.4f} vs {y:.4f} ({zz:.1f} standard errors)"   # tolerance: 5 standard errors
    # and the two really are different samples of a non-degenerate distribution
    assert not np.array_equal(a, b) and (a.std(0) > 0).all()


KS_STEPS = (120, 180, 240, 299)   # late, well separated steps: one observation per env and step, independent across envs
KS_FIELDS = {"trade_vol/step": 0, "bid_price": 1, "ask_price": 2, "ask_vol": 3, "bid_vol": 4, "bid_touch_vol": 5, "ask_touch_vol": 7}


def _assert_ks_match(a_hist, b_hist, alpha=1e-3):
    """Distributional match: two-sample Kolmogorov-Smirnov tests on the per-step traded volume,
    the touch prices, the spread, the side totals and the touch volumes.  Observations within one env are autocorrelated,
    so each test takes ONE observation per env (the value at a fixed step) — independent draws, as the test assumes.
    Tolerance: no test rejects at alpha = 0.001 (seeds are fixed: the outcome is deterministic)."""
    from scipy.stats import ks_2samp

    worst = (1.0, None)
    for step in KS_STEPS:
        cols = {name: (a_hist[:, step, w].astype(np.float64), b_hist[:, step, w].astype(np.float64)) for name, w in KS_FIELDS.items()}
        cols["spread"] = (a_hist[:, step, 2].astype(np.float64) - a_hist[:, step, 1], b_hist[:, step, 2].astype(np.float64) - b_hist[:, step, 1])
        for name, (x, y) in cols.items():
            res = ks_2samp(x, y)
            assert res.pvalue >= alpha, f"KS rejects {name} at step {step}: D = {res.statistic:.3f}, p = {res.pvalue:.2e}"
            worst = min(worst, (res.pvalue, f"{name}@{step}"))
    return worst


def _oracle_hist(oracle, groups, n_envs, keyed, seed):
    out = []
    for e in range(n_envs):
        env = oracle.StepEnvNumpy(0, 0, 1, 1_000_000)
        env.set_groups(groups)
        # stream mode: each env is its own reference-style sim_runner seeded seed + e; keyed mode: Philox (seed, env e)
        env.run_agents(N_STEPS, seed + (0 if keyed else e), env_id=e, keyed=keyed)
        out.append(env._history()[:, :9])
    return np.stack(out)


@pytest.mark.parametrize("groups_fn", [workloads.c3_groups, workloads.c4_groups], ids=["random", "random+momentum"])
def test_philox_contract_matches_reference_stream_statistically(oracle, groups_fn):
    groups = groups_fn()
    hk, hs = _oracle_hist(oracle, groups, 192, True, 11), _oracle_hist(oracle, groups, 192, False, 1234)
    _assert_match(_env_summaries(hk[:96]), _env_summaries(hs[:96]))
    _assert_ks_match(hk, hs)
    # the KS test has teeth: the same population with a different activity rate is told apart
    other = [g.copy() for g in groups]
    other[0]["rate"] = 0.5
    with pytest.raises(AssertionError, match="KS rejects"):
        _assert_ks_match(hk, _oracle_hist(oracle, other, 192, False, 99))


@pytest.mark.gpu
def test_gpu_agents_match_reference_stream_statistically(core, oracle):
    """The CUDA path (Philox contract, 1024 envs) against the reference-style single-stream runs of the oracle."""
    from bourse_b200 import abi
    groups = workloads.c3_groups()
    env = core.BatchedEnv(1024, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=32768, max_trades=32768, max_steps=N_STEPS,
                          max_queue=128, price_window=(20, 180), live_cap=128)
    env.set_agents(groups)
    env.run_agents(N_STEPS, 11)
    assert not env.env_errors().any()
    hist = env.history_all(N_STEPS)
    hs = _oracle_hist(oracle, groups, 256, False, 1234)
    _assert_match(_env_summaries(hist), _env_summaries(hs[:128]))
    _assert_ks_match(hist, hs)


def _oracle_market_hist(oracle, groups, assets, n_assets, n_markets, keyed, seed):
    out = []
    for m in range(n_markets):
        env = oracle.MarketEnv(0, 0, [1] * n_assets, 1_000_000)
        env.set_groups(groups, assets)
        # stream mode: market_sim_runner seeded seed + m (crates/step_sim/src/runner.rs:107-131); keyed: Philox (seed, market m)
        env.run_agents(N_STEPS, seed + (0 if keyed else m), market_id=m, keyed=keyed)
        out.append(np.stack([env.history(a)[:, :9] for a in range(n_assets)]))
    return np.stack(out)   # [markets, assets, steps, 9]


def test_market_twins_match_reference_stream_statistically(oracle):
    """The multi-asset example population (crates/step_sim/examples/multi_asset/main.rs): market-keyed Philox runs against
    reference-style market_sim_runner runs, per asset, across 64 markets."""
    groups, assets = workloads.market_example_groups()
    keyed = _oracle_market_hist(oracle, groups, assets, 2, 64, True, 21)
    stream = _oracle_market_hist(oracle, groups, assets, 2, 64, False, 4321)
    for a in range(2):
        _assert_match(_env_summaries(keyed[:, a]), _env_summaries(stream[:, a]))
    # the two assets of one market are separate books: their histories differ
    assert not np.array_equal(keyed[:, 0], keyed[:, 1])


@pytest.mark.gpu
def test_gpu_market_twins_match_reference_stream_statistically(oracle):
    """k_sim<.., MKT> (512 two-asset markets on the dense engine) against reference-style market_sim_runner runs."""
    from bourse_b200 import abi, core
    groups, assets = workloads.market_example_groups()
    env = core.BatchedEnv(1024, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1, max_orders=32768, max_trades=32768, max_steps=N_STEPS,
                          max_queue=96, price_window=(20, 180), live_cap=128, assets=2)
    env.set_agents(groups, assets=assets)
    env.run_agents(N_STEPS, 21)
    assert not env.env_errors().any()
    hist = env.history_all(N_STEPS).reshape(512, 2, N_STEPS, 9)
    stream = _oracle_market_hist(oracle, groups, assets, 2, 64, False, 4321)
    for a in range(2):
        _assert_match(_env_summaries(hist[:, a]), _env_summaries(stream[:, a]))
