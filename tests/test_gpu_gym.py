"""GPU suite for the device-resident vectorised loop (bourse_b200.gym.VectorEnv, bb_step_device / bb_level2_device):
per env it must be exactly StepEnvNumpy.submit_instructions + step + level_2_data of the reference
(rust/src/step_sim_numpy.rs:233-275), checked against one oracle StepEnvNumpy per env on the same rows."""
import numpy as np
import pytest

from bourse_b200 import abi, gym

pytestmark = pytest.mark.gpu


def _random_actions(rng, n_envs, rows, issued, tick):
    """Rows mixing limit / market orders (some off the tick grid), cancels and modifies of issued ids, and no-ops."""
    op = np.zeros((n_envs, rows), np.uint32)
    cols = {k: np.zeros((n_envs, rows), np.uint64) for k in ("bid", "vol", "trader", "price", "order_id", "market", "has_price", "has_vol")}
    for e in range(n_envs):
        for r in range(rows):
            u = rng.random()
            if u < 0.15:
                continue
            if u < 0.65 or issued[e] == 0:
                op[e, r] = abi.OP_NEW
                cols["bid"][e, r] = rng.random() < 0.5
                cols["vol"][e, r] = rng.integers(1, 50)
                cols["trader"][e, r] = rng.integers(0, 1000)
                cols["market"][e, r] = rng.random() < 0.07
                cols["price"][e, r] = tick * rng.integers(40, 61) + (1 if tick > 1 and rng.random() < 0.05 else 0)
            elif u < 0.85:
                op[e, r] = abi.OP_CANCEL
                cols["order_id"][e, r] = rng.integers(0, issued[e])
            else:
                op[e, r] = abi.OP_MODIFY
                cols["order_id"][e, r] = rng.integers(0, issued[e])
                hp, hv = rng.random() < 0.6, rng.random() < 0.6
                cols["has_price"][e, r], cols["has_vol"][e, r] = hp, hv
                cols["price"][e, r] = tick * rng.integers(40, 61)
                cols["vol"][e, r] = rng.integers(1, 50)
    return op, cols


def _oracle_step(o, op, cols, e, rows):
    ids = np.full(rows, abi.NO_ID, np.uint64)
    bad = False
    for r in range(rows):
        k = op[e, r]
        if k == abi.OP_NEW:
            price = None if cols["market"][e, r] else int(cols["price"][e, r])
            try:
                ids[r] = o.place_order(bool(cols["bid"][e, r]), int(cols["vol"][e, r]), int(cols["trader"][e, r]), price)
            except ValueError:
                bad = True
        elif k == abi.OP_CANCEL:
            o.cancel_order(int(cols["order_id"][e, r]))
        elif k == abi.OP_MODIFY:
            o.modify_order(int(cols["order_id"][e, r]), int(cols["price"][e, r]) if cols["has_price"][e, r] else None,
                           int(cols["vol"][e, r]) if cols["has_vol"][e, r] else None)
    o.step()
    return ids, bad


@pytest.mark.parametrize("kw,tick", [(dict(), 1), (dict(), 2), (dict(price_window=(32, 96), live_cap=254), 1)], ids=["paged_t1", "paged_t2", "dense"])
def test_vector_env_matches_one_step_env_per_env(oracle, kw, tick):
    n_envs, rows, n_steps, seed = 40, 7, 30, 11
    rng = np.random.default_rng(5)
    v = gym.VectorEnv(n_envs, rows, seed, 0, tick, 1000, max_orders=1024, max_trades=4096, max_steps=64, **kw)
    orcs = [oracle.StepEnv(seed + e, 0, tick, 1000) for e in range(n_envs)]
    issued = np.zeros(n_envs, np.int64)
    obs0 = v.reset().numpy()
    assert np.array_equal(obs0, np.stack([o.level_2_data_array() for o in orcs]))
    flagged = np.zeros(n_envs, bool)
    for _ in range(n_steps):
        op, cols = _random_actions(rng, n_envs, rows, issued, tick)
        obs, ids = v.step(gym.pack_actions(op, **cols))
        obs, ids = obs.numpy(), ids.numpy()
        for e in range(n_envs):
            want, bad = _oracle_step(orcs[e], op, cols, e, rows)
            flagged[e] |= bad
            assert np.array_equal(ids[e], want), e
            assert np.array_equal(obs[e], orcs[e].level_2_data_array()), e
            issued[e] = len(orcs[e].get_orders())
    err = v.env.env_errors()
    assert np.array_equal((err & 0x200) != 0, flagged) and not (err & ~np.uint32(0x200)).any()
    assert tick == 1 or flagged.any()
    for e in range(0, n_envs, 7):
        assert v.env.get_trades(e) == orcs[e].get_trades() and v.env.get_orders(e) == orcs[e].get_orders()
    if flagged.any():
        with pytest.raises(RuntimeError):
            v.check_errors()
    v.close()


def test_vector_env_zero_copy_with_torch(oracle):
    """Actions produced on the device (a torch tensor) go in without a host round trip; observations and ids are read
    through the CUDA array interface as torch views."""
    import torch

    n_envs, rows = 256, 4
    v = gym.VectorEnv(n_envs, rows, 3, 0, 1, 1000, level_1=True, max_orders=512, max_trades=512, max_steps=16)
    obs_t = torch.as_tensor(v.obs, device="cuda")
    ids_t = torch.as_tensor(v.ids, device="cuda")
    assert obs_t.shape == (n_envs, 9) and obs_t.data_ptr() == v.obs.ptr
    v.reset()
    # every env: bid 10 @ 50, ask 10 @ 52, ask 5 @ 50 (crosses half of the bid), no-op
    a = gym.pack_actions(np.tile(np.array([abi.OP_NEW, abi.OP_NEW, abi.OP_NEW, abi.OP_NOOP], np.uint32), (n_envs, 1)),
                         bid=np.tile([1, 0, 0, 0], (n_envs, 1)), vol=np.tile([10, 10, 5, 0], (n_envs, 1)),
                         price=np.tile([50, 52, 50, 0], (n_envs, 1)), trader=7)
    dev_actions = torch.from_numpy(a.view(np.uint8).reshape(n_envs, rows, 32)).cuda()
    torch.cuda.synchronize()
    v.step(dev_actions)
    v.env.synchronize()
    ids = ids_t.cpu().numpy()
    assert (ids[:, :3] == np.arange(3, dtype=np.uint64)).all() and (ids[:, 3] == abi.NO_ID).all()
    obs = obs_t.cpu().numpy().astype(np.uint32)
    for e in (0, 100, 255):
        o = oracle.StepEnvNumpy(3 + e, 0, 1, 1000)
        o.submit_instructions((np.array([1, 1, 1, 0], np.uint32), np.array([True, False, False, False]), np.array([10, 10, 5, 0], np.uint32),
                               np.full(4, 7, np.uint32), np.array([50, 52, 50, 0], np.uint32), np.zeros(4, np.uint64)))
        o.step()
        assert np.array_equal(obs[e], o.level_1_data()), e
    # depending on the shuffle the ask at 50 either traded with the bid or rests in front of it
    assert set(np.unique(obs[:, 0])) <= {0, 5}
    v.check_errors()
    v.close()


def test_dlpack_export_and_import(oracle):
    """DLPack at the Python edge (BASELINE north_star): observations / ids leave as DLPack capsules (`torch.from_dlpack`, zero
    copy), and an action block that speaks ONLY DLPack (no CUDA array interface) is accepted by `step`."""
    import torch

    n_envs, rows = 32, 3
    v = gym.VectorEnv(n_envs, rows, 11, 0, 1, 1000, max_orders=256, max_trades=256, max_steps=8)
    assert v.obs.__dlpack_device__() == (2, 0)            # kDLCUDA, device 0
    obs_t, ids_t = torch.from_dlpack(v.obs), torch.from_dlpack(v.ids)
    assert obs_t.data_ptr() == v.obs.ptr and tuple(obs_t.shape) == (n_envs, 45) and obs_t.dtype == torch.uint32
    assert ids_t.data_ptr() == v.ids.ptr and ids_t.dtype == torch.uint64
    v.reset()
    a = gym.pack_actions(np.tile(np.array([abi.OP_NEW, abi.OP_NEW, abi.OP_NOOP], np.uint32), (n_envs, 1)), bid=np.tile([1, 0, 0], (n_envs, 1)),
                         vol=np.tile([10, 4, 0], (n_envs, 1)), price=np.tile([50, 50, 0], (n_envs, 1)), trader=1)

    class DLPackOnly:   # a producer without __cuda_array_interface__
        def __init__(self, t): self.t = t
        def __dlpack__(self, stream=None): return self.t.__dlpack__()
        def __dlpack_device__(self): return self.t.__dlpack_device__()

    dev = torch.from_numpy(a.view(np.uint8).reshape(n_envs, rows, 32)).cuda()
    torch.cuda.synchronize()
    v.step(DLPackOnly(dev))
    v.env.synchronize()
    obs = obs_t.cpu().numpy().view(np.uint32)
    for e in (0, 31):
        o = oracle.StepEnv(11 + e, 0, 1, 1000)
        o.place_order(True, 10, 1, 50); o.place_order(False, 4, 1, 50); o.step()
        assert np.array_equal(obs[e], o.level_2_data_array()), e
    assert list(ids_t.cpu().numpy().view(np.uint64)[0]) == [0, 1, abi.NO_ID]
    with pytest.raises(ValueError):   # a host tensor is not a device array
        gym._dlpack_import(DLPackOnly(torch.zeros(4, dtype=torch.int32)))
    del obs_t, ids_t
    v.close()


def test_step_device_argument_checks():
    from bourse_b200 import core
    e = core.BatchedEnv(4, 0, 0, 1, 1000, assets=2, max_orders=64, max_trades=64, max_steps=8, max_queue=16)
    p = e.device_alloc(64)
    with pytest.raises(ValueError, match="single-asset"):
        e.step_device(p, p, 0)
    e.device_free(p)
    e2 = core.BatchedEnv(1, 0, 0, 1, 1000, max_orders=64, max_trades=64, max_steps=8, max_queue=16)
    e2.submit([abi.ACT_NEW], [1], [1], [0], [10])
    p = e2.device_alloc(64)
    with pytest.raises(ValueError, match="pending"):
        e2.step_device(p, p, 0)
    e2.device_free(p)


@pytest.mark.parametrize("kw,groups_name", [(dict(price_window=(20, 180), live_cap=254), "c3"), (dict(), "c3"), (dict(), "c4")],
                         ids=["dense_random", "fast_random", "fast_momentum"])
def test_vector_env_with_background_agents(oracle, kw, groups_name):
    """bb_run_agents_with_rows: every step is { built-in agents update; the action rows; Env::step } in one launch; per
    env it must equal the oracle's agents_update + place / cancel calls + keyed step (one shuffled queue)."""
    from bourse_b200 import workloads
    groups = workloads.c3_groups() if groups_name == "c3" else workloads.c4_groups()
    n_envs, rows, n_steps, seed = 24, 4, 25, 77
    rng = np.random.default_rng(8)
    v = gym.VectorEnv(n_envs, rows, 0, 0, 1, 1_000_000, agents=groups, agent_seed=seed, max_orders=8192, max_trades=8192,
                      max_steps=64, max_queue=160, **kw)
    orcs = [oracle.StepEnv(0, 0, 1, 1_000_000) for _ in range(n_envs)]
    for o in orcs:
        o.set_groups(groups)
    v.reset()
    mine = [[] for _ in range(n_envs)]   # ids this "learner" placed, per env
    for s in range(n_steps):
        op = np.zeros((n_envs, rows), np.uint32)
        cols = {k: np.zeros((n_envs, rows), np.uint64) for k in ("bid", "vol", "trader", "price", "order_id", "market")}
        for e in range(n_envs):
            for r in range(rows):
                u = rng.random()
                if u < 0.2:
                    continue
                if u < 0.75 or not mine[e]:
                    op[e, r] = abi.OP_NEW
                    cols["bid"][e, r], cols["vol"][e, r] = rng.random() < 0.5, rng.integers(1, 30)
                    cols["trader"][e, r], cols["market"][e, r] = 500 + r, rng.random() < 0.1
                    cols["price"][e, r] = 2 * rng.integers(40, 61)
                else:
                    op[e, r] = abi.OP_CANCEL
                    cols["order_id"][e, r] = mine[e][rng.integers(len(mine[e]))]
        obs, ids = v.step(gym.pack_actions(op, **cols))
        obs, ids = obs.numpy(), ids.numpy()
        for e, o in enumerate(orcs):
            o.agents_update(seed, e)
            want = np.full(rows, abi.NO_ID, np.uint64)
            for r in range(rows):
                if op[e, r] == abi.OP_NEW:
                    want[r] = o.place_order(bool(cols["bid"][e, r]), int(cols["vol"][e, r]), int(cols["trader"][e, r]),
                                            None if cols["market"][e, r] else int(cols["price"][e, r]))
                    mine[e].append(int(want[r]))
                elif op[e, r] == abi.OP_CANCEL:
                    o.cancel_order(int(cols["order_id"][e, r]))
            o.step_keyed(seed, e)
            assert np.array_equal(ids[e], want), (s, e)
            assert np.array_equal(obs[e], o.level_2_data_array()), (s, e)
    v.check_errors()
    for e in range(0, n_envs, 5):
        assert v.env.get_trades(e) == orcs[e].get_trades() and v.env.get_orders(e) == orcs[e].get_orders()
        assert np.array_equal(v.env.history(e), orcs[e]._history())
    # the learner's orders traded against the background population
    assert any(t[4] in mine[0] or t[5] in mine[0] for t in orcs[0].get_trades())
    # plain agent launches continue the same books (the on-chip agent tables are rebuilt per launch)
    v.env.run_agents(3, seed)
    for e in (0, 7):
        orcs[e].run_agents(3, seed, env_id=e, keyed=True)
        assert np.array_equal(v.env.history(e), orcs[e]._history())
    v.close()


def test_background_agent_rows_refuse_modify_and_unknown_ids():
    from bourse_b200 import workloads
    v = gym.VectorEnv(2, 2, 0, 0, 1, 1_000_000, agents=workloads.c3_groups(), agent_seed=1, max_orders=1024, max_trades=1024, max_steps=8,
                      max_queue=128)
    v.reset()
    v.step(gym.pack_actions(np.array([[abi.OP_MODIFY, abi.OP_NOOP], [abi.OP_NOOP, abi.OP_NOOP]], np.uint32), order_id=0, price=50, vol=1))
    err = v.env.env_errors()
    assert err[0] == 0x400 and err[1] == 0
    v.step(gym.pack_actions(np.array([[abi.OP_NOOP, abi.OP_NOOP], [abi.OP_CANCEL, abi.OP_NOOP]], np.uint32), order_id=100000))
    assert v.env.env_errors()[1] == 0x10
    v.close()


def test_step_device_inside_a_cuda_graph(oracle):
    """The vector step is one kernel launch on the caller's stream, so an RL loop can capture it (with its policy) into a
    CUDA graph: a captured step replayed K times equals K eager steps on a twin env fed the same action tensor."""
    import torch

    n_envs, rows, k = 64, 3, 5
    kw = dict(max_orders=512, max_trades=1024, max_steps=32)
    a, b = gym.VectorEnv(n_envs, rows, 2, 0, 1, 1000, **kw), gym.VectorEnv(n_envs, rows, 2, 0, 1, 1000, **kw)
    stream = torch.cuda.Stream()
    a.env.set_stream(stream.cuda_stream)
    a.reset(); b.reset()
    acts = torch.zeros((n_envs, rows, 8), dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(0)

    blocks = []   # every action block the envs executed, for the oracle

    def fill():
        blk = gym.pack_actions(np.full((n_envs, rows), abi.OP_NEW, np.uint32), bid=rng.random((n_envs, rows)) < 0.5,
                               vol=rng.integers(1, 20, (n_envs, rows)), price=rng.integers(45, 56, (n_envs, rows)), trader=3)
        blocks.append(blk)
        acts.copy_(torch.from_numpy(blk.view(np.int32).reshape(n_envs, rows, 8)))
        torch.cuda.synchronize()

    fill()
    with torch.cuda.stream(stream):
        a.step(acts)                       # warm-up outside the capture (attribute / occupancy queries happen here)
    stream.synchronize()
    b.step(acts); b.env.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        a.step(acts)
    for _ in range(k):                     # (the capture itself does not execute the step)
        fill()
        g.replay()
        torch.cuda.synchronize()
        b.step(acts); b.env.synchronize()
        assert np.array_equal(a.obs.numpy(), b.obs.numpy()) and np.array_equal(a.ids.numpy(), b.ids.numpy())
    assert a.env.n_steps(0) == b.env.n_steps(0) == k + 1
    for e in (0, 17, 63):
        assert a.env.get_trades(e) == b.env.get_trades(e) and np.array_equal(a.env.history(e), b.env.history(e))
    assert len(a.env.get_trades(0)) > 0
    # ... and both equal the oracle: one StepEnv per env fed the same blocks (the captured step was replayed, never re-recorded)
    assert len(blocks) == k + 1
    for e in (0, 17, 63):
        o = oracle.StepEnv(2 + e, 0, 1, 1000)
        for blk in blocks:
            for r in range(rows):
                x = blk[e, r]
                o.place_order(bool(int(x["op_flags"]) & abi.F_BID), int(x["vol"]), int(x["trader"]), int(x["price"]))
            o.step()
        assert np.array_equal(a.obs.numpy()[e], o.level_2_data_array()), e
        assert a.env.get_trades(e) == o.get_trades() and a.env.get_orders(e) == o.get_orders(), e
    # agent launches stage records without per-step capacity checks and refuse to be captured
    from bourse_b200 import workloads
    c = gym.VectorEnv(4, 2, 0, 0, 1, 1_000_000, agents=workloads.c3_groups(), agent_seed=1, max_orders=1024, max_trades=1024, max_steps=8, max_queue=128)
    c.env.set_stream(stream.cuda_stream)
    c.reset()
    small = torch.zeros((4, 2, 8), dtype=torch.int32, device="cuda")
    with torch.cuda.stream(stream):
        c.step(small)
    stream.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with pytest.raises(ValueError, match="cannot be captured"):
        with torch.cuda.graph(g2, stream=stream):
            c.step(small)
    a.close(); b.close()


@pytest.mark.parametrize("bg", [False, True], ids=["rows_only", "background_agents"])
def test_vector_env_runs_past_max_steps(oracle, bg):
    """An open-ended loop: the per-step history is a scratch ring for VectorEnv (bb_clear_history when it is full), so the
    loop runs past max_steps and keeps matching the oracle."""
    from bourse_b200 import workloads
    groups = workloads.c3_groups() if bg else None
    n_envs, rows, seed = 6, 2, 13
    v = gym.VectorEnv(n_envs, rows, 4, 0, 1, 1_000_000, agents=groups, agent_seed=seed, max_orders=8192, max_trades=8192, max_steps=8,
                      max_queue=128)
    orcs = [oracle.StepEnv(4 + e, 0, 1, 1_000_000) for e in range(n_envs)]
    if bg:
        for o in orcs:
            o.set_groups(groups)
    v.reset()
    rng = np.random.default_rng(1)
    for s in range(30):
        bid = rng.random((n_envs, rows)) < 0.5
        vol = rng.integers(1, 20, (n_envs, rows)); price = 2 * rng.integers(45, 56, (n_envs, rows))
        obs, _ = v.step(gym.pack_actions(np.full((n_envs, rows), abi.OP_NEW, np.uint32), bid=bid, vol=vol, price=price, trader=9))
        obs = obs.numpy()
        for e, o in enumerate(orcs):
            if bg:
                o.agents_update(seed, e)
            for r in range(rows):
                o.place_order(bool(bid[e, r]), int(vol[e, r]), 9, int(price[e, r]))
            o.step_keyed(seed, e) if bg else o.step()
            assert np.array_equal(obs[e], o.level_2_data_array()), (s, e)
    v.check_errors()
    assert v.env.get_trades(0) == orcs[0].get_trades() and v.env.n_steps(0) <= 8
    v.close()
