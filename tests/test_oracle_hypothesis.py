"""CPU suite: property-based cross-check of the C++ oracle against the independent pure-Python restatement
(oracle/pybook.py) on adversarial short streams — tiny price / id / time domains so that equal-(price, time) key
collisions, zero volumes (N5), market sentinels (N3), trading toggles (N6), dead-order cancels and
modifies of every kind (N4) all occur within a few dozen instructions — plus the invariants of SURVEY.md 8c(ii)."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from bourse_b200 import abi

from .test_oracle_crosscheck import run_pybook

MAXP = 2**32 - 1

instr = st.one_of(
    st.tuples(st.just("new"), st.booleans(), st.integers(0, 6), st.sampled_from([0, 2, 4, 6, 8, MAXP - 1, MAXP]), st.booleans()),
    st.tuples(st.just("cancel"), st.integers(0, 40)),
    st.tuples(st.just("modify"), st.integers(0, 40), st.one_of(st.none(), st.sampled_from([0, 2, 3, 4, 6, MAXP])),
              st.one_of(st.none(), st.integers(0, 8))),
    st.tuples(st.just("trading"), st.booleans()),
)


def build(instrs, dts, tick):
    out = np.zeros(len(instrs), dtype=abi.INSTR_DTYPE)
    t, n_issued = 5, 0
    for i, (x, dt) in enumerate(zip(instrs, dts)):
        t = max(0, t + dt)
        of, oid, price, vol = abi.OP_NOOP, 0, 0, 0
        if x[0] == "new":
            _, bid, vol, price, market = x
            if not market and price % tick:
                price -= price % tick
            of = abi.OP_NEW | (abi.F_BID if bid else 0) | (abi.F_MARKET if market else 0)
            n_issued += 1
        elif x[0] == "cancel" and n_issued:
            of, oid = abi.OP_CANCEL, x[1] % n_issued
        elif x[0] == "modify" and n_issued:
            _, oid, p, v = x
            oid %= n_issued
            of = abi.OP_MODIFY | (abi.F_HAS_PRICE if p is not None else 0) | (abi.F_HAS_VOL if v is not None else 0)
            price, vol = p or 0, v or 0
        elif x[0] == "trading":
            of, vol = abi.OP_SET_TRADING, int(x[1])
        out[i] = (t, of | abi.F_EMIT, oid, price, vol, i % 7, 0)
    return out


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(instr, min_size=1, max_size=60), st.data(), st.sampled_from([1, 2]))
def test_oracle_equals_python_restatement_on_adversarial_streams(oracle, instrs, data, tick):
    dts = data.draw(st.lists(st.sampled_from([0, 0, 1, 1, 3, -1]), min_size=len(instrs), max_size=len(instrs)))
    s = build(instrs, dts, tick)
    ob = oracle.OrderBook(0, tick)
    obs = ob.replay(s, obs_cap=len(s))
    pb, pobs = run_pybook(s, tick)
    assert ob.get_trades() == pb.trades
    assert ob.get_orders() == pb.order_tuples()
    assert np.array_equal(obs, pobs)
    # invariants (8c ii): every trade is priced at the passive order's price; executed volume is conserved when no
    # modify changed a volume; statuses agree with end times
    orders = ob.get_orders()
    for t, pbid, price, vol, a, p in ob.get_trades():
        assert orders[p][0] == pbid and orders[a][0] != pbid
    if not any(x[0] == "modify" for x in instrs):
        executed = sum(o[5] - o[4] for o in orders)
        assert executed == 2 * sum(tr[3] for tr in ob.get_trades())
    for o in orders:
        assert (o[3] == 2**64 - 1) == (o[1] in (0, 1))


def random_adversarial_stream(rng, n, tick):
    """The same adversarial instruction space, drawn with numpy (for the batched GPU fuzz test)."""
    prices = [0, 2, 4, 6, 8, MAXP - 1, MAXP]
    instrs = []
    for _ in range(n):
        u = rng.random()
        if u < 0.5:
            instrs.append(("new", bool(rng.random() < 0.5), int(rng.integers(0, 7)), prices[int(rng.integers(len(prices)))],
                           bool(rng.random() < 0.15)))
        elif u < 0.7:
            instrs.append(("cancel", int(rng.integers(0, 41))))
        elif u < 0.95:
            p = None if rng.random() < 0.35 else [0, 2, 3, 4, 6, MAXP][int(rng.integers(6))]
            v = None if rng.random() < 0.35 else int(rng.integers(0, 9))
            instrs.append(("modify", int(rng.integers(0, 41)), p, v))
        else:
            instrs.append(("trading", bool(rng.random() < 0.6)))
    dts = [[0, 0, 1, 1, 3, -1][int(rng.integers(6))] for _ in range(n)]
    return build(instrs, dts, tick)
