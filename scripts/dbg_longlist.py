"""Does the MomentumAgent / NoiseAgent cancel sweep stay bit-exact once a live-order list is longer than one warp batch?
(A population without RandomAgents: the book is often one-sided, limit orders pile up far from the touch and never fill.)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from bourse_b200 import abi, core
from oracle import oracle as orc
M = (1000, 17, 1, 0.019482422074020622, 10, 0.41498651540274606, 6.351126071309, 0.7483027064602775, 1.820404354554679, 0.0, 0.9776575619920471)
N = (2000, 15, 1, 0.3854460237242081, 0.19789807296854475, 0.3151828300857516, 7, 0.6608400378801695, 1.0670506319030661)
for name, gs in (("m+n", lambda m: [m.momentum_group(*M), m.noise_group(*N)]), ("n+m", lambda m: [m.noise_group(*N), m.momentum_group(*M)]),
                 ("n only, p_cancel 0.02", lambda m: [m.noise_group(2000, 15, 1, 0.385, 0.198, 0.02, 7, 0.66, 1.07)])):
    e = core.BatchedEnv(8, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=16384, max_trades=32768, max_steps=64, max_queue=512,
                        pages_smem=64, pages_total=64)
    e.set_agents(gs(core)); e.run_agents(60, 5)
    bad = []
    for env in range(8):
        o = orc.StepEnvNumpy(0, 0, 1, 1_000_000); o.set_groups(gs(orc)); o.run_agents(60, 5, env_id=env, keyed=True)
        hg, ho = e.history(env), o._history()
        if not np.array_equal(hg, ho):
            bad.append((env, int(np.flatnonzero((hg != ho).any(axis=1))[0])))
        if env == 0:
            act = [sum(1 for od in o.get_orders() if lo <= od[7] < lo + 100 and od[1] == 1) for lo in (1000, 2000)]
    print(name, "errors", np.unique(e.env_errors()), "mismatch (env, first step):", bad, "| active momentum / noise orders at the end (env 0):", act)
