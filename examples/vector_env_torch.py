"""A device-resident RL-style loop: a (random) policy written in torch trades against 100 background RandomAgents in each
of 4096 envs; actions, order ids and observations never leave the GPU.

    python examples/vector_env_torch.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from bourse_b200 import abi, core, gym  # noqa: E402

N_ENVS, ROWS = 4096, 4
env = gym.VectorEnv(N_ENVS, ROWS, seed=0, start_time=0, tick_size=1, step_size=1_000_000, max_queue=128, max_orders=16384,
                    max_trades=16384, max_steps=256,
                    agents=[core.random_group(50, (40, 60), (10, 20), 2, 0.8), core.random_group(50, (10, 90), (50, 70), 2, 0.2)],
                    agent_seed=101)
obs = torch.as_tensor(env.reset(), device="cuda")           # zero-copy view of the [N_ENVS, 45] observation buffer
ids = torch.as_tensor(env.ids, device="cuda")               # ... and of the [N_ENVS, ROWS] order-id buffer
actions = torch.zeros((N_ENVS, ROWS, 8), dtype=torch.int32, device="cuda")   # packed bb_instr rows: (t lo, t hi, op_flags, order_id, price, vol, trader, aux)
gen = torch.Generator(device="cuda").manual_seed(0)
for step in range(100):
    mid = ((obs[:, 1].long() + obs[:, 2].long()) // 2).clamp(40, 160)        # quote around the mid price (bid, ask = words 1, 2)
    side = torch.randint(0, 2, (N_ENVS, ROWS), device="cuda", generator=gen)
    off = torch.randint(0, 6, (N_ENVS, ROWS), device="cuda", generator=gen)
    price = ((mid[:, None] + torch.where(side == 1, -off, off)) // 2 * 2)    # on the agents' tick grid
    actions[:, :, 2] = (abi.OP_NEW | torch.where(side == 1, abi.F_BID, 0)).int()
    actions[:, :, 4] = price.int()
    actions[:, :, 5] = 5
    actions[:, :, 6] = 1000
    env.step(actions)                                                        # one launch: agents + these rows + Env::step
env.env.synchronize()
env.check_errors()
print("mean traded volume in the last step:", obs[:, 0].float().mean().item(), "| last ids of env 0:", ids[0].tolist())
print(env.env.stats())
