"""Times the large-economy path on BASELINE config D (1 x (100k persons + 5k firms), 8 goods, stack 10).
Device-resident actions, CUDA events around each step (the step synchronises internally once per round)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from fastace_b200 import _abi, scenario
from fastace_b200.env import BatchedEconomy

dims = (1, 100000, 5000, 8, 10)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
env = BatchedEconomy(dims)
env.set_state(scenario.generic_initial_state(dims, 21))
orders = scenario.OrderStream(dims, 38)
acts = [env.alloc_actions(scenario.synthetic_actions(dims, seed=22, step=t, perms=orders.next(), **scenario.BENCH_PRESET)) for t in range(steps)]
packed = [env.pack_device("actions", a) for a in acts]
out = env.alloc_outputs()
pout = env.pack_device("out", out)
ms, rounds = [], []
for t in range(steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.time_step(packed[t], pout, flags=_abi.IDX_MODULO); e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1)); rounds.append(env.large_stats())
print("ms per step:", [round(x, 3) for x in ms])
print("rounds:", rounds)
m = float(np.median(ms[2:]))
print(f"median {m:.3f} ms/step -> {105000 / (m * 1e-3):.3e} agent-steps/s")
