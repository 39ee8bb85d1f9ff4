"""One batched policy decision step at config B/C size (4096 economies), for kernel-level profiling."""
import sys
import torch
sys.path.insert(0, ".")
from fastace_b200 import _abi, policy, scenario
from fastace_b200.env import BatchedEconomy

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dims = (E, 100, 10, 2, 10)
env = BatchedEconomy(dims)
env.set_state(scenario.custom_initial_state(dims, 3)[0])
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
nets = policy.DecisionNets(numGoods=2).cuda().eval()
pol = policy.BatchedPolicy(env, nets, fused=(len(sys.argv) > 3 and sys.argv[3] == 'fused'))
orders = scenario.OrderStream(dims, 4)
out = env.alloc_outputs()
perms = [tuple(torch.from_numpy(a).cuda() for a in orders.next()) for _ in range(reps + 2)]
for k in range(2):
    pol.step(perms[k], out, flags=_abi.IDX_ABSOLUTE)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(reps):
    pol.step(perms[2 + k], out, flags=_abi.IDX_ABSOLUTE)
e1.record(); torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / reps)
