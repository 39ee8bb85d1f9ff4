"""A/B of the trainer's update (re-evaluation forward + backward + Adam) on one recorded batch of episodes:
eager autograd layers vs train_layers.py (fused element-wise halves, split-K weight gradients)."""
import sys
import torch
sys.path.insert(0, ".")
from fastace_b200 import _abi, policy, scenario, trainer
from fastace_b200.env import BatchedEconomy

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dims = (4096, 100, 10, 2, 10)
env = BatchedEconomy(dims)
state = scenario.custom_initial_state(dims, 3)[0]
torch.backends.cuda.matmul.allow_tf32 = (len(sys.argv) <= 2 or sys.argv[2] != 'fp32')
torch.manual_seed(0)
nets = policy.DecisionNets(numGoods=2).cuda()
with torch.no_grad():
    for name, prm in nets.named_parameters():
        if ".last" in name and "offerEncoder" not in name and "jobOfferEncoder" not in name:
            prm.mul_(0.05)
pol = policy.BatchedPolicy(env, nets, fused=True)
out = env.alloc_outputs()
env.set_state(state, time=0)
ep = trainer.run_episode(pol, scenario.OrderStream(dims, 4), out, T, flags=_abi.IDX_ABSOLUTE)
init = {k: v.detach().clone() for k, v in nets.state_dict().items()}
res = {}
for rep in range(3):
    for fl in (False, True):
        nets.load_state_dict(init)
        a2c = trainer.AdvantageActorCritic(nets, lr=0.0, adam_kwargs=dict(fused=True), fused_layers=fl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = a2c.train_on_episode(ep); e1.record(); torch.cuda.synchronize()
        res.setdefault(fl, []).append(e0.elapsed_time(e1))
        g = torch.cat([p.grad.flatten() for p in nets.parameters() if p.grad is not None])
        res.setdefault(("g", fl), g.clone())
print("eager   update ms:", [round(x) for x in res[False]])
print("layers  update ms:", [round(x) for x in res[True]])
ga, gb = res[("g", False)], res[("g", True)]
print("grad max rel diff:", float((ga - gb).abs().max() / ga.abs().max()), "loss", loss)
