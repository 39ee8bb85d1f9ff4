"""One A2C update (rollout + re-evaluation backward + Adam) at 4096 economies x 20 steps: eager TF32 vs bf16 autocast."""
import sys
import torch
sys.path.insert(0, ".")
from fastace_b200 import _abi, policy, scenario, trainer
from fastace_b200.env import BatchedEconomy

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dims = (E, 100, 10, 2, 10)
env = BatchedEconomy(dims)
state = scenario.custom_initial_state(dims, 3)[0]
torch.backends.cuda.matmul.allow_tf32 = True
for label, dt, fused in (("tf32", None, False), ("bf16 autocast", torch.bfloat16, False), ("tf32 update, fused rollout", None, True)):
    torch.manual_seed(0)
    nets = policy.DecisionNets(numGoods=2).cuda()
    with torch.no_grad():
        for name, prm in nets.named_parameters():
            if ".last" in name and "offerEncoder" not in name and "jobOfferEncoder" not in name:
                prm.mul_(0.05)
    a2c = trainer.AdvantageActorCritic(nets, autocast_dtype=dt, adam_kwargs=dict(fused=True))
    pol = policy.BatchedPolicy(env, nets, autocast_dtype=dt, fused=fused)
    out = env.alloc_outputs()
    for it in range(2):
        env.set_state(state, time=0)
        orders = scenario.OrderStream(dims, 4)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        ep = trainer.run_episode(pol, orders, out, 20, flags=_abi.IDX_ABSOLUTE)
        ev[1].record()
        loss = a2c.train_on_episode(ep)
        ev[2].record(); torch.cuda.synchronize()
    print(f"{label}: rollout {ev[0].elapsed_time(ev[1]):.0f} ms, update {ev[1].elapsed_time(ev[2]):.0f} ms, loss {loss:.4g}, dropped {a2c.last_dropped}, peak {torch.cuda.max_memory_allocated()/2**30:.1f} GB")
