"""Batched advantage actor-critic (fastace_b200/trainer.py) against the reference's own trainer.
tests/golden/a2c_episode.npz was produced by tests/golden/gen_a2c_golden.py from the UNMODIFIED
/root/reference/src/neural/{advantageActorCritic,decisionNetHandler,decisionNets}.cpp
(oracle/_ref/libfastace_refa2c.so).  Everything is fp32 autograd: tolerances are 2e-4 relative on
log-probabilities / losses 1e-4 of the largest entry for gradients, 2e-3 of the largest Adam update for parameters."""
import os
import sys

import numpy as np
import pytest
import torch

from fastace_b200 import policy, trainer

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "a2c_episode.npz")
sys.path.insert(0, os.path.join(HERE, "golden"))
NETS = trainer.NET_ORDER


def _sub(z, prefix):
    return {k[len(prefix):]: z[k] for k in z if k.startswith(prefix)}


def build(case, cfg):
    """nets + trainer initialised like the reference instance of this case"""
    nets = policy.DecisionNets(**cfg)
    nets.load_reference_parameters(_sub(case, "init/"))
    sched = case["sched"]
    a2c = trainer.AdvantageActorCritic(
        nets, lrs={n: float(case["lrs"][i]) for i, n in enumerate(NETS)},
        episodeBatchSizeForLRDecay=int(sched[0]), patienceForLRDecay=int(sched[1]), multiplierForLRDecay=float(sched[2]),
        cosinePeriod=int(sched[3]), discount=float(case["discount"]), adam_kwargs=dict(foreach=False))
    return nets, a2c


def snapshot_for(case, cfg, ep, t):
    """env-layout state tensors (E=1) of step t of episode ep, from the injected agent states"""
    G = cfg["numGoods"]
    P, F, nM, nJ = int(case["P"]), int(case["F"]), int(case["nM"]), int(case["nJ"])
    st = _sub(case, f"ep{ep}/state/")
    up, pp = case["util_params"], case["prod_params"]
    T = torch.from_numpy
    capM = F * G
    m_good = np.zeros((1, capM), np.uint8); m_price = np.zeros((1, capM)); j_wage = np.zeros((1, F))
    m_good[0, :nM] = case["good"]; m_price[0, :nM] = case["price"]; j_wage[0, :nJ] = case["wage"]
    return {
        "m_count": torch.tensor([nM], dtype=torch.int32), "m_good": T(m_good), "m_price": T(m_price),
        "j_count": torch.tensor([nJ], dtype=torch.int32), "j_wage": T(j_wage),
        "p_money": T(st["p_money"][t][None]), "p_labor_input": T(st["p_labor"][t][None]),
        "p_inv": T(np.ascontiguousarray(st["p_inv"][t].T[None])),
        "p_util_tfp": T(up[:, 0][None].copy()), "p_util_share": T(np.ascontiguousarray(up[:, 1:G + 2].T[None])),
        "p_util_rho": T(up[:, G + 2][None].copy()),
        "f_money": T(st["f_money"][t][None]), "f_labor": T(st["f_labor"][t][None]),
        "f_inv": T(np.ascontiguousarray(st["f_inv"][t].T[None])),
        "f_prod_tfp": T(np.ascontiguousarray(pp[:, :, 0].T[None])),                         # [1][G][F]
        "f_prod_share": T(np.ascontiguousarray(pp[:, :, 1:G + 2].transpose(1, 2, 0)[None])),   # [1][G][G+1][F]
        "f_prod_rho": T(np.ascontiguousarray(pp[:, :, G + 2].T[None])),
    }


def episode_for(case, cfg, nets, ep):
    T_, P, F = int(case["T"]), int(case["P"]), int(case["F"])
    st = _sub(case, f"ep{ep}/state/")
    epi = trainer.Episode()
    infos = []
    for t in range(T_):
        snap = snapshot_for(case, cfg, ep, t)
        draws = {k: torch.from_numpy(v) for k, v in _sub(case, f"ep{ep}/draws/{t}/").items()}
        with torch.no_grad():
            _, info = policy.evaluate(nets, snap, draws)
        infos.append(info)
        # env convention: f_profit[t] is the reward of step t-1 (the harness records it with offset 1 at step t)
        epi.append(snap, draws, info, torch.from_numpy(st["p_reward"][t][None]), torch.from_numpy(st["f_profit"][t][None]))
    return epi, infos


def check_case(case, cfg):
    nets, a2c = build(case, cfg)
    P, F, T_ = int(case["P"]), int(case["F"]), int(case["T"])
    for ep in range(int(case["episodes"])):
        epi, infos = episode_for(case, cfg, nets, ep)
        rec = case[f"ep{ep}/records"]    # [kind][t][agent]
        kinds = ("logp_purchase", "logp_firmPurchase", "logp_laborSearch", "logp_consumption", "logp_production",
                 "logp_offer", "logp_jobOffer")
        for t in range(T_):
            for k, name in enumerate(kinds):
                firm = name in ("logp_firmPurchase", "logp_production", "logp_offer", "logp_jobOffer")
                want = rec[k, t, P:] if firm else rec[k, t, :P]
                np.testing.assert_allclose(infos[t][name][0].numpy(), want, rtol=2e-4, atol=2e-5, equal_nan=True, err_msg=f"{name} t={t}")
            np.testing.assert_allclose(infos[t]["value_person"][0].numpy(), rec[7, t, :P], rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(infos[t]["value_firm"][0].numpy(), rec[7, t, P:], rtol=2e-4, atol=2e-5)
        loss = a2c.train_on_episode(epi)
        assert abs(loss - float(case[f"ep{ep}/loss"])) <= 2e-4 * abs(float(case[f"ep{ep}/loss"])) + 1e-4
        grads, params = _sub(case, f"ep{ep}/grad/"), _sub(case, f"ep{ep}/param/")
        for net in policy.NET_NAMES:
            for pname, p in nets.net(net).named_parameters():
                key = f"{net}/{pname}"
                g = p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float32)
                if net == "productionNet" and ep > 0:
                    continue    # never zeroed nor stepped in the reference wiring: its .grad accumulates over episodes
                scale = max(np.abs(grads[key]).max(), 1e-6)
                assert np.abs(g - grads[key]).max() <= 1e-4 * scale, (ep, key, np.abs(g - grads[key]).max(), scale)
                init = case["init/" + key]
                step = max(np.abs(params[key] - init).max(), 1e-7)
                assert np.abs(p.detach().numpy() - params[key]).max() <= 2e-3 * step + 1e-7, (ep, key)
        want_lrs = case[f"ep{ep}/lrs"]
        got = a2c.learning_rates()
        for i, n in enumerate(NETS):
            assert abs(got[n] - want_lrs[i]) <= 1e-12 + 1e-9 * want_lrs[i], (ep, n, got[n], want_lrs[i])
    return nets, a2c


@pytest.fixture(scope="module")
def gold():
    z = dict(np.load(GOLD))
    cfg = {k.split("/", 1)[1]: int(v[0]) for k, v in z.items() if k.startswith("cfg/")}
    return z, cfg


def test_reference_episodes_with_markets(gold):
    """3 episodes: recorded log-probs/values, losses, gradients, Adam updates, LR schedule of the reference"""
    z, cfg = gold
    nets, a2c = check_case(_sub(z, "market/"), cfg)
    # the reference wiring never steps productionNet (advantageActorCritic.cpp:111-116)
    init = _sub(_sub(z, "market/"), "init/")
    for pname, p in nets.productionNet.named_parameters():
        assert np.array_equal(p.detach().numpy(), init[f"productionNet/{pname}"])


def test_reference_episode_with_empty_markets(gold):
    """no offers: purchase log-probs are NaN ('no decision') and skipped, job search 0, encodings zero"""
    z, cfg = gold
    case = _sub(z, "empty/")
    assert np.isnan(case["ep0/records"][0, :, :int(case["P"])]).all()
    check_case(case, cfg)


def test_live_against_compiled_reference():
    gen = pytest.importorskip("gen_a2c_golden")
    if not os.path.exists(gen.LIB):
        pytest.skip("oracle/_ref/libfastace_refa2c.so not built here")
    L = gen.load()
    cfg = dict(gen.CFG, numGoods=3, stackSize=4)
    case = gen.run_reference(L, cfg, P=4, F=3, T=5, episodes=2, seed=23, sched=(1, 1, 0.5, 3))
    check_case(case, cfg)


def test_batch_of_economies_is_the_mean_of_single_economy_gradients(gold):
    """E economies at once == average of E reference-style single-economy updates (gradient level)"""
    z, cfg = gold
    case = _sub(z, "market/")
    nets, a2c = build(case, cfg)
    epi, _ = episode_for(case, cfg, nets, 0)
    epi2, _ = episode_for(case, cfg, nets, 1)
    # two different single-economy episodes
    grads = []
    for e in (epi, epi2):
        n1, a1 = build(case, cfg)
        a1.train_on_episode(e)
        grads.append({k: p.grad.clone() for k, p in n1.named_parameters() if p.grad is not None})
    # the same two stacked on the economy axis
    both = trainer.Episode()
    cat = lambda a, b: {k: torch.cat([a[k], b[k]], dim=0) for k in a}
    for t in range(len(epi)):
        both.steps.append((cat(epi.steps[t][0], epi2.steps[t][0]), cat(epi.steps[t][1], epi2.steps[t][1])))
        for name in ("value_person", "value_firm", "p_reward", "f_profit"):
            getattr(both, name).append(torch.cat([getattr(epi, name)[t], getattr(epi2, name)[t]], dim=0))
    n2, a2 = build(case, cfg)
    a2.train_on_episode(both)
    for k, p in n2.named_parameters():
        if p.grad is None:
            continue
        want = 0.5 * (grads[0][k] + grads[1][k])
        scale = max(want.abs().max().item(), 1e-6)
        assert (p.grad - want).abs().max().item() <= 1e-4 * scale, k


def test_lr_scheduler_rules():
    """advantageActorCritic.cpp:43-72 by hand: batches of 2, patience 2, x0.5 decay, kick every 3*2*2 = 12 updates"""
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    s = trainer.LRScheduler(opt, 2, 2, 0.5, 3)
    seq = [5, 5, 4, 4, 6, 6, 6, 6, 1, 1, 2, 2]     # batch sums 10, 8 (better), 12 (bad 1), 12 (bad 2 -> decay), 2, 4
    lrs = []
    for x in seq:
        s.update_lr(x); lrs.append(s.get_lr())
    assert lrs[:7] == [1.0] * 7 and lrs[7] == 0.5 and lrs[8:11] == [0.5] * 3
    assert lrs[11] == 1.0      # 12th update: reverse annealing 1/0.5
    assert s.cosineTimer == 0 and s.numBadBatches == 1 and s.bestBatchLoss == 2


# ---- N>1: economies sharded over ranks, one gradient all-reduce (gloo, world_size 2) -----------------
def _free_port():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _stack_episodes(a, b):
    both = trainer.Episode()
    cat = lambda x, y: {k: torch.cat([x[k], y[k]], dim=0) for k in x}
    for t in range(len(a)):
        both.steps.append((cat(a.steps[t][0], b.steps[t][0]), cat(a.steps[t][1], b.steps[t][1])))
        for name in ("value_person", "value_firm", "p_reward", "f_profit"):
            getattr(both, name).append(torch.cat([getattr(a, name)[t], getattr(b, name)[t]], dim=0))
    return both


def _rank_worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = dict(np.load(GOLD))
    cfg = {k.split("/", 1)[1]: int(v[0]) for k, v in z.items() if k.startswith("cfg/")}
    case = _sub(z, "market/")
    nets, a2c = build(case, cfg)
    losses = []
    for it in range(2):
        epi, _ = episode_for(case, cfg, nets, rank)     # rank r owns the economy of golden episode r
        losses.append(a2c.train_on_episode(epi))
    torch.save({"params": {k: p.detach().clone() for k, p in nets.named_parameters()}, "losses": losses,
                "lrs": a2c.learning_rates()}, os.path.join(tmpdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_equals_single_process(tmp_path, gold):
    import torch.multiprocessing as mp
    mp.spawn(_rank_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    z, cfg = gold
    case = _sub(z, "market/")
    nets, a2c = build(case, cfg)
    losses = []
    for it in range(2):
        e0, _ = episode_for(case, cfg, nets, 0)
        e1, _ = episode_for(case, cfg, nets, 1)
        losses.append(a2c.train_on_episode(_stack_episodes(e0, e1)))
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    for k, p in nets.named_parameters():
        assert torch.equal(r0["params"][k], r1["params"][k]), k               # replicas stay in lock-step
        init = torch.from_numpy(_sub(case, "init/")[k.replace(".", "/", 1)])
        step = (p.detach() - init).abs().max().item()
        assert (r0["params"][k] - p.detach()).abs().max().item() <= 2e-3 * step + 1e-7, k
    assert r0["lrs"] == r1["lrs"] == a2c.learning_rates()
    for a, b in zip(r0["losses"], losses):
        assert abs(a - b) <= 1e-4 * abs(b) + 1e-4


@pytest.mark.gpu
def test_train_on_env_episodes_gpu():
    """run_episode on the CUDA env + one update: finite loss, parameters move, re-evaluation reproduces the
    acting-time values and log-probabilities (same snapshot, same draws)"""
    from fastace_b200 import _abi, scenario
    from fastace_b200.env import BatchedEconomy
    dims = (64, 40, 6, 2, 8)
    env = BatchedEconomy(dims)
    env.set_state(scenario.custom_initial_state(dims, 3)[0])
    torch.manual_seed(0)
    nets = policy.DecisionNets(numGoods=2, stackSize=8, hiddenSize=32, nHidden=4, nHiddenSmall=2).cuda()
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    pol = policy.BatchedPolicy(env, nets, generator=gen)
    a2c = trainer.AdvantageActorCritic(nets, lr=1e-4, wiring="intended")
    out = env.alloc_outputs()
    orders = scenario.OrderStream(dims, 4)
    before = {k: p.detach().clone() for k, p in nets.named_parameters()}
    ep = trainer.run_episode(pol, orders, out, 6, flags=_abi.IDX_ABSOLUTE)
    assert len(ep) == 6 and env.get_time() == 6
    with torch.no_grad():
        _, info = policy.evaluate(nets, *ep.steps[3])
    torch.testing.assert_close(info["value_person"], ep.value_person[3], rtol=1e-4, atol=1e-5)
    assert int(ep.steps[3][0]["m_count"].sum()) > 0       # markets are populated by step 3
    loss = a2c.train_on_episode(ep)
    assert np.isfinite(loss)
    moved = sum(int(not torch.equal(before[k], p.detach())) for k, p in nets.named_parameters())
    assert moved == len(before)
    env.close()


def test_non_finite_economies_are_dropped(gold):
    """an economy whose episode recorded NaN (e.g. a NaN wage poisoned its money) is left out of the update:
    the result equals the single-economy update of the healthy one"""
    z, cfg = gold
    case = _sub(z, "market/")
    nets, a2c = build(case, cfg)
    good_ep, _ = episode_for(case, cfg, nets, 0)
    bad_ep, _ = episode_for(case, cfg, nets, 1)
    bad_ep.steps[2][0]["f_money"][0, 1] = float("nan")
    both = trainer.Episode()
    for t in range(len(good_ep)):
        snap = {k: torch.cat([good_ep.steps[t][0][k], bad_ep.steps[t][0][k]]) for k in good_ep.steps[t][0]}
        draws = {k: torch.cat([good_ep.steps[t][1][k], bad_ep.steps[t][1][k]]) for k in good_ep.steps[t][1]}
        with torch.no_grad():
            _, info = policy.evaluate(nets, snap, draws)
        both.append(snap, draws, info, torch.cat([good_ep.p_reward[t], bad_ep.p_reward[t]]),
                    torch.cat([good_ep.f_profit[t], bad_ep.f_profit[t]]))
    assert both.finite.tolist() == [True, False]
    loss = a2c.train_on_episode(both)
    assert a2c.last_dropped == 1 and np.isfinite(loss)
    n1, a1 = build(case, cfg)
    loss1 = a1.train_on_episode(good_ep)
    assert abs(loss - loss1) <= 1e-5 * abs(loss1)
    init = _sub(case, "init/")
    for (k, p), (_, q) in zip(nets.named_parameters(), n1.named_parameters()):
        step = (q.detach() - torch.from_numpy(init[k.replace(".", "/", 1)])).abs().max().item()
        assert (p.detach() - q.detach()).abs().max().item() <= 2e-3 * step + 1e-7, k   # same bar as the Adam-step checks
    # propagate = the reference's behaviour: NaN loss
    n2, a2 = build(case, cfg)
    a2.nan_policy = "propagate"
    assert np.isnan(a2.train_on_episode(both))


@pytest.mark.gpu
def test_value_nets_learn_on_env_episodes():
    """beyond gradient parity: a few updates on fresh episodes of the CUDA env reduce the critics' error
    (mean squared advantage of persons and firms) — the trainer, the recorded episodes and the env fit together"""
    from fastace_b200 import _abi, scenario
    from fastace_b200.env import BatchedEconomy
    dims = (128, 40, 6, 2, 8)
    env = BatchedEconomy(dims)
    state = scenario.custom_initial_state(dims, 3)[0]
    torch.manual_seed(0)
    nets = policy.DecisionNets(numGoods=2, stackSize=8, hiddenSize=32, nHidden=3, nHiddenSmall=2).cuda()
    with torch.no_grad():                      # finite log-normal heads at random init (see bench.py --train)
        for name, prm in nets.named_parameters():
            if ".last" in name and "offerEncoder" not in name and "jobOfferEncoder" not in name:
                prm.mul_(0.05)
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    pol = policy.BatchedPolicy(env, nets, generator=gen)
    # critics only (the policy heads keep lr 0, so the returns they are regressed on stay comparable between episodes)
    a2c = trainer.AdvantageActorCritic(nets, lr=0.0, lrs={"valueNet": 1e-2, "firmValueNet": 1e-2}, wiring="intended")
    out = env.alloc_outputs()
    errs = []
    for k in range(10):
        env.set_state(state, time=0)
        ep = trainer.run_episode(pol, scenario.OrderStream(dims, 4), out, 6, flags=_abi.IDX_ABSOLUTE)
        _, adv_p, _, adv_f = a2c.returns_and_advantages(ep)
        keep = ep.finite
        err = torch.stack([a[keep] for a in adv_p]).pow(2).mean().item() + torch.stack([a[keep] for a in adv_f[:-1]]).pow(2).mean().item()
        errs.append(err)
        assert np.isfinite(a2c.train_on_episode(ep))
    assert errs[-1] < 0.9 * errs[0] and min(errs[5:]) < min(errs[:3]), errs
    env.close()


@pytest.mark.gpu
def test_cuda_training_gradients_match_the_reference(gold):
    """the reference-pinned episode trained on the GPU: loss and every gradient against the reference's own
    (the CPU tests pin the same numbers; this covers the CUDA autograd path the trainer really runs on)"""
    z, cfg = gold
    case = _sub(z, "market/")
    nets, _ = build(case, cfg)
    epi, _ = episode_for(case, cfg, nets, 0)
    nets.cuda()
    a2c = trainer.AdvantageActorCritic(nets, lr=1e-4, discount=float(case["discount"]))
    dev = lambda d: {k: v.cuda() for k, v in d.items()}
    epi.steps = [(dev(a), dev(b)) for a, b in epi.steps]
    for name in ("value_person", "value_firm", "p_reward", "f_profit"):
        setattr(epi, name, [x.cuda() for x in getattr(epi, name)])
    epi.finite = epi.finite.cuda()
    loss = a2c.train_on_episode(epi)
    assert abs(loss - float(case["ep0/loss"])) <= 2e-4 * abs(float(case["ep0/loss"])) + 1e-4
    grads = _sub(case, "ep0/grad/")
    checked = 0
    for net in policy.NET_NAMES:
        for pname, p in nets.net(net).named_parameters():
            if p.grad is None:
                continue
            key = f"{net}/{pname}"
            scale = max(np.abs(grads[key]).max(), 1e-6)
            assert np.abs(p.grad.cpu().numpy() - grads[key]).max() <= 2e-4 * scale, key
            checked += 1
    assert checked > 100


@pytest.mark.gpu
@pytest.mark.parametrize("rows,K,H,residual", [(20000, 100, 100, True), (20000, 19, 100, False), (8192, 32, 32, True), (37, 9, 16, False)])
def test_training_layer_matches_eager_autograd(rows, K, H, residual):
    """train_layers.TanhLayer (csrc/layer_kernels.cuh + split-K weight gradient) against eager autograd, fp32"""
    from fastace_b200 import train_layers
    torch.manual_seed(rows + K)
    lin = torch.nn.Linear(K, H).cuda()
    x = torch.randn(rows, K, device="cuda", requires_grad=True)
    gy = torch.randn(rows, H, device="cuda")
    y0 = (x + torch.tanh(lin(x))) if residual else torch.tanh(lin(x))
    y0.backward(gy)
    ref = (x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None; lin.weight.grad = None; lin.bias.grad = None
    y1 = train_layers.tanh_layer(x, lin, residual)
    y1.backward(gy)
    torch.testing.assert_close(y1.detach(), y0.detach(), rtol=1e-6, atol=1e-6)
    for got, want in zip((x.grad, lin.weight.grad, lin.bias.grad), ref):
        assert (got - want).abs().max().item() <= 2e-5 * want.abs().max().item()
