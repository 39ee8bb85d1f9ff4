"""`run` / `train` of the reference's Python front end (py/main.py:112-155 -> src/pybindings.cpp:78-114 ->
neural::train, src/neural/neuralScenarios.cpp:205-327) on the batched env — §8 row (f-3).

Same arguments as py/main.py's wrappers: the two ctypes structs (`CustomScenarioParams`, `TrainingParams`, byte-identical
layouts, created by the library's own create_scenario_params / create_training_params), `fromPretrained`,
`perturbationSize`; `train` returns the list of episode losses and writes the final learning rates back into
`trainingParams`, checkpoints the 11 nets every `checkpointEveryNEpisodes`, reloads the last checkpoint when an episode's
loss is NaN (or stops with "Training failed before first checkpoint"), prints the running average every
`updateEveryNEpisodes`.  What differs, by construction: every "episode" is `numEconomies` economies stepped at once on
the GPU (the reference steps one), so an episode loss is the mean over those economies.

Checkpoints: `<saveDir>/<net>.pt` for the reference's eleven names (decisionNetHandler.cpp:726-757).  save_models
writes torch state_dicts with the reference's parameter names; save_models_reference_format writes the TorchScript
archives the reference's torch::load reads; load_models reads either (the reference's own torch::save files included)."""
import math
import os

import numpy as np
import torch

from . import _abi, policy, scenario, trainer
from .env import BatchedEconomy

DEFAULT_SAVE_DIR = "../models/"      # neuralConstants.h:32


def save_models(nets, saveDir=DEFAULT_SAVE_DIR):
    os.makedirs(saveDir, exist_ok=True)
    for name in policy.NET_NAMES:
        torch.save({k: v.detach().cpu() for k, v in nets.net(name).state_dict().items()}, os.path.join(saveDir, name + ".pt"))


class _ParameterTree(torch.nn.Module):
    """parameter-only mirror of a module tree (same child and parameter names): what torch::load reads by name"""

    def __init__(self):
        super().__init__()


def _mirror(module):
    tree = _ParameterTree()
    for n, p in module.named_parameters(recurse=False):
        tree.register_parameter(n, torch.nn.Parameter(p.detach().cpu().clone(), requires_grad=False))
    for n, child in module.named_children():
        tree.add_module(n, _mirror(child))
    return tree


def save_models_reference_format(nets, saveDir=DEFAULT_SAVE_DIR):
    """The eleven files as the reference's own torch::save writes them (decisionNetHandler.cpp:726-738): a TorchScript
    archive per net whose attribute tree carries the parameters under the reference's names, so that the reference's
    DecisionNetHandler::load_models (torch::load, :745-757) reads them."""
    os.makedirs(saveDir, exist_ok=True)
    for name in policy.NET_NAMES:
        torch.jit.save(torch.jit.script(_mirror(nets.net(name))), os.path.join(saveDir, name + ".pt"))


def load_models(nets, saveDir=DEFAULT_SAVE_DIR):
    for name in policy.NET_NAMES:
        path = os.path.join(saveDir, name + ".pt")
        import zipfile
        with zipfile.ZipFile(path) as z:
            scripted = any(n.endswith("constants.pkl") or "/code/" in n for n in z.namelist())
        if scripted:      # the reference's torch::save (or save_models_reference_format): a TorchScript archive
            sd = {k: v.detach() for k, v in torch.jit.load(path, map_location="cpu").named_parameters()}
        else:
            sd = torch.load(path, map_location="cpu", weights_only=True)
        own = dict(nets.net(name).named_parameters())
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    own[k].copy_(v)


def perturb_models(nets, pct, generator=None):
    """DecisionNetHandler::perturb_models (decisionNetHandler.cpp:763-775): every Linear weight becomes
    sqrt(1-pct) * W + sqrt(xavier_var * pct) * N(0,1) (perturb_layer, decisionNets.cpp:14-39); a net perturbs its own
    layers, the two encoders are perturbed once each."""
    assert 0.0 <= pct <= 1.0
    seen = set()
    with torch.no_grad():
        for name in policy.NET_NAMES:
            for mod_name, mod in nets.net(name).named_modules():
                if not isinstance(mod, torch.nn.Linear) or id(mod) in seen:
                    continue
                if name not in ("offerEncoder", "jobOfferEncoder") and ("offerEncoder" in mod_name or "jobOfferEncoder" in mod_name):
                    continue
                seen.add(id(mod))
                out_f, in_f = mod.weight.shape
                std = math.sqrt(2.0 / (in_f + out_f) * pct)
                noise = torch.randn(mod.weight.shape, generator=generator, device="cpu").to(mod.weight.device) * std
                mod.weight.mul_(math.sqrt(1 - pct)).add_(noise)


def _build(scenarioParams, trainingParams, numEconomies, device):
    tp = trainingParams
    dims = (numEconomies, int(scenarioParams.numPeople), int(scenarioParams.numFirms), 2, int(tp.stackSize))
    env = BatchedEconomy(dims, device=device)
    nets = policy.DecisionNets(numGoods=2, stackSize=int(tp.stackSize), encodingSize=int(tp.encodingSize),
                               hiddenSize=int(tp.hiddenSize), nHidden=int(tp.nHidden), nHiddenSmall=int(tp.nHiddenSmall))
    nets = nets.to(torch.device("cuda", device))
    return dims, env, nets


def market_info(env, economy=0, goods=("good1", "good2")):
    """print_info of src/pybindings.cpp:20-75 for one economy: average price per good and average wage, from the
    device-side reduction (fastace_env_market_stats) — a few numbers cross the bus, not the books"""
    ms = env.market_stats()
    sums = ms["sum_quantity_per_price"][economy].tolist()
    counts = ms["offers"][economy].tolist()
    lines = [f"Time = {env.get_time()}:"]
    if sum(counts) > 0:
        for g, gname in enumerate(goods):
            if counts[g] > 0:
                lines.append(f"{gname}: Avg. price = {sums[g] / counts[g]:g} (num. offers = {counts[g]})")
            else:
                lines.append(f"{gname}: Avg. price = NA (num. offers = 0)")
    else:
        lines.append("[No offers]")
    nj = int(ms["job_offers"][economy])
    if nj > 0:
        lines.append(f"Avg. wage per unit of labor = {float(ms['sum_wage_per_labor'][economy]) / nj:g} (num. offers = {nj})")
    else:
        lines.append("[No job offers]")
    return "\n".join(lines)


def run(scenarioParams, trainingParams, numEconomies=1, device=0, saveDir=DEFAULT_SAVE_DIR, seed=0, fused=True, quiet=False):
    """lib.run: load the saved nets, step one episode without gradients, print the market after every step."""
    dims, env, nets = _build(scenarioParams, trainingParams, numEconomies, device)
    load_models(nets, saveDir)
    env.set_state(scenario.custom_initial_state(dims, seed, scenarioParams)[0])
    pol = policy.BatchedPolicy(env, nets.eval(), fused=fused, two_phase=True)
    orders = scenario.DeviceOrderStream(env, seed + 1, chunk=int(trainingParams.episodeLength))   # std::shuffle on the device
    out = env.alloc_outputs()
    log = []
    for _ in range(int(trainingParams.episodeLength)):
        pol.step(orders.next(), out, flags=_abi.IDX_ABSOLUTE)
        log.append(market_info(env))
        if not quiet:
            print(log[-1])
    env.close()
    return log


def train(scenarioParams, trainingParams, fromPretrained=False, perturbationSize=0.0, numEconomies=64, device=0,
          saveDir=DEFAULT_SAVE_DIR, seed=0, quiet=False, trainer_kwargs=None, fused_rollout=False):
    """lib.train: returns the episode losses; learning rates are written back into trainingParams.
    fused_rollout=True acts through the fused net kernel (bf16 operands; ~2x faster episodes) — the update always
    re-evaluates the recorded steps in fp32 autograd."""
    tp = trainingParams
    dims, env, nets = _build(scenarioParams, tp, numEconomies, device)
    if fromPretrained:                                   # train_from_pretrained, neuralScenarios.cpp:313-327
        load_models(nets, saveDir)
        if perturbationSize > 0.0:
            perturb_models(nets, perturbationSize)
    lrs = {n: float(getattr(tp, n + "LR")) for n in trainer.NET_ORDER}
    a2c = trainer.AdvantageActorCritic(
        nets, lrs=lrs, episodeBatchSizeForLRDecay=int(tp.episodeBatchSizeForLRDecay), patienceForLRDecay=int(tp.patienceForLRDecay),
        multiplierForLRDecay=float(tp.multiplierForLRDecay), cosinePeriod=int(tp.reverseAnnealingPeriod), **(trainer_kwargs or {}))
    # phase-wise: the consumption decision and the firms' decisions see what they see in the reference; a person's
    # purchase decision is still taken from its start-of-step money and labour (in the reference: after its own job
    # search, neuralPersonDecisionMaker.cpp:56-70) — DESIGN.md §7
    pol = policy.BatchedPolicy(env, nets, two_phase=True, fused=fused_rollout)
    out = env.alloc_outputs()
    say = (lambda *a: None) if quiet else print
    losses = [0.0] * int(tp.numEpisodes)
    for i in range(int(tp.numEpisodes)):
        state, discount = scenario.custom_initial_state(dims, seed + i * numEconomies, scenarioParams)   # scenario->setup()
        env.set_state(state, time=0)
        a2c.discount = torch.from_numpy(discount).to(torch.device("cuda", device))     # UtilMaxer::get_discountRate per person
        orders = scenario.DeviceOrderStream(env, seed + 7919 * (i + 1), chunk=int(tp.episodeLength))   # economy.cpp:110-111 on the device
        ep = trainer.run_episode(pol, orders, out, int(tp.episodeLength), flags=_abi.IDX_ABSOLUTE)
        loss = a2c.train_on_episode(ep)
        if math.isnan(loss):                              # neuralScenarios.cpp:229-243
            if i >= int(tp.checkpointEveryNEpisodes):
                say(f"In episode {i + 1}: NaN encountered; reverting to last checkpoint.")
                load_models(nets, saveDir)
                loss = losses[i - 1]
            else:
                say("Training failed before first checkpoint.")
                break
        elif (i + 1) % int(tp.checkpointEveryNEpisodes) == 0 or i == int(tp.numEpisodes) - 1:
            save_models(nets, saveDir)
        losses[i] = loss
        every = int(tp.updateEveryNEpisodes)
        if every != 0 and ((i + 1) % every == 0 or i + 1 == int(tp.numEpisodes)):
            window = every if every <= int(tp.numEpisodes) else int(tp.numEpisodes)
            avg = float(np.mean([losses[i - j] for j in range(window)]))
            say(f"Episode {i + 1}: Average loss over past {window} episodes = {avg:.6e}")
    for n, lr in a2c.learning_rates().items():            # neuralScenarios.cpp:264-272
        setattr(tp, n + "LR", lr)
    env.close()
    return losses


# ---- behind the C symbols `run` / `train` (csrc/legacy_capi.cpp): addresses of the caller's structs ---------------------
def _env_options():
    return dict(numEconomies=int(os.environ.get("FASTACE_NUM_ECONOMIES", "64")), device=int(os.environ.get("FASTACE_DEVICE", "0")),
                seed=int(os.environ.get("FASTACE_SEED", "0")))


def _c_train(out_addr, sp_addr, tp_addr, from_pretrained, perturbation):
    """train(double* output, const CustomScenarioParams*, TrainingParams*, bool, double) of src/pybindings.cpp:92-114"""
    import ctypes as C
    sp = _abi.CustomScenarioParams.from_address(sp_addr)
    tp = _abi.TrainingParams.from_address(tp_addr)          # learning rates are written back in place
    losses = train(sp, tp, fromPretrained=bool(from_pretrained), perturbationSize=float(perturbation), **_env_options())
    out = (C.c_double * int(tp.numEpisodes)).from_address(out_addr)
    for i, v in enumerate(losses):
        out[i] = float(v)
    return 0


def _c_run(sp_addr, tp_addr, _unused, _flag, _value):
    """run(CustomScenarioParams, TrainingParams) of src/pybindings.cpp:78-89 (the C side hands over its by-value copies)"""
    sp = _abi.CustomScenarioParams.from_address(sp_addr)
    tp = _abi.TrainingParams.from_address(tp_addr)
    run(sp, tp, **_env_options())
    return 0
