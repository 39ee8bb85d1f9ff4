"""Advantage actor-critic over E economies at once — §8 row (f-2).

The reference trains on ONE economy per episode, agent by agent: for every person and firm it walks the
episode backwards, builds the return q and the advantage q - V, and calls backward() on
sum_t (q_t - V_t)^2 + sum_t logp_t * advantage_t (src/neural/advantageActorCritic.cpp:244-455), then
steps nine Adam optimisers (:569-600).  Here the same loss is assembled for all agents of all economies
as tensors; the episode is stored as per-step (snapshot, draws) pairs and re-evaluated with autograd one
step at a time, so activation memory is one step deep whatever the episode length.  The gradient is the
MEAN over economies (all ranks) of the reference's per-economy gradient; with one economy it is the
reference's, bit for bit up to fp32 summation order (tests/test_trainer.py pins it against the compiled
reference trainer).

Reference behaviours kept on purpose (mode "reference", the default), each switchable:
  * loss = +logp*advantage (advantageActorCritic.cpp:259) — sign="reference"; "ascent" uses -logp*advantage.
  * the reparameterised sample stays in the graph (decisionNetHandler.cpp:31-32) — see policy.evaluate.
  * optimiser wiring (advantageActorCritic.cpp:93-122): the optimiser called laborSearchNetOptim holds
    firmPurchaseNet's parameters and LR, consumptionNetOptim holds laborSearchNet's, productionNetOptim
    holds consumptionNet's; productionNet is never stepped.  wiring="intended" gives each net its own.
  * valueNetLoss is never accumulated (only firmValueNetLoss is, :394), so the value net's scheduler sees 0.
Multi-GPU: economies are sharded over ranks; gradients and the tracked losses are summed with ONE
all-reduce over a flat bucket (NCCL on GPUs, gloo in the CPU tests) before the optimiser steps."""
import math

import torch
import torch.distributed as dist

from . import policy

NET_ORDER = ("purchaseNet", "firmPurchaseNet", "laborSearchNet", "consumptionNet", "productionNet",
             "offerNet", "jobOfferNet", "valueNet", "firmValueNet")

# optimiser name -> net whose parameters() and learning rate it is built on (advantageActorCritic.cpp:87-143)
REFERENCE_WIRING = {
    "purchaseNet": "purchaseNet", "firmPurchaseNet": "firmPurchaseNet", "laborSearchNet": "firmPurchaseNet",
    "consumptionNet": "laborSearchNet", "productionNet": "consumptionNet", "offerNet": "offerNet",
    "jobOfferNet": "jobOfferNet", "valueNet": "valueNet", "firmValueNet": "firmValueNet",
}
INTENDED_WIRING = {n: n for n in NET_ORDER}

DEFAULT_LEARNING_RATE = 1e-5                    # neuralConstants.h:25-29
DEFAULT_EPISODE_BATCH_SIZE_FOR_LR_DECAY = 10
DEFAULT_PATIENCE_FOR_LR_DECAY = 5
DEFAULT_MULTIPLIER_FOR_LR_DECAY = 0.5
DEFAULT_REVERSE_ANNEALING_PERIOD = 3


class LRScheduler:
    """advantageActorCritic.cpp:8-72: plateau decay on batches of episode losses plus a periodic 1/decay kick."""

    def __init__(self, optimizer, episodeBatchSize, patience, decayMultiplier, cosinePeriod, name=""):
        self.optimizer, self.name = optimizer, name
        self.episodeBatchSize, self.patience = episodeBatchSize, patience
        self.decayMultiplier, self.cosinePeriod = decayMultiplier, cosinePeriod
        self.lossHistory = []
        self.bestBatchLoss = math.inf          # HUGE_VALF
        self.numBadBatches = 0
        self.cosineTimer = 0

    def get_lr(self):
        return self.optimizer.param_groups[0]["lr"]

    def scale_lr(self, multiplier):
        for group in self.optimizer.param_groups:
            group["lr"] = group["lr"] * multiplier

    def update_lr(self, loss):
        self.lossHistory.append(loss)
        if len(self.lossHistory) == self.episodeBatchSize:
            recent = 0.0
            for x in self.lossHistory:
                recent += x
            if recent < self.bestBatchLoss:
                self.bestBatchLoss = recent
                self.numBadBatches = 0
            else:
                self.numBadBatches += 1
            if self.numBadBatches >= self.patience:
                self.scale_lr(self.decayMultiplier)
                self.numBadBatches = 0
            self.lossHistory.clear()
        self.cosineTimer += 1
        if self.cosineTimer == self.cosinePeriod * self.episodeBatchSize * self.patience:
            self.scale_lr(1.0 / self.decayMultiplier)
            self.cosineTimer = 0


class Episode:
    """One batch of E episodes: per-step (snapshot, draws) for re-evaluation, the values predicted while
    acting, and the env's rewards.  f_profit[t] is the env's profit output of step t, i.e. the firm's money
    at its turn in step t minus that at its turn in step t-1: the reward of the firm's decisions in step t-1
    (neuralFirmDecisionMaker.cpp:65-74 records it with offset 1)."""

    def __init__(self):
        self.steps, self.value_person, self.value_firm, self.p_reward, self.f_profit = [], [], [], [], []
        self.finite = None      # [E] bool: every recorded number of the economy's episode is finite

    def __len__(self):
        return len(self.steps)

    def append(self, snap, draws, info, p_reward, f_profit):
        self.steps.append((snap, draws))
        self.value_person.append(info["value_person"].detach())
        self.value_firm.append(info["value_firm"].detach())
        self.p_reward.append(p_reward.detach().clone())
        self.f_profit.append(f_profit.detach().clone())
        ok = torch.isfinite(p_reward).all(1) & torch.isfinite(f_profit).all(1)
        empty = (snap["m_count"] == 0).view(-1, 1)
        for key, value in info.items():
            fin = torch.isfinite(value)
            if key in ("logp_purchase", "logp_firmPurchase"):
                fin = fin | (torch.isnan(value) & empty)          # nan = no decision to make, by design
            ok = ok & fin.all(1)
        self.finite = ok if self.finite is None else (self.finite & ok)

    def select(self, keep):
        """the sub-batch of economies keep (index tensor)"""
        sub = Episode()
        pick = lambda d: {k: v.index_select(0, keep) for k, v in d.items()}
        sub.steps = [(pick(a), pick(b)) for a, b in self.steps]
        for name in ("value_person", "value_firm", "p_reward", "f_profit"):
            setattr(sub, name, [x.index_select(0, keep) for x in getattr(self, name)])
        sub.finite = self.finite.index_select(0, keep)
        return sub


def _masked(logp):
    """nan = 'no decision was made', skipped by the reference (advantageActorCritic.cpp:257-260)"""
    return torch.where(torch.isnan(logp), torch.zeros_like(logp), logp)


class AdvantageActorCritic:
    def __init__(self, nets, lr=DEFAULT_LEARNING_RATE, lrs=None,
                 episodeBatchSizeForLRDecay=DEFAULT_EPISODE_BATCH_SIZE_FOR_LR_DECAY,
                 patienceForLRDecay=DEFAULT_PATIENCE_FOR_LR_DECAY,
                 multiplierForLRDecay=DEFAULT_MULTIPLIER_FOR_LR_DECAY,
                 cosinePeriod=DEFAULT_REVERSE_ANNEALING_PERIOD,
                 discount=0.9, wiring="reference", sign="reference", sample_grad="reference",
                 autocast_dtype=None, process_group=None, adam_kwargs=None, nan_policy="drop_economy",
                 fused_layers=True):
        """lrs: optional {net name: lr} (TrainingParams.purchaseNetLR ..., neuralScenarios.h:173-181).
        nan_policy: the reference abandons an episode whose loss is NaN and reloads its last checkpoint
        (neuralScenarios.cpp:229-243).  With E episodes per update the equivalent is "drop_economy": economies
        whose episode recorded a non-finite value, log-probability or reward are left out of the update (the mean
        runs over the kept ones; `last_dropped` counts them); "propagate" keeps them and returns NaN like the
        reference."""
        self.nan_policy, self.last_dropped = nan_policy, 0
        # CUDA re-evaluation through train_layers.py (fused element-wise halves, split-K weight gradients); CPU / autocast: eager
        self.fused_layers = fused_layers
        self.nets, self.discount = nets, discount
        self.sign = 1.0 if sign == "reference" else -1.0
        self.sample_grad, self.autocast_dtype, self.group = sample_grad, autocast_dtype, process_group
        lrs = dict(lrs or {})
        table = REFERENCE_WIRING if wiring == "reference" else INTENDED_WIRING
        kw = dict(adam_kwargs or {})
        self.optims, self.schedulers = {}, {}
        for name in NET_ORDER:
            src = table[name]
            self.optims[name] = torch.optim.Adam(list(nets.net(src).parameters()), lr=lrs.get(src, lr), **kw)
            self.schedulers[name] = LRScheduler(self.optims[name], episodeBatchSizeForLRDecay, patienceForLRDecay,
                                                multiplierForLRDecay, cosinePeriod, name)
        self.tracked = {name: 0.0 for name in NET_ORDER}
        seen, self._params = set(), []
        for p in nets.parameters():
            if id(p) not in seen:
                seen.add(id(p)); self._params.append(p)

    # ---- pieces of train_on_episode (advantageActorCritic.cpp:602-624) ----
    def zero_all_grads(self):
        # every parameter, not only those an optimiser owns: with the reference's wiring productionNet belongs to no
        # optimiser (advantageActorCritic.cpp:103-120); its gradient would otherwise pile up from episode to episode
        # (and be multiplied by the world size at every all-reduce)
        for p in self._params:
            p.grad = None

    def update_lr_schedulers(self):
        for name in NET_ORDER:
            self.schedulers[name].update_lr(self.tracked[name])
        self.tracked = {name: 0.0 for name in NET_ORDER}

    def all_optims_step(self):
        for name in NET_ORDER:
            self.optims[name].step()

    def learning_rates(self):
        return {name: self.schedulers[name].get_lr() for name in NET_ORDER}

    def returns_and_advantages(self, ep, discount=None):
        """q and advantage series (get_advantage_for_person / _for_firm, :269-299, :345-375): fp64 returns,
        advantages stored in fp32 like the reference's torch::empty(time) buffer."""
        T = len(ep)
        q = torch.zeros_like(ep.p_reward[0], dtype=torch.float64)
        q_p, adv_p = [None] * T, [None] * T
        for t in range(T - 1, -1, -1):
            q = ep.p_reward[t].double() + (self.discount if discount is None else discount) * q
            q_p[t] = q
            adv_p[t] = (q - ep.value_person[t].double()).float()
        q = torch.zeros_like(ep.f_profit[0], dtype=torch.float64)
        q_f, adv_f = [None] * T, [None] * T
        for t in range(T - 2, -1, -1):          # the last step's payoff is never seen (:353-354)
            q = ep.f_profit[t + 1].double() + q  # no discounting with firms (:365)
            q_f[t] = q
            adv_f[t] = (q - ep.value_firm[t].double()).float()
        return q_p, adv_p, q_f, adv_f

    def train_on_episode(self, ep):
        """One update from a batch of episodes.  Returns the mean per-economy episode loss (what the
        reference's train_on_episode returns for its single economy)."""
        self.zero_all_grads()
        self.update_lr_schedulers()
        T = len(ep)
        dev = ep.p_reward[0].device
        E_local = ep.p_reward[0].shape[0]
        self.last_dropped = 0
        discount = self.discount        # a scalar, or one rate per person [E, P] (UtilMaxer::get_discountRate)
        if self.nan_policy == "drop_economy" and ep.finite is not None and not bool(ep.finite.all()):
            keep = ep.finite.nonzero().flatten()
            self.last_dropped = E_local - keep.numel()
            ep = ep.select(keep)
            E_local = keep.numel()
            if torch.is_tensor(discount) and discount.dim() == 2:
                discount = discount.index_select(0, keep)
        count = torch.tensor([float(E_local)], dtype=torch.float64, device=dev)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(count, group=self.group)
        if count.item() == 0:
            return float("nan")                 # nothing usable: no step (the reference would reload its checkpoint)
        scale = 1.0 / count.item()
        q_p, adv_p, q_f, adv_f = self.returns_and_advantages(ep, discount)
        names = ("purchaseNet", "laborSearchNet", "consumptionNet", "firmPurchaseNet", "productionNet", "offerNet",
                 "jobOfferNet", "firmValueNet")
        sums = torch.zeros(len(names) + 1, dtype=torch.float64, device=dev)   # tracked policy losses + total loss
        for t in range(T):
            snap, draws = ep.steps[t]
            with torch.enable_grad():
                policy._TRAIN_LAYERS = self.fused_layers
                try:
                    _, info = policy.evaluate(self.nets, snap, draws, self.autocast_dtype, self.sample_grad)
                finally:
                    policy._TRAIN_LAYERS = False
                err = q_p[t] - info["value_person"].double()
                terms = [self.sign * (_masked(info["logp_purchase"]) * adv_p[t]).sum(),
                         self.sign * (_masked(info["logp_laborSearch"]) * adv_p[t]).sum(),
                         self.sign * (_masked(info["logp_consumption"]) * adv_p[t]).sum()]
                loss = (err * err).sum() + terms[0].double() + terms[1].double() + terms[2].double()
                if t < T - 1:
                    ferr = q_f[t] - info["value_firm"].double()
                    fterms = [self.sign * (_masked(info["logp_firmPurchase"]) * adv_f[t]).sum(),
                              self.sign * (_masked(info["logp_production"]) * adv_f[t]).sum(),
                              self.sign * (_masked(info["logp_offer"]) * adv_f[t]).sum(),
                              self.sign * (_masked(info["logp_jobOffer"]) * adv_f[t]).sum()]
                    fval = (ferr * ferr).sum()
                    loss = loss + fval + fterms[0].double() + fterms[1].double() + fterms[2].double() + fterms[3].double()
                    terms = terms + fterms + [fval]
                (loss * scale).backward()
            with torch.no_grad():
                sums[:len(terms)] += torch.stack([x.detach().double() for x in terms])
                sums[-1] += loss.detach()
        self._all_reduce(sums)
        sums = (sums * scale).tolist()
        for name, value in zip(names, sums[:-1]):
            self.tracked[name] += value
        self.all_optims_step()
        return sums[-1]

    def _all_reduce(self, sums):
        """sum gradients and loss statistics over ranks: one flat bucket, one collective"""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        params = self._params
        for p in params:                       # a rank whose economies were all dropped still takes part
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        flat = torch.cat([p.grad.reshape(-1).float() for p in params] + [sums.float()])
        # statistics ride along in fp32; they are only used for LR scheduling and reporting
        dist.all_reduce(flat, group=self.group)
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
        sums.copy_(flat[off:].double())


def run_episode(pol, orders, out, length, flags=0):
    """Roll the batched policy through `length` env steps and collect the Episode (neural::train's inner loop,
    neuralScenarios.cpp:214-220).  `orders.next()` supplies the visiting orders of each step."""
    ep = Episode()
    for _ in range(length):
        record = []
        info = pol.step(orders.next(), out, flags=flags, record=record)
        snap, draws = record[0]
        ep.append(snap, draws, info, out["p_reward"], out["f_profit"])
    return ep
