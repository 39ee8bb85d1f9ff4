"""Batched policy forward — the caller on the other side of the plugin boundary.

The reference evaluates its 11 decision networks one agent at a time (batch-1 forwards,
/root/reference/src/neural/decisionNetHandler.cpp:49-69, 390-643).  Here the same networks
(/root/reference/src/neural/decisionNets.cpp) carry a leading batch dimension and run for every
agent of every economy at once on torch CUDA tensors; inputs are read zero-copy from the env's
device buffers (BatchedEconomy.device_state_tensors) and outputs are written in the action layout
of include/fastace_b200.h.  Tensor cores are used through torch (TF32 / bf16 autocast) — this file
contains no hand-written kernel; the env step stays the hand-written path.

Parameter names are the reference's (`dimReduce`, `hidden{i}`, `last`, `flatten`, `first`,
`hidden_firstStage{i}`, `hidden_secondStage_{a,b}{i}`, `last_{a,b}`, `offerFlatten`, `jobOfferFlatten`,
sub-modules `offerEncoder` / `jobOfferEncoder`), so a state_dict maps 1:1 onto the reference's
`named_parameters()` (checked against the compiled reference modules in tests/test_policy.py).
"""
import math

import torch
from torch import nn

SQRT2PI = 2 / (math.sqrt(4 / math.pi) * math.sqrt(0.5))  # neuralConstants.h:10: 2 / (M_2_SQRTPI * M_SQRT1_2)


def _xavier(module):
    """xavier_init (decisionNets.cpp:6-12): xavier-normal weights, bias 0.01."""
    if isinstance(module, nn.Linear):
        nn.init.xavier_normal_(module.weight)
        nn.init.constant_(module.bias, 0.01)


_FUSED = False    # set by BatchedPolicy(fused=True) around its no-grad forward: residual stacks run as one CUDA kernel
_TRAIN_LAYERS = False   # set by the trainer: autograd layers with hand-written element-wise halves (train_layers.py)


def _residual_stack(x, layers):
    """x <- x + tanh(h(x)) for every layer (the hidden stack of every decision net)"""
    if (_FUSED and x.is_cuda and not torch.is_grad_enabled() and len(layers) > 0 and x.shape[-1] % 2 == 0
            and x.shape[-1] <= 128 and all(h.in_features == h.out_features == x.shape[-1] for h in layers)):
        from . import fused_mlp
        return fused_mlp.residual_tanh_stack(x, layers)
    if _TRAIN_LAYERS and len(layers) > 0:
        from . import train_layers
        if train_layers.usable(x):
            for h in layers:
                x = train_layers.tanh_layer(x, h, True)
            return x
    for h in layers:
        x = x + torch.tanh(h(x))
    return x


def _body(x, first, layers, last, activation=None):
    """act(last(residual stack(tanh(first(x)))))  — one net body; the fused path runs it as a single kernel"""
    if _FUSED and x.is_cuda and not torch.is_grad_enabled():
        from . import fused_mlp
        if fused_mlp.supports(first, layers, last, x):
            return fused_mlp.net_forward(x, first, layers, last, activation)
    if first is not None:
        if _TRAIN_LAYERS:
            from . import train_layers
            x = train_layers.tanh_layer(x, first, False) if train_layers.usable(x) else torch.tanh(first(x))
        else:
            x = torch.tanh(first(x))
    x = _residual_stack(x, layers)
    if last is None:
        return x
    x = last(x)
    return torch.sigmoid(x) if activation == "sigmoid" else torch.tanh(x) if activation == "tanh" else x


class _Hidden(nn.Module):
    """registers `hidden0..hiddenN-1` (or another prefix) as direct children, like the reference"""

    def _add_layers(self, prefix, sizes):
        layers = []
        for i, (a, b) in enumerate(sizes):
            lin = nn.Linear(a, b)
            self.add_module(f"{prefix}{i}", lin)
            layers.append(lin)
        return layers


class OfferEncoder(_Hidden):
    """decisionNets.cpp:42-80.  x: [..., numFeatures] -> [..., encodingSize]"""

    def __init__(self, stackSize, numFeatures, hiddenSize, numHidden, encodingSize):
        super().__init__()
        self.stackSize, self.numHidden, self.encodingSize = stackSize, numHidden, encodingSize
        self.dimReduce = nn.Linear(numFeatures, hiddenSize)
        self._hidden = self._add_layers("hidden", [(hiddenSize, hiddenSize)] * numHidden)
        self.last = nn.Linear(hiddenSize, encodingSize)
        self.apply(_xavier)

    def forward(self, x):
        return _body(x, self.dimReduce, self._hidden, self.last, "tanh")


class IndexedEncodings:
    """The stack of encodings an agent looks at, given as (table of all encoded offers of its economy, indices):
    table [E, N, enc], idx [E, A, S], valid [E,1,1] (False = empty market, encodings are zeros,
    decisionNetHandler.cpp:542-565).  `flatten` is linear, so flatten(gather(table)) == gather(flatten(table)): the
    nets apply their Linear(enc -> 1) to the N table rows and gather scalars instead of gathering [E,A,S,enc] rows."""

    def __init__(self, table, idx, valid, flat=False):
        self.table, self.idx, self.valid, self.flat = table, idx, valid, flat   # flat: return [E*A, S] rows

    def flattened(self, flatten):
        flat = flatten(self.table).squeeze(-1)                               # [E, N]
        A = self.idx.shape[1]
        picked = torch.gather(flat.unsqueeze(1).expand(-1, A, -1), 2, self.idx)   # [E, A, S]
        out = torch.where(self.valid, picked, flatten.bias.to(picked.dtype))       # flatten(0) = bias
        return out.reshape(-1, out.shape[-1]) if self.flat else out


def _flat(flatten, enc):
    if isinstance(enc, IndexedEncodings):
        return enc.flattened(flatten)
    return flatten(enc).squeeze(-1)


def _head_features(flatten, offerEncodings, *rest):
    x = torch.tanh(_flat(flatten, offerEncodings))
    return torch.cat([x, *rest], dim=-1)


class PurchaseNet(_Hidden):
    """decisionNets.cpp:91-146.  [B,S,enc],[B,U],[B,1],[B,1],[B,G] -> probabilities [B,S]"""

    def __init__(self, offerEncoder, numUtilParams, numGoods, hiddenSize, numHidden):
        super().__init__()
        assert numHidden >= 1
        self.flatten = nn.Linear(offerEncoder.encodingSize, 1)
        nf = offerEncoder.stackSize + numUtilParams + numGoods + 2
        self._hidden = self._add_layers("hidden", [(nf if i == 0 else hiddenSize, hiddenSize) for i in range(numHidden)])
        self.last = nn.Linear(hiddenSize, offerEncoder.stackSize)
        self.apply(_xavier)
        self.offerEncoder = offerEncoder   # registered after apply(), as in the reference (:121)

    def forward(self, offerEncodings, utilParams, budget, labor, inventory):
        x = _head_features(self.flatten, offerEncodings, utilParams, budget, labor, inventory)
        return _body(x, self._hidden[0], self._hidden[1:], self.last, "sigmoid")


class ConsumptionNet(_Hidden):
    """decisionNets.cpp:157-197.  -> [B, G, 2] (mu, log sigma); the reference reshapes to {G,2}"""

    def __init__(self, numUtilParams, numGoods, hiddenSize, numHidden):
        super().__init__()
        self.numGoods = numGoods
        self.first = nn.Linear(numUtilParams + numGoods + 2, hiddenSize)
        self._hidden = self._add_layers("hidden", [(hiddenSize, hiddenSize)] * numHidden)
        self.last = nn.Linear(hiddenSize, numGoods * 2)
        self.apply(_xavier)

    def forward(self, utilParams, money, labor, inventory):
        x = torch.cat([utilParams, money, labor, inventory], dim=-1)
        return _body(x, self.first, self._hidden, self.last).reshape(*x.shape[:-1], self.numGoods, 2)


class OfferNet(_Hidden):
    """decisionNets.cpp:208-300.  -> [B, G, 4] (prop_mu, prop_logsigma, price_mu, price_logsigma)"""

    def __init__(self, offerEncoder, numUtilParams, numGoods, hiddenSize_firstStage, hiddenSize_secondStage,
                 numHidden_firstStage, numHidden_secondStage):
        super().__init__()
        assert numHidden_firstStage >= 1 and numHidden_secondStage >= 1
        self.numGoods = numGoods
        self.flatten = nn.Linear(offerEncoder.encodingSize, 1)
        nf = offerEncoder.stackSize + numUtilParams + numGoods + 2
        h1, h2 = hiddenSize_firstStage, hiddenSize_secondStage
        self._first = self._add_layers("hidden_firstStage", [(nf if i == 0 else h1, h1) for i in range(numHidden_firstStage)])
        sizes2 = [(h1 if i == 0 else h2, h2) for i in range(numHidden_secondStage)]
        # the reference registers a_i and b_i alternately (decisionNets.cpp:245-260)
        self._a, self._b = [], []
        for i, (a, b) in enumerate(sizes2):
            la, lb = nn.Linear(a, b), nn.Linear(a, b)
            self.add_module(f"hidden_secondStage_a{i}", la)
            self.add_module(f"hidden_secondStage_b{i}", lb)
            self._a.append(la)
            self._b.append(lb)
        self.last_a = nn.Linear(h2, numGoods * 2)
        self.last_b = nn.Linear(h2, numGoods * 2)
        self.apply(_xavier)
        self.offerEncoder = offerEncoder

    def forward(self, offerEncodings, utilParams, money, labor, inventory):
        x = _head_features(self.flatten, offerEncodings, utilParams, money, labor, inventory)
        x = _body(x, self._first[0], self._first[1:], None)
        xa = _body(x, None, self._a, self.last_a).reshape(*x.shape[:-1], self.numGoods, 2)
        xb = _body(x, None, self._b, self.last_b).reshape(*x.shape[:-1], self.numGoods, 2)
        return torch.cat([xa, xb], dim=-1)


class JobOfferNet(_Hidden):
    """decisionNets.cpp:318-376.  -> [B, 4] (labor_mu, labor_logsigma, wage_mu, wage_logsigma).
    The reference does NOT register its encoder as a sub-module (:352)."""

    def __init__(self, offerEncoder, numUtilParams, numGoods, hiddenSize, numHidden):
        super().__init__()
        assert numHidden >= 1
        self.flatten = nn.Linear(offerEncoder.encodingSize, 1)
        nf = offerEncoder.stackSize + numUtilParams + numGoods + 2
        self._hidden = self._add_layers("hidden", [(nf if i == 0 else hiddenSize, hiddenSize) for i in range(numHidden)])
        self.last = nn.Linear(hiddenSize, 4)
        self.apply(_xavier)
        object.__setattr__(self, "offerEncoder", offerEncoder)   # plain attribute: not in named_parameters()

    def forward(self, offerEncodings, utilParams, money, labor, inventory):
        x = _head_features(self.flatten, offerEncodings, utilParams, money, labor, inventory)
        return _body(x, self._hidden[0], self._hidden[1:], self.last)


class ValueNet(_Hidden):
    """decisionNets.cpp:387-447.  -> [B, 1]"""

    def __init__(self, offerEncoder, jobOfferEncoder, numUtilParams, numGoods, hiddenSize, numHidden):
        super().__init__()
        assert numHidden >= 1
        self.offerFlatten = nn.Linear(offerEncoder.encodingSize, 1)
        self.jobOfferFlatten = nn.Linear(jobOfferEncoder.encodingSize, 1)
        nf = offerEncoder.stackSize + jobOfferEncoder.stackSize + numUtilParams + numGoods + 2
        self._hidden = self._add_layers("hidden", [(nf if i == 0 else hiddenSize, hiddenSize) for i in range(numHidden)])
        self.last = nn.Linear(hiddenSize, 1)
        self.apply(_xavier)
        self.offerEncoder = offerEncoder
        self.jobOfferEncoder = jobOfferEncoder

    def forward(self, offerEncodings, jobOfferEncodings, utilParams, money, labor, inventory):
        ox = torch.tanh(_flat(self.offerFlatten, offerEncodings))
        jx = torch.tanh(_flat(self.jobOfferFlatten, jobOfferEncodings))
        x = torch.cat([ox, jx, utilParams, money, labor, inventory], dim=-1)
        return _body(x, self._hidden[0], self._hidden[1:], self.last)


NET_NAMES = ["offerEncoder", "jobOfferEncoder", "purchaseNet", "firmPurchaseNet", "laborSearchNet", "consumptionNet",
             "productionNet", "offerNet", "jobOfferNet", "valueNet", "firmValueNet"]


class DecisionNets(nn.Module):
    """The 11 networks of DecisionNetHandler (decisionNetHandler.cpp:126-222), sharing the two encoders."""

    def __init__(self, numGoods=2, stackSize=10, encodingSize=10, hiddenSize=100, nHidden=12, nHiddenSmall=6):
        super().__init__()
        G = numGoods
        self.numGoods, self.stackSize, self.encodingSize = G, stackSize, encodingSize
        U, PF = G + 3, (G + 3) * G
        self.offerEncoder = OfferEncoder(stackSize, G + 1, hiddenSize, nHidden, encodingSize)
        self.jobOfferEncoder = OfferEncoder(stackSize, 2, hiddenSize, nHidden, encodingSize)
        self.purchaseNet = PurchaseNet(self.offerEncoder, U, G, hiddenSize, nHidden)
        self.firmPurchaseNet = PurchaseNet(self.offerEncoder, PF, G, hiddenSize, nHidden)
        self.laborSearchNet = PurchaseNet(self.jobOfferEncoder, U, G, hiddenSize, nHidden)
        self.consumptionNet = ConsumptionNet(U, G, hiddenSize, nHidden)
        self.productionNet = ConsumptionNet(PF, G, hiddenSize, nHidden)
        self.offerNet = OfferNet(self.offerEncoder, PF, G, hiddenSize, hiddenSize, nHidden, nHiddenSmall)
        self.jobOfferNet = JobOfferNet(self.jobOfferEncoder, PF, G, hiddenSize, nHidden)
        self.valueNet = ValueNet(self.offerEncoder, self.jobOfferEncoder, U, G, hiddenSize, nHidden)
        self.firmValueNet = ValueNet(self.offerEncoder, self.jobOfferEncoder, PF, G, hiddenSize, nHidden)

    def net(self, name):
        return getattr(self, name)

    def load_reference_parameters(self, named):
        """named: {"<net>/<reference parameter name>": array}; copies into the matching tensors."""
        with torch.no_grad():
            for key, value in named.items():
                net, pname = key.split("/", 1)
                dict(self.net(net).named_parameters())[pname].copy_(torch.as_tensor(value))


# ---- sampling rules (decisionNetHandler.cpp:27-46, 368-387) ---------------------------
def sample_normal(params, noise=None, detach=False):
    """params[..., 0] = mu, params[..., 1] = log sigma.  Returns (x, log p(x)) with the reference's
    log-density -0.5*((x-mu)/sigma)^2 - log(sigma*sqrt(2*pi)).  detach=False keeps x = eps*sigma + mu in
    the autograd graph like the reference does."""
    mu, sigma = params[..., 0], torch.exp(params[..., 1])
    if noise is None:
        noise = torch.randn_like(mu)
    x = noise * sigma + mu
    if detach:
        x = x.detach()
    logp = -0.5 * ((x - mu) / sigma) ** 2 - torch.log(sigma * SQRT2PI)
    return x, logp


def sample_logit_normal(params, noise=None, detach=False):
    x, logp = sample_normal(params, noise, detach)
    return torch.sigmoid(x), logp       # no Jacobian term, as in the reference (:38-41)


def sample_log_normal(params, noise=None, detach=False):
    x, logp = sample_normal(params, noise, detach)
    return torch.exp(x), logp           # no Jacobian term (:43-46)


def sample_bernoulli(probas, uniform=None):
    """take_i = u_i < p_i ; log pi = sum_i log p_i or log(1-p_i)  (decisionNetHandler.cpp:368-387)"""
    if uniform is None:
        uniform = torch.rand_like(probas)
    take = uniform < probas
    logp = torch.where(take, torch.log(probas), torch.log(1 - probas)).sum(dim=-1)
    return take, logp


# ---- batched decisions: snapshot -> draws -> evaluate ------------------------------------------------
SNAPSHOT_KEYS = ("m_count", "m_good", "m_price", "j_count", "j_wage",
                 "p_money", "p_inv", "p_util_tfp", "p_util_share", "p_util_rho",
                 "f_money", "f_labor", "f_inv", "f_prod_tfp", "f_prod_share", "f_prod_rho")


def snapshot(state):
    """Copy of the state tensors (env layout, fastace_state_t) the decisions of one step read."""
    return {k: state[k].clone() for k in SNAPSHOT_KEYS}


def snapshot_dims(snap):
    E, G, P = snap["p_inv"].shape
    return E, P, snap["f_money"].shape[1], G


def draw(snap, stackSize, generator=None):
    """All random numbers of one step: offer indices (torch::randint(0, count, S) per agent,
    decisionNetHandler.cpp:327-365), the uniforms of the Bernoulli takes (:372, 450) and the standard
    normals of sample_normal (:31).  count == 0 -> index 0 (masked out by `valid` in evaluate)."""
    E, P, F, G = snapshot_dims(snap)
    S, dev = stackSize, snap["p_money"].device
    rand = lambda *shape: torch.rand(*shape, device=dev, generator=generator)
    randn = lambda *shape: torch.randn(*shape, device=dev, generator=generator)

    def indices(count, n_agents):
        c = count.long().clamp(min=1).view(E, 1, 1)
        u = rand(E, n_agents, S)
        return torch.minimum((u * c.to(u.dtype)).long(), c - 1)

    nM, nJ = snap["m_count"], snap["j_count"]
    return {
        "pidxM": indices(nM, P), "pidxJ": indices(nJ, P), "fidxM": indices(nM, F), "fidxJ": indices(nJ, F),
        "u_job": rand(E, P, S), "u_good": rand(E, P, S), "u_fgood": rand(E, F, S),
        "n_cons": randn(E, P, G), "n_prod": randn(E, F, G), "n_amt": randn(E, F, G), "n_price": randn(E, F, G),
        "n_lab": randn(E, F), "n_wage": randn(E, F),
    }


PERSON_OUT = ("p_reward", "p_job_ok", "p_good_ok", "old_j_left", "old_j_taken")
FIRM_OUT = ("f_profit", "f_good_ok", "old_m_left", "old_m_taken")
FIRM_SNAPSHOT_KEYS = ("f_money", "f_labor", "f_inv")      # what the person phase changes of the firms' inputs


CONSUME_SNAPSHOT = {"c_money": "p_money", "c_labor": "p_labor", "c_inv": "p_inv"}   # the consumption decision's inputs


def evaluate(nets, snap, draws, autocast_dtype=None, sample_grad="reference", sides=("persons", "consume", "firms"),
             enc_cache=None, sink=None):
    """Forward of the 11 nets for every agent of every economy + sampling with the given draws.
    Differentiable when autograd is enabled (the trainer re-evaluates recorded steps with it).

    sample_grad: "reference" keeps the reparameterised sample attached to the graph exactly as the
    reference does (decisionNetHandler.cpp:31-32: the quadratic term of the log-density then has zero
    gradient, only -log sigma trains); "score_function" detaches the sample (the textbook estimator).

    sides: which decisions to evaluate — "persons" (value, job search, purchases), "consume", "firms".  Phase-wise
    stepping takes the consumption decision after the trades and the firms' decisions after the person phase; a
    recorded phase-wise step holds those later inputs (c_money / c_labor / c_inv, and the f_* fields), so
    re-evaluating it with all sides reproduces every decision.

    enc_cache: a dict shared by the phase-wise calls of ONE step (no autograd): the offer encodings are computed by the
    first call that needs them and reused by the later ones — the books do not change before the firms post.

    sink: the env's action tensors (fastace_actions_t layout).  Given (no autograd, CUDA), the sampling rules run as the
    kernels of csrc/policy_kernels.cuh, which write the decisions straight into `sink`; `decoded` then holds nothing
    for those keys.

    Returns (decoded, info): decoded = agent-major action tensors, info = log-probabilities [E,agents]
    and state values."""
    E, P, F, G = snapshot_dims(snap)
    S = draws["pidxM"].shape[-1]
    st = snap
    f32 = torch.float32
    ctx = torch.autocast("cuda", dtype=autocast_dtype) if autocast_dtype is not None else _NullCtx()
    with ctx:
        # --- market snapshot + encoder forward (update_encodedOffers / JobOffers, :236-275)
        do_p, do_c, do_f = "persons" in sides, "consume" in sides, "firms" in sides
        nM, nJ = st["m_count"].long(), st["j_count"].long()
        validM = (nM > 0).view(E, 1, 1)
        validJ = (nJ > 0).view(E, 1, 1)
        encM = encJ = None
        if do_p or do_f:                                                             # the consumption net reads no offers
            if enc_cache is not None and "encM" in enc_cache:
                encM, encJ = enc_cache["encM"], enc_cache["encJ"]
            else:
                good = st["m_good"].long().clamp(0, G - 1)
                qty = torch.nn.functional.one_hot(good, G).to(f32)                   # quantities = e_g * 1.0
                feats = torch.cat([qty, st["m_price"].to(f32).unsqueeze(-1)], dim=-1)    # [E, capM, G+1]
                encM = nets.offerEncoder(feats)                                      # [E, capM, enc]
                jfeat = torch.stack([torch.full_like(st["j_wage"], 0.5), st["j_wage"]], dim=-1).to(f32)
                encJ = nets.jobOfferEncoder(jfeat)                                   # [E, F, enc]
                if enc_cache is not None:
                    enc_cache["encM"], enc_cache["encJ"] = encM, encJ

        def gather(enc, idx, valid):    # the agents' stacks of encodings, kept as (table, indices): see IndexedEncodings
            return IndexedEncodings(enc, idx, valid, flat=True)

        # agents as rows of 2-D matrices: nn.Linear then runs as ONE addmm with the bias in the GEMM epilogue
        rows = lambda x: x.reshape(-1, x.shape[-1])

        decoded, info, heads = {}, {}, {}
        if do_p or do_c:
            # --- persons
            pidxM, pidxJ = draws["pidxM"], draws["pidxJ"]
            util = torch.cat([st["p_util_tfp"].unsqueeze(1), st["p_util_share"], st["p_util_rho"].unsqueeze(1)], dim=1)
            util = util.permute(0, 2, 1).to(f32)                                 # [E,P,G+3]: tfp, shares, rho (:30-38)
            money = st["p_money"].to(f32).unsqueeze(-1)
            if "p_labor_input" in st:                                             # tests inject it; the env path uses 0
                labor0 = st["p_labor_input"].to(f32).unsqueeze(-1)
            else:
                labor0 = torch.zeros_like(money)                                  # laborSupplied was just reset (person.cpp:24)
            inv = st["p_inv"].permute(0, 2, 1).to(f32)
            util, money, labor0, inv = rows(util), rows(money), rows(labor0), rows(inv)
            if do_p:
                eM, eJ = gather(encM, pidxM, validM), gather(encJ, pidxJ, validJ)
                heads["p_value"] = nets.valueNet(eM, eJ, util, money, labor0, inv).reshape(E, P)
                heads["p_job_p"] = nets.laborSearchNet(eJ, util, money, labor0, inv).reshape(E, P, S)
                heads["p_good_p"] = nets.purchaseNet(eM, util, money, labor0, inv).reshape(E, P, S)
            if do_c:
                if "c_money" in st:      # phase-wise step: the person's state when it chooses what to consume
                    money = rows(st["c_money"].to(f32).unsqueeze(-1))
                    labor0 = rows(st["c_labor"].to(f32).unsqueeze(-1))
                    inv = rows(st["c_inv"].permute(0, 2, 1).to(f32))
                heads["cons"] = nets.consumptionNet(util, money, labor0, inv).reshape(E, P, G, 2)
        if do_f:
            # --- firms
            fidxM, fidxJ = draws["fidxM"], draws["fidxJ"]
            pf = torch.cat([st["f_prod_tfp"].unsqueeze(2), st["f_prod_share"], st["f_prod_rho"].unsqueeze(2)], dim=2)
            pf = pf.permute(0, 3, 1, 2).reshape(E, F, G * (G + 3)).to(f32)       # per good: tfp, shares, rho (:36-48)
            fmoney = st["f_money"].to(f32).unsqueeze(-1)
            flabor = st["f_labor"].to(f32).unsqueeze(-1)
            finv = st["f_inv"].permute(0, 2, 1).to(f32)
            feM, feJ = gather(encM, fidxM, validM), gather(encJ, fidxJ, validJ)
            pf, fmoney, flabor, finv = rows(pf), rows(fmoney), rows(flabor), rows(finv)
            heads["f_value"] = nets.firmValueNet(feM, feJ, pf, fmoney, flabor, finv).reshape(E, F)
            heads["f_good_p"] = nets.firmPurchaseNet(feM, pf, fmoney, flabor, finv).reshape(E, F, S)
            heads["prod"] = nets.productionNet(pf, fmoney, flabor, finv).reshape(E, F, G, 2)
            heads["offer"] = nets.offerNet(feM, pf, fmoney, flabor, finv).reshape(E, F, G, 4)
            heads["job"] = nets.jobOfferNet(feJ, pf, fmoney, flabor, finv).reshape(E, F, 4)
    # --- sampling (fp32) and the empty-market conventions: purchase log-prob NaN = "no decision", job search 0.0
    #     (decisionNetHandler.cpp:398-403, 476-480)
    detach = sample_grad != "reference"
    if sink is not None and not torch.is_grad_enabled() and st["p_money"].is_cuda:
        from . import fused_sampling as fs
        if do_p:
            lp_job = fs.bernoulli(heads["p_job_p"], draws["u_job"], draws["pidxJ"], validJ, False, sink["p_job_idx"], sink["p_job_take"])
            lp_good = fs.bernoulli(heads["p_good_p"], draws["u_good"], draws["pidxM"], validM, True, sink["p_good_idx"], sink["p_good_take"])
            info.update({"value_person": heads["p_value"].float(), "logp_purchase": lp_good, "logp_laborSearch": lp_job})
        if do_c:
            info["logp_consumption"] = fs.normal(heads["cons"], 0, draws["n_cons"], "logit", sink["p_consume"])
        if do_f:
            lp_fgood = fs.bernoulli(heads["f_good_p"], draws["u_fgood"], draws["fidxM"], validM, True, sink["f_good_idx"], sink["f_good_take"])
            lp_offer = fs.normal(heads["offer"], 0, draws["n_amt"], "logit", sink["f_offer_amt"])
            fs.normal(heads["offer"], 2, draws["n_price"], "log", sink["f_offer_price"], logp=lp_offer)
            lp_jobo = fs.normal(heads["job"], 0, draws["n_lab"], "log", sink["f_job_labor"])
            fs.normal(heads["job"], 2, draws["n_wage"], "log", sink["f_job_wage"], logp=lp_jobo)
            info.update({"value_firm": heads["f_value"].float(), "logp_firmPurchase": lp_fgood,
                         "logp_production": fs.normal(heads["prod"], 0, draws["n_prod"], "logit", sink["f_prod"]),
                         "logp_offer": lp_offer, "logp_jobOffer": lp_jobo})
        return decoded, info
    detach = sample_grad != "reference"
    if do_p:
        p_job_take, lp_job = sample_bernoulli(heads["p_job_p"].float(), draws["u_job"])
        p_good_take, lp_good = sample_bernoulli(heads["p_good_p"].float(), draws["u_good"])
        decoded.update({"p_job_idx": draws["pidxJ"], "p_job_take": p_job_take & validJ, "p_good_idx": draws["pidxM"],
                        "p_good_take": p_good_take & validM})
        info.update({
            "value_person": heads["p_value"].float(),
            "logp_purchase": torch.where(validM.view(E, 1), lp_good, torch.full_like(lp_good, float("nan"))),
            "logp_laborSearch": torch.where(validJ.view(E, 1), lp_job, torch.zeros_like(lp_job)),
        })
    if do_c:
        cons_x, lp_cons = sample_logit_normal(heads["cons"].float(), draws["n_cons"], detach)
        decoded["p_consume"] = cons_x
        info["logp_consumption"] = lp_cons.sum(-1)
    if do_f:
        f_good_take, lp_fgood = sample_bernoulli(heads["f_good_p"].float(), draws["u_fgood"])
        prod_x, lp_prod = sample_logit_normal(heads["prod"].float(), draws["n_prod"], detach)
        offer, job = heads["offer"].float(), heads["job"].float()
        amt_x, lp_amt = sample_logit_normal(offer[..., 0:2], draws["n_amt"], detach)
        price_x, lp_price = sample_log_normal(offer[..., 2:4], draws["n_price"], detach)
        lab_x, lp_lab = sample_log_normal(job[..., 0:2], draws["n_lab"], detach)
        wage_x, lp_wage = sample_log_normal(job[..., 2:4], draws["n_wage"], detach)
        decoded.update({"f_good_idx": draws["fidxM"], "f_good_take": f_good_take & validM, "f_prod": prod_x,
                        "f_offer_amt": amt_x, "f_offer_price": price_x, "f_job_labor": lab_x, "f_job_wage": wage_x})
        info.update({
            "value_firm": heads["f_value"].float(),
            "logp_firmPurchase": torch.where(validM.view(E, 1), lp_fgood, torch.full_like(lp_fgood, float("nan"))),
            "logp_production": lp_prod.sum(-1),
            "logp_offer": lp_amt.sum(-1) + lp_price.sum(-1),
            "logp_jobOffer": lp_lab + lp_wage,
        })
    return decoded, info


class BatchedPolicy:
    """Runs the 7 decisions of every agent of every economy at once and writes the action tensors
    of one env step (fastace_actions_t layout).

    Semantics note (documented deviation): the reference asks the nets in the middle
    of an agent's turn, so e.g. a person's purchase decision sees its money after its job search.
    A batched forward cannot interleave with the matching inside one fused env step, so here all
    decisions of a step are taken from the state at the start of the step (labour input 0 for
    persons, as in the reference's first decision of a turn).  The networks, the index draws, the
    sampling rules and the action decode are the reference's."""

    def __init__(self, env, nets, generator=None, autocast_dtype=None, fused=False, two_phase=False):
        """fused=True: the residual hidden stacks run through the hand-written kernel (csrc/mlp_stack.cuh, bf16
        tensor-core operands, fp32 accumulate/residual) instead of eager torch — a rollout-only fast mode.
        two_phase=True (phase-wise stepping): step() runs the persons' trades (FASTACE_STEP_PERSONS_TRADE), takes the
        consumption decision from each person's money / laborSupplied / inventory after its trades
        (FASTACE_STEP_PERSONS_CONSUME), then takes the firms' decisions from the state the firms actually see in the
        reference — after every person has acted: money after sales and hires, inventories after sales, laborHired of
        this step — and completes the step (FASTACE_STEP_FIRMS)."""
        self.env, self.nets, self.gen, self.autocast_dtype, self.fused = env, nets, generator, autocast_dtype, fused
        self.two_phase = two_phase
        self.E, self.P, self.F, self.G, self.S = env.dims.tuple
        self.state = env.device_state_tensors()
        self.dev = self.state["p_money"].device
        self.actions = env.alloc_actions()
        self.packed = env.pack_device("actions", self.actions)

    @torch.no_grad()
    def decide(self, perms, record=None):
        """Fill self.actions for one step.  perms = (perm_person, perm_firm) int32 [E,P] / [E,F] (host or device).
        Returns a dict with log-probabilities and state values (device tensors); when `record` is a list,
        (snapshot, draws) of the step is appended to it so a trainer can re-evaluate the step with autograd."""
        global _FUSED
        snap = snapshot(self.state) if record is not None else self.state
        draws = draw(snap, self.S, self.gen)
        _FUSED = self.fused
        try:
            decoded, info = evaluate(self.nets, snap, draws, self.autocast_dtype, sink=self.actions if self.fused else None)
        finally:
            _FUSED = False
        if record is not None:
            record.append((snap, draws))
        a = self.actions
        a["perm_person"].copy_(torch.as_tensor(perms[0]))
        a["perm_firm"].copy_(torch.as_tensor(perms[1]))
        for key, value in decoded.items():
            a[key].copy_(value if value.dim() == 2 else value.permute(0, 2, 1))   # agent-major -> [E][slot|good][agent]
        return info

    def _write(self, decoded):
        for key, value in decoded.items():
            self.actions[key].copy_(value if value.dim() == 2 else value.permute(0, 2, 1))   # agent-major -> [E][slot|good][agent]

    def _evaluate(self, snap, draws, sides, enc_cache=None):
        global _FUSED
        _FUSED = self.fused
        try:
            return evaluate(self.nets, snap, draws, self.autocast_dtype, sides=sides, enc_cache=enc_cache,
                            sink=self.actions if self.fused else None)
        finally:
            _FUSED = False

    @torch.no_grad()
    def step(self, perms, out, flags=0, record=None):
        """decide + one env step (device path, current stream)."""
        if not self.two_phase:
            info = self.decide(perms, record)
            self.env.time_step(self.packed, out, flags=flags)
            return info
        from . import _abi

        def phase_out(names):                         # each call gets only its own output arrays
            if isinstance(out, _abi.StepOut):
                sub = _abi.StepOut()
                for n in names:
                    setattr(sub, n, getattr(out, n))
                sub._keepalive = out
                return sub
            return {k: v for k, v in out.items() if k in names}

        a = self.actions
        a["perm_person"].copy_(torch.as_tensor(perms[0]))
        a["perm_firm"].copy_(torch.as_tensor(perms[1]))
        snap = snapshot(self.state) if record is not None else self.state
        draws = draw(snap, self.S, self.gen)          # the books (and so the index ranges) do not change before the firms post
        enc_cache = {}                                # offer encodings of the step: computed once, used by persons and firms
        decoded, info = self._evaluate(snap, draws, ("persons",), enc_cache)
        self._write(decoded)
        self.env.time_step(self.packed, phase_out(("p_job_ok", "p_good_ok", "old_j_left", "old_j_taken")),
                           flags=flags | _abi.STEP_PERSONS_TRADE)
        # consumption: decided from the person's money, laborSupplied and inventory after its trades
        # (neuralPersonDecisionMaker.cpp:93-111)
        live = {k: self.state[v] for k, v in CONSUME_SNAPSHOT.items()}
        if record is not None:
            live = {k: v.clone() for k, v in live.items()}
        snap = dict(snap, **live)
        decoded_c, info_c = self._evaluate(snap, draws, ("consume",))
        self._write(decoded_c)
        info.update(info_c)
        self.env.time_step(self.packed, phase_out(("p_reward",)), flags=flags | _abi.STEP_PERSONS_CONSUME)
        if record is not None:
            for k in FIRM_SNAPSHOT_KEYS:              # the firms' inputs as they stand after the person phase
                snap[k] = self.state[k].clone()
            record.append((snap, draws))
        decoded_f, info_f = self._evaluate(snap, draws, ("firms",), enc_cache)
        self._write(decoded_f)
        self.env.time_step(self.packed, phase_out(FIRM_OUT), flags=flags | _abi.STEP_FIRMS)
        info.update(info_f)
        return info


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
