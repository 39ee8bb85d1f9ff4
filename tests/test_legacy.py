"""py/main.py's `train` / `run` on the batched env (fastace_b200/legacy.py) and the checkpoint helpers."""
import math
import os

import numpy as np
import pytest
import torch

from fastace_b200 import legacy, policy, scenario


def _small_nets():
    torch.manual_seed(0)
    return policy.DecisionNets(numGoods=2, stackSize=4, encodingSize=3, hiddenSize=8, nHidden=2, nHiddenSmall=1)


def test_checkpoint_round_trip_uses_reference_file_names(tmp_path):
    nets = _small_nets()
    legacy.save_models(nets, str(tmp_path) + "/")
    assert sorted(os.listdir(tmp_path)) == sorted(n + ".pt" for n in policy.NET_NAMES)   # decisionNetHandler.cpp:726-738
    other = _small_nets()
    with torch.no_grad():
        for p in other.parameters():
            p.add_(1.0)
    legacy.load_models(other, str(tmp_path) + "/")
    for (k, a), (_, b) in zip(nets.named_parameters(), other.named_parameters()):
        assert torch.equal(a, b), k


def test_perturb_models_keeps_xavier_variance():
    """perturb_layer (decisionNets.cpp:14-39): W <- sqrt(1-pct) W + sqrt(xavier_var*pct) N(0,1); biases untouched"""
    torch.manual_seed(1)
    nets = policy.DecisionNets(numGoods=2, hiddenSize=100, nHidden=2, nHiddenSmall=1)
    w0 = nets.valueNet.hidden1.weight.detach().clone()
    b0 = nets.valueNet.hidden1.bias.detach().clone()
    enc0 = nets.offerEncoder.hidden0.weight.detach().clone()
    gen = torch.Generator(); gen.manual_seed(3)
    legacy.perturb_models(nets, 0.25, generator=gen)
    w1 = nets.valueNet.hidden1.weight.detach()
    assert torch.equal(nets.valueNet.hidden1.bias.detach(), b0)
    xavier_var = 2.0 / 200
    noise = w1 - math.sqrt(0.75) * w0
    assert abs(noise.var().item() / (xavier_var * 0.25) - 1) < 0.1
    assert abs(w1.var().item() / xavier_var - 1) < 0.1
    # the shared encoder is perturbed exactly once although five nets hold it
    enc_noise = nets.offerEncoder.hidden0.weight.detach() - math.sqrt(0.75) * enc0
    assert abs(enc_noise.var().item() / (xavier_var * 0.25) - 1) < 0.1
    legacy.perturb_models(nets, 0.0)
    assert torch.equal(nets.valueNet.hidden1.weight.detach(), w1)


@pytest.mark.gpu
def test_train_and_run_like_py_main(tmp_path, capsys):
    sp = scenario.scenario_params(20, 4)
    tp = scenario.training_params()
    tp.numEpisodes, tp.episodeLength, tp.updateEveryNEpisodes, tp.checkpointEveryNEpisodes = 4, 5, 2, 2
    tp.stackSize, tp.encodingSize, tp.hiddenSize, tp.nHidden, tp.nHiddenSmall = 4, 3, 16, 2, 1
    tp.episodeBatchSizeForLRDecay, tp.patienceForLRDecay = 1, 1
    tp.purchaseNetLR = 1e-4
    save = str(tmp_path) + "/"
    torch.manual_seed(0)
    losses = legacy.train(sp, tp, numEconomies=16, saveDir=save, trainer_kwargs=dict(wiring="intended"))
    out = capsys.readouterr().out
    assert len(losses) == 4 and all(np.isfinite(losses))
    assert "Episode 2: Average loss over past 2 episodes" in out and "Episode 4:" in out
    assert os.path.exists(save + "offerNet.pt") and os.path.exists(save + "firmValueNet.pt")
    assert tp.purchaseNetLR != 1e-4 or tp.valueNetLR != 1e-5        # schedulers wrote the learning rates back
    # continue from the checkpoint with perturbed weights, then roll the saved nets out
    more = legacy.train(sp, tp, fromPretrained=True, perturbationSize=0.1, numEconomies=16, saveDir=save, quiet=True)
    assert len(more) == 4
    log = legacy.run(sp, tp, saveDir=save, quiet=True)
    # print_info (src/pybindings.cpp:60-75) after every step: the book posted during that step
    assert len(log) == 5 and log[0].startswith("Time = 1:") and log[4].startswith("Time = 5:")
    assert all(("Avg. price" in e or "[No offers]" in e) and ("Avg. wage" in e or "[No job offers]" in e) for e in log)


def test_reads_checkpoints_written_by_the_reference():
    """tests/golden/ref_checkpoint/*.pt were written by the reference's own torch::save (gen_checkpoint_golden.py)"""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_checkpoint")
    z = np.load(os.path.join(d, "params.npz"))
    cfg = {k.split("/", 1)[1]: int(z[k][0]) for k in z.files if k.startswith("cfg/")}
    nets = policy.DecisionNets(**cfg)
    legacy.load_models(nets, d + "/")
    checked = 0
    for key in z.files:
        if key.startswith("cfg/"):
            continue
        net, pname = key.split("/", 1)
        assert np.array_equal(dict(nets.net(net).named_parameters())[pname].detach().numpy(), z[key]), key
        checked += 1
    assert checked > 100


def test_reference_reads_checkpoints_written_here(tmp_path):
    """save_models_reference_format -> the reference's torch::load (compiled reference modules, when built here)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    gen = pytest.importorskip("gen_policy_golden")
    ck = pytest.importorskip("gen_checkpoint_golden")
    if not os.path.exists(gen.LIB):
        pytest.skip("oracle/_ref/libfastace_refnets.so not built here")
    L = ck.bind(gen.load())
    if not hasattr(L, "refnets_load"):
        pytest.skip("prebuilt harness without refnets_load")
    cfg = gen.CFG
    torch.manual_seed(9)
    nets = policy.DecisionNets(**cfg)
    legacy.save_models_reference_format(nets, str(tmp_path) + "/")
    h = L.refnets_create(cfg["stackSize"], cfg["encodingSize"], cfg["hiddenSize"], cfg["nHidden"], cfg["nHiddenSmall"], cfg["numGoods"], 1)
    for i in range(11):
        name = L.refnets_net_name(h, i).decode()
        assert L.refnets_load(h, i, os.path.join(str(tmp_path), name + ".pt").encode()) == 0, name
    for key, value in gen.export_params(L, h).items():
        net, pname = key.split("/", 1)
        assert np.array_equal(dict(nets.net(net).named_parameters())[pname].detach().numpy(), value), key
    # and the round trip through our own reader
    other = policy.DecisionNets(**cfg)
    legacy.load_models(other, str(tmp_path) + "/")
    for (k, a), (_, b) in zip(nets.named_parameters(), other.named_parameters()):
        assert torch.equal(a, b), k
