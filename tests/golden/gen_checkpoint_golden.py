"""Generates tests/golden/ref_checkpoint/: the eleven `.pt` files written by the reference's own torch::save path
(oracle/_ref/libfastace_refnets.so -> nn::Module::save, as DecisionNetHandler::save_models does) for a tiny
architecture, plus params.npz with the same parameters exported through named_parameters()."""
import ctypes as C
import os
import sys

import numpy as np
import torch  # noqa: F401  (the harness shares the process-wide libtorch; serialisation needs it initialised)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_policy_golden as gen  # noqa: E402

OUT = os.path.join(HERE, "ref_checkpoint")


def bind(L):
    L.refnets_save.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.refnets_load.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.refnets_net_name.restype = C.c_char_p
    L.refnets_net_name.argtypes = [C.c_void_p, C.c_int]
    return L


if __name__ == "__main__":
    L = bind(gen.load())
    cfg = gen.CFG
    h = L.refnets_create(cfg["stackSize"], cfg["encodingSize"], cfg["hiddenSize"], cfg["nHidden"], cfg["nHiddenSmall"], cfg["numGoods"], 5)
    os.makedirs(OUT, exist_ok=True)
    for i in range(11):
        name = L.refnets_net_name(h, i).decode()
        assert L.refnets_save(h, i, os.path.join(OUT, name + ".pt").encode()) == 0
    np.savez_compressed(os.path.join(OUT, "params.npz"), **gen.export_params(L, h),
                        **{"cfg/" + k: np.array([v]) for k, v in cfg.items()})
    print("wrote", OUT, sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
