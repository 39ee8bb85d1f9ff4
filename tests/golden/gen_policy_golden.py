"""Generates tests/golden/policy_nets.npz from the reference's UNMODIFIED LibTorch decision networks
(oracle/_ref/libfastace_refnets.so, built by `make -C oracle refnets`): every named parameter of the 11
nets (tiny sizes so the fixture stays small) plus batch-1 forward results on random inputs."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(ROOT, "oracle", "_ref", "libfastace_refnets.so")

CFG = dict(stackSize=5, encodingSize=4, hiddenSize=16, nHidden=3, nHiddenSmall=2, numGoods=2)
fp = C.POINTER(C.c_float)


def f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(fp)


def load():
    L = C.CDLL(LIB)
    L.refnets_create.restype = C.c_void_p
    L.refnets_create.argtypes = [C.c_int] * 6 + [C.c_uint64]
    L.refnets_num_params.argtypes = [C.c_void_p]
    L.refnets_param_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    L.refnets_param_data.argtypes = [C.c_void_p, C.c_int, fp]
    L.refnets_encode.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, fp]
    L.refnets_purchase.argtypes = [C.c_void_p, C.c_int, fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]
    L.refnets_consumption.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, C.c_float, C.c_float, fp, fp]
    L.refnets_offer.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]
    L.refnets_joboffer.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]
    L.refnets_value.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]
    return L


def export_params(L, h):
    out = {}
    shape = (C.c_int64 * 2)()
    name = C.create_string_buffer(256)
    for i in range(L.refnets_num_params(h)):
        n = L.refnets_param_info(h, i, name, 256, shape)
        buf = np.zeros(n, dtype=np.float32)
        L.refnets_param_data(h, i, buf.ctypes.data_as(fp))
        shp = (shape[0], shape[1]) if shape[1] else (shape[0],)
        out[name.value.decode()] = buf.reshape(shp)
    return out


def forwards(L, h, cfg, rng, nsamples):
    S, enc, G = cfg["stackSize"], cfg["encodingSize"], cfg["numGoods"]
    U, PF = G + 3, (G + 3) * G
    d = {}
    x, px = f32(rng.normal(size=(7, G + 1))); y = np.zeros((7, enc), np.float32)
    L.refnets_encode(h, 0, px, 7, y.ctypes.data_as(fp)); d["enc_goods_in"], d["enc_goods_out"] = x, y
    x, px = f32(rng.normal(size=(6, 2))); y = np.zeros((6, enc), np.float32)
    L.refnets_encode(h, 1, px, 6, y.ctypes.data_as(fp)); d["enc_jobs_in"], d["enc_jobs_out"] = x, y
    oe = rng.normal(size=(nsamples, S, enc)).astype(np.float32)
    je = rng.normal(size=(nsamples, S, enc)).astype(np.float32)
    up = rng.normal(size=(nsamples, U)).astype(np.float32)
    fpar = rng.normal(size=(nsamples, PF)).astype(np.float32)
    money = rng.uniform(1, 20, nsamples).astype(np.float32)
    labor = rng.uniform(0, 1, nsamples).astype(np.float32)
    inv = rng.uniform(0, 10, size=(nsamples, G)).astype(np.float32)
    d.update(oe=oe, je=je, up=up, fpar=fpar, money=money, labor=labor, inv=inv)
    res = {k: [] for k in ("purchase", "firmPurchase", "laborSearch", "consumption", "production", "offer", "joboffer", "value", "firmValue")}
    for i in range(nsamples):
        def call(fn, *args, shape):
            y = np.zeros(shape, np.float32)
            fn(*args, y.ctypes.data_as(fp))
            return y
        a = lambda v: np.ascontiguousarray(v).ctypes.data_as(fp)
        res["purchase"].append(call(L.refnets_purchase, h, 0, a(oe[i]), a(up[i]), U, money[i], labor[i], a(inv[i]), shape=(S,)))
        res["firmPurchase"].append(call(L.refnets_purchase, h, 1, a(oe[i]), a(fpar[i]), PF, money[i], labor[i], a(inv[i]), shape=(S,)))
        res["laborSearch"].append(call(L.refnets_purchase, h, 2, a(je[i]), a(up[i]), U, money[i], labor[i], a(inv[i]), shape=(S,)))
        res["consumption"].append(call(L.refnets_consumption, h, 0, a(up[i]), U, money[i], labor[i], a(inv[i]), shape=(G, 2)))
        res["production"].append(call(L.refnets_consumption, h, 1, a(fpar[i]), PF, money[i], labor[i], a(inv[i]), shape=(G, 2)))
        res["offer"].append(call(L.refnets_offer, h, a(oe[i]), a(fpar[i]), PF, money[i], labor[i], a(inv[i]), shape=(G, 4)))
        res["joboffer"].append(call(L.refnets_joboffer, h, a(je[i]), a(fpar[i]), PF, money[i], labor[i], a(inv[i]), shape=(4,)))
        res["value"].append(call(L.refnets_value, h, 0, a(oe[i]), a(je[i]), a(up[i]), U, money[i], labor[i], a(inv[i]), shape=(1,)))
        res["firmValue"].append(call(L.refnets_value, h, 1, a(oe[i]), a(je[i]), a(fpar[i]), PF, money[i], labor[i], a(inv[i]), shape=(1,)))
    for k, v in res.items():
        d["out_" + k] = np.stack(v)
    return d


def main():
    L = load()
    h = L.refnets_create(CFG["stackSize"], CFG["encodingSize"], CFG["hiddenSize"], CFG["nHidden"], CFG["nHiddenSmall"], CFG["numGoods"], 1234)
    data = {"cfg/" + k: np.array([v]) for k, v in CFG.items()}
    for k, v in export_params(L, h).items():
        data["param/" + k] = v
    data.update(forwards(L, h, CFG, np.random.default_rng(0), 6))
    path = os.path.join(HERE, "policy_nets.npz")
    np.savez_compressed(path, **data)
    print("policy_nets", os.path.getsize(path) // 1024, "KiB", sum(1 for k in data if k.startswith("param/")), "parameters")


if __name__ == "__main__":
    main()
