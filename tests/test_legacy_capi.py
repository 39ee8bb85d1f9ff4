"""`run` / `train`, the two C entry points of libpybindings.so that the reference's Python front end binds
(/root/reference/src/pybindings.h:16-27, py/main.py:96-126).

CPU part (here): the symbols are exported with the reference's calling convention, and — where /root/reference is
present — the reference's UNMODIFIED py/main.py binds against this library (prototypes, struct layouts, defaults) as
it does against its own.  GPU part: train() through the same ctypes prototypes for two episodes."""
import ctypes as C
import os
import sys
import types

import numpy as np
import pytest

from fastace_b200 import _abi, lib

REF_MAIN = "/root/reference/py/main.py"


def test_symbols_and_struct_layouts(native_lib):
    L = native_lib
    assert hasattr(L, "run") and hasattr(L, "train")
    assert C.sizeof(_abi.CustomScenarioParams) == 344 and C.sizeof(_abi.TrainingParams) == 136


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="needs the reference checkout (this container only)")
def test_unmodified_reference_front_end_binds(native_lib, tmp_path, monkeypatch):
    """py/main.py does ctypes.CDLL("../bin/libpybindings.so") at import: give it this library under that name."""
    (tmp_path / "bin").mkdir()
    (tmp_path / "py").mkdir()
    os.symlink(lib.LIB_PATH, tmp_path / "bin" / "libpybindings.so")
    monkeypatch.chdir(tmp_path / "py")
    for name in ("matplotlib", "matplotlib.pyplot"):          # plotting is not installed here and not exercised
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    src = open(REF_MAIN).read()
    mod = types.ModuleType("reference_main")
    mod.__file__ = str(tmp_path / "py" / "main.py")           # it chdirs next to "itself" and opens ../bin/libpybindings.so
    exec(compile(src, REF_MAIN, "exec"), mod.__dict__)        # the file itself, unmodified, never copied
    # its ctypes mirrors are byte-compatible with ours and its factory calls return the reference's defaults
    assert C.sizeof(mod.CustomScenarioParams) == C.sizeof(_abi.CustomScenarioParams)
    assert C.sizeof(mod.TrainingParams) == C.sizeof(_abi.TrainingParams)
    sp = mod.lib.create_scenario_params(48, 12)
    tp = mod.lib.create_training_params()
    assert (sp.numPeople, sp.numFirms) == (48, 12) and sp.money_mu == 10.0 and sp.firm_money_mu == 50.0
    assert tp.stackSize == 10 and tp.hiddenSize == 100 and tp.nHidden == 12 and tp.episodeLength == 20
    assert mod.lib.train.argtypes[3] is C.c_bool and mod.lib.run.restype is None


@pytest.mark.gpu
def test_train_through_the_reference_prototypes(native_lib, tmp_path, monkeypatch):
    """what py/main.py's train() wrapper does (py/main.py:112-126), against this library: numEpisodes finite losses,
    the eleven checkpoint files under ../models/, learning rates written back"""
    (tmp_path / "py").mkdir()
    monkeypatch.chdir(tmp_path / "py")
    monkeypatch.setenv("FASTACE_NUM_ECONOMIES", "8")
    L = C.CDLL(lib.LIB_PATH)
    L.create_scenario_params.argtypes = [C.c_uint, C.c_uint]
    L.create_scenario_params.restype = _abi.CustomScenarioParams
    L.create_training_params.restype = _abi.TrainingParams
    L.train.argtypes = [C.POINTER(C.c_double), C.POINTER(_abi.CustomScenarioParams), C.POINTER(_abi.TrainingParams), C.c_bool, C.c_double]
    L.train.restype = None
    sp = L.create_scenario_params(12, 3)
    tp = L.create_training_params()
    tp.numEpisodes, tp.episodeLength, tp.checkpointEveryNEpisodes, tp.updateEveryNEpisodes = 2, 5, 1, 1
    tp.hiddenSize, tp.nHidden, tp.nHiddenSmall = 32, 2, 1
    lr_before = tp.purchaseNetLR
    losses = C.ARRAY(C.c_double, tp.numEpisodes)()
    L.train(losses, C.POINTER(_abi.CustomScenarioParams)(sp), C.POINTER(_abi.TrainingParams)(tp), False, 0.0)
    assert all(np.isfinite(list(losses))), list(losses)
    files = sorted(os.listdir(tmp_path / "models"))
    assert len(files) == 11 and all(f.endswith(".pt") for f in files), files
    assert tp.purchaseNetLR > 0 and np.isfinite(tp.purchaseNetLR) and lr_before > 0
