"""Profiling aid (loads fastace_b200/libfastace_b200_timing.so, built here with
`python -c "from fastace_b200 import build; build.build(extra_flags=['-DFASTACE_CTA_TIMING'], out='fastace_b200/libfastace_b200_timing.so')"`): when does every economy's warp of match_kernel
start and end, and on which SM?  Prints the launch ramp, the spread of the warps' durations and the idle share."""
import ctypes as C
import sys

import numpy as np

from fastace_b200 import _abi, lib, scenario
from fastace_b200.env import BatchedEconomy

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 12
COMPACT = len(sys.argv) > 3 and sys.argv[3] == "compact"
DIMS = (E, 100, 10, 2, 10)
import os
lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(lib.__file__)), "libfastace_b200_timing.so")
L = lib.load()
env = BatchedEconomy(DIMS)
env.set_state(scenario.custom_initial_state(DIMS, 2024)[0])
orders = scenario.OrderStream(DIMS, 7)
import torch
for t in range(STEPS):
    act = scenario.synthetic_actions(DIMS, seed=99, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
    if COMPACT:      # the call the specialised match_kernel<G, kModeCompact ...> is compiled for
        st = env.get_state()
        cz = _abi.compact_actions_for_counts(act, st["j_count"], st["m_count"], True)
        dact = env.pack_device("compact", env.alloc_compact_actions(cz))
    else:
        dact = env.alloc_actions(act)
    dout = env.alloc_outputs()
    torch.cuda.synchronize()
    env.time_step(dact, dout, flags=_abi.IDX_MODULO)
    torch.cuda.synchronize()
    buf = np.zeros((E, 3), dtype=np.uint64)
    rc = L.fastace_debug_cta_times(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), E)
    assert rc == 0
    t0 = buf[:, 0].astype(np.int64); t1 = buf[:, 1].astype(np.int64); sm = buf[:, 2].astype(np.int64)
    k0 = t0.min(); dur = (t1 - t0) / 1e3; start = (t0 - k0) / 1e3; end = (t1 - k0) / 1e3
    sm_end = np.array([end[sm == s].max() for s in np.unique(sm)])
    sm_start = np.array([start[sm == s].min() for s in np.unique(sm)])
    print(f"step {t:2d}: kernel {end.max():6.1f} us | starts p50 {np.percentile(start,50):5.1f} p90 {np.percentile(start,90):5.1f} max {start.max():5.1f} | "
          f"warp duration mean {dur.mean():5.1f} p10 {np.percentile(dur,10):5.1f} p50 {np.percentile(dur,50):5.1f} p90 {np.percentile(dur,90):5.1f} max {dur.max():5.1f} | "
          f"ends mean {end.mean():5.1f} p50 {np.percentile(end,50):5.1f} | SM last-end mean {sm_end.mean():5.1f} min {sm_end.min():5.1f} | SMs {len(sm_end)} CTAs/SM max {np.bincount(sm).max()}")
np.save("gpurun_out/cta_times.npy", buf)
