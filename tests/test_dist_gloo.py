"""N>1 path on CPU: world_size-2 gloo.  Economies shard by index with no data-path
collective; each rank steps only its block (here with the CPU oracle standing in for the
GPU, which this container does not have) and the union must equal the single-process run.
Also covers the max-over-ranks timing reduction used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastace_b200 import _abi, scenario, sharding

DIMS_TOTAL = (7, 20, 4, 2, 6)  # 7 economies over 2 ranks -> ragged shards 4 + 3
STEPS = 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice(d, lo, hi):
    return {k: np.ascontiguousarray(v[lo:hi]) for k, v in d.items()}


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.loader import Oracle
    orc = Oracle()
    E, P, F, G, S = DIMS_TOTAL
    lo, hi = sharding.shard_range(E, rank, world)
    dims = (hi - lo, P, F, G, S)
    full_state = scenario.custom_initial_state(DIMS_TOTAL, 31)[0]
    st = _slice(full_state, lo, hi)
    orders = scenario.OrderStream(DIMS_TOTAL, 8)
    rewards = []
    for t in range(STEPS):
        act = _slice(scenario.synthetic_actions(DIMS_TOTAL, seed=2, step=t, perms=orders.next(), **scenario.BENCH_PRESET), lo, hi)
        out = _abi.alloc_host("out", dims, names=("p_reward", "f_profit"))
        orc.step(dims, st, act, out, flags=_abi.IDX_MODULO, time_before=t)
        rewards.append(out["p_reward"].copy())
    dist.barrier()
    # timing reduction: every rank must see the max
    t_max = sharding.reduce_max(10.0 + rank, dist)
    assert t_max == 10.0 + world - 1
    # no data-path collective was needed; gather only to check the result
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), lo=lo, hi=hi, p_money=st["p_money"], m_count=st["m_count"],
             rewards=np.stack(rewards))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-process run over all economies
    st = scenario.custom_initial_state(DIMS_TOTAL, 31)[0]
    orders = scenario.OrderStream(DIMS_TOTAL, 8)
    rewards = []
    for t in range(STEPS):
        act = scenario.synthetic_actions(DIMS_TOTAL, seed=2, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        out = _abi.alloc_host("out", DIMS_TOTAL, names=("p_reward", "f_profit"))
        oracle.step(DIMS_TOTAL, st, act, out, flags=_abi.IDX_MODULO, time_before=t)
        rewards.append(out["p_reward"].copy())
    rewards = np.stack(rewards)
    covered = np.zeros(DIMS_TOTAL[0], dtype=bool)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        assert not covered[lo:hi].any()
        covered[lo:hi] = True
        assert np.array_equal(z["p_money"], st["p_money"][lo:hi])
        assert np.array_equal(z["m_count"], st["m_count"][lo:hi])
        assert np.array_equal(z["rewards"], rewards[:, lo:hi])
    assert covered.all()


@pytest.mark.parametrize("total,world", [(4096, 8), (7, 2), (5, 8), (65536, 8), (1, 1)])
def test_shard_ranges_partition(total, world):
    seen = []
    for r in range(world):
        lo, hi = sharding.shard_range(total, r, world)
        seen.extend(range(lo, hi))
        for e in (lo, hi - 1):
            if lo < hi:
                assert sharding.owner_of(e, total, world) == r
    assert seen == list(range(total))
