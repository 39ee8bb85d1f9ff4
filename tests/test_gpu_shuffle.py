"""Visiting orders generated on the device == libstdc++'s std::shuffle with std::minstd_rand0 on the host
(fastace_shuffle_orders, /root/reference/src/base/economy.cpp:110-111), cumulatively over 40 steps."""
import numpy as np
import pytest

from fastace_b200 import scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,chunks", [
    ((1, 48, 12, 2, 10), [1] * 40),              # config A shape
    ((300, 100, 10, 2, 10), [7, 1, 12, 20]),     # config B shape, several steps per launch
    ((3, 1, 1, 1, 1), [40]),
    ((2, 3000, 700, 8, 10), [5, 5]),             # shared-memory variant at its upper end
])
def test_device_orders_equal_std_shuffle(native_lib, dims, chunks):
    import torch
    from fastace_b200.env import BatchedEconomy
    env = BatchedEconomy(dims)
    host = scenario.OrderStream(dims, 4242)
    first = True
    for steps in chunks:
        pp, pf = env.shuffle_orders(seed=4242, restart=first, steps=steps)
        pp16 = torch.empty(pp.shape, dtype=torch.int16, device=pp.device)
        first = False
        torch.cuda.synchronize()
        pp, pf = pp.cpu().numpy(), pf.cpu().numpy()
        for t in range(steps):
            hp, hf = host.next()
            assert np.array_equal(pp[t], hp) and np.array_equal(pf[t], hf), (dims, t)
    env.close()


def test_device_orders_16bit_form_and_restart(native_lib):
    import torch
    from fastace_b200.env import BatchedEconomy
    dims = (65, 100, 10, 2, 10)
    env = BatchedEconomy(dims)
    dev = torch.device("cuda", 0)
    for seed in (1, 99):                        # a restart re-seeds and starts from the identity order
        host = scenario.OrderStream(dims, seed)
        pp = torch.empty((6, 65, 100), dtype=torch.int16, device=dev)
        pf = torch.empty((6, 65, 10), dtype=torch.int16, device=dev)
        env.shuffle_orders(seed=seed, restart=True, steps=6, perm_person=pp, perm_firm=pf)
        torch.cuda.synchronize()
        for t in range(6):
            hp, hf = host.next()
            assert np.array_equal(pp[t].cpu().numpy().view(np.uint16), hp) and np.array_equal(pf[t].cpu().numpy().view(np.uint16), hf)
    env.close()


def test_device_orders_config_d_shape(native_lib):
    """one economy of 100 000 persons + 5 000 firms: beyond shared memory and beyond 16-bit ids (global-memory
    variant, per-element draws for n > 46340: urngrange / n < n)"""
    from fastace_b200.env import BatchedEconomy
    import torch
    dims = (1, 100000, 5000, 8, 10)
    env = BatchedEconomy(dims)
    host = scenario.OrderStream(dims, 7)
    pp, pf = env.shuffle_orders(seed=7, restart=True, steps=3)
    torch.cuda.synchronize()
    for t in range(3):
        hp, hf = host.next()
        assert np.array_equal(pp[t].cpu().numpy(), hp) and np.array_equal(pf[t].cpu().numpy(), hf), t
    env.close()
