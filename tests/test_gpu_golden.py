"""The CUDA path (through the C ABI, host-pointer entry point) against the golden vectors
generated from the UNMODIFIED reference.  Integer / index / counter state bit-exact; floating
point within 1e-5 relative (the device pow differs from glibc's in the last ulp)."""
import numpy as np
import pytest

from fastace_b200 import _abi
from tests import golden_util as GU
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GU.NAMES)
def test_cuda_reproduces_reference(native_lib, name):
    from fastace_b200.env import BatchedEconomy
    g = GU.Golden(name)
    env = BatchedEconomy(g.dims)
    env.set_state(g.initial_state())
    before = g.initial_state()
    for t in range(g.steps):
        out = _abi.alloc_host("out", g.dims)
        env.time_step_host(g.actions(t), out, flags=g.flags)
        want_out = g.outputs(t)
        got = env.get_state()
        want = dict(g.state(t))
        H.compare_outputs({k: out[k] for k in want_out}, want_out, g.dims, before)
        for k in ("p_util_tfp", "p_util_share", "p_util_rho", "f_prod_tfp", "f_prod_share", "f_prod_rho"):
            want[k] = before[k]
        H.compare_states(got, want, g.dims)
        before = want
    env.close()
