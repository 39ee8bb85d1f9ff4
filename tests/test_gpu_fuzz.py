"""Randomised shapes and action recipes through all three step paths (lane-parallel warp kernel, serial warp kernel,
large-economy path) against the oracle.  compute-sanitizer is closed on this GPU pool, so breadth of shapes is what
stands in for it: odd sizes around every tiling boundary (32-person windows, 4-person staging vectors, F*G near the
254-entry book limit, S around the 12/16 unroll variants)."""
import numpy as np
import pytest

from fastace_b200 import _abi, scenario
from tests.test_gpu_parity import _run_episode

pytestmark = pytest.mark.gpu


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        G = int(rng.integers(1, 9))
        F = int(rng.integers(1, 254 // G + 1))
        if rng.random() < 0.5:
            F = min(F, int(rng.integers(1, 13)))
        P = int(rng.choice([0, 1, 2, 3, 31, 32, 33, 63, 64, 65, 96, 100, 127, 128, 129, 200]))
        S = int(rng.choice([0, 1, 3, 4, 5, 8, 10, 11, 12, 13, 15, 16]))
        E = int(rng.integers(1, 6))
        preset = dict(take_prob=float(rng.choice([0.1, 0.5, 0.9, 1.0])), labor_mu=float(rng.choice([0.5, 1.0, 2.3])),
                      wage_scale=float(rng.choice([0.05, 0.1, 1.0])), price_scale=float(rng.choice([0.5, 1.5, 5.0])),
                      prod_scale=float(rng.choice([0.2, 0.5, 1.0])))
        out.append(((E, P, F, G, S), preset, int(rng.integers(0, 1 << 30))))
    return out


@pytest.mark.parametrize("mode", [0, _abi.STEP_SERIAL, _abi.STEP_LARGE], ids=["parallel", "serial", "large"])
def test_random_shapes_and_recipes(oracle, mode):
    for dims, preset, seed in _cases(20, seed=2024 + mode):
        idx = _abi.IDX_MODULO if seed % 3 else _abi.IDX_ABSOLUTE
        try:
            _run_episode(oracle, dims, 6, seed=seed % 100000, preset=preset, flags=idx | mode)
        except AssertionError as err:
            raise AssertionError(f"dims={dims} preset={preset} seed={seed} mode={mode}: {err}") from err
