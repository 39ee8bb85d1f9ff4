"""Known-answer tests for the function plugins on the path (CES::f, CobbDouglas::f,
CES constructor normalisation).  The expected bit patterns were produced by the
UNMODIFIED reference classes (src/functions/vecToScalar.cpp:45-47, 105-118) through
oracle/_ref (fastace_ref_ces_f / fastace_ref_cobb_douglas_f); when oracle/_ref is present
they are re-derived live as well."""
import ctypes as C

import numpy as np
import pytest

# (tfp, raw shares, elasticity of substitution, inputs, reference result as float.hex())
CES_KAT = [
    (1.0, (0.5, 0.5, 0.5), 1.3, (1.0, 5.0, 5.0), "0x1.62f24f4048ba8p+0"),      # SimpleScenario person A
    (1.0, (0.2, 0.6, 0.4), 1.3, (0.5, 2.5, 7.5), "0x1.b45a85475b420p-1"),      # SimpleScenario person B
    (0.5, (1.0, 0.0, 1.0), 3.0, (1.5, 4.0, 6.0), "0x1.55555571f7693p+0"),      # SimpleScenario firm, good 0
    (1.0, (1.0, 0.0, 1.0), 5.0, (0.0, 0.0, 0.0), "0x1.5798ee2308c3ap-27"),     # all-zero inputs: only eps
    (1.0, (0.4, 0.4, 0.1), 10.0, (1.0, 3.3, 0.25), "0x1.6674f1e447828p+0"),    # CustomScenario means
    (1.3, (0.41, 0.09, 0.38), 7.7, (9.5, 2.0, 11.0), "0x1.5ffdd72238b7dp+3"),
    (2.0, (0.3, 0.2, 0.1, 0.25, 0.15), 0.5, (1, 2, 3, 4, 5), "0x1.8fae0c2896f28p+2"),
]
CD_KAT = [
    (1.0, (0.5, 0.5), (4.0, 9.0), "0x1.8000000000000p+2"),
    (2.0, (0.3, 0.3, 0.4), (1.5, 2.5, 3.5), "0x1.3a15a4c88c403p+2"),
]


@pytest.mark.parametrize("tfp,share,el,x,want", CES_KAT)
def test_ces_f_matches_reference_bits(oracle, tfp, share, el, x, want):
    s, rho = oracle.ces_params(share, el)
    got = oracle.ces_f(tfp, s, rho, x)
    assert got.hex() == want


@pytest.mark.parametrize("tfp,el,x,want", CD_KAT)
def test_cobb_douglas_f_matches_reference_bits(oracle, tfp, el, x, want):
    assert oracle.cobb_douglas_f(tfp, el, x).hex() == want


def test_ces_params_rule(oracle):
    # shares normalised to 1, rho = 1/(1 - sigma)  (vecToScalar.cpp:105-110)
    s, rho = oracle.ces_params((0.5, 0.5, 0.5), 1.3)
    assert np.array_equal(s, np.array([0.5, 0.5, 0.5]) / 1.5)
    assert rho == 1 / (1 - 1.3)


def test_live_against_reference_when_present(oracle):
    loader = pytest.importorskip("oracle.loader")
    if not loader.have_reference():
        pytest.skip("oracle/_ref not built here")
    L = loader.ref_lib()
    rng = np.random.default_rng(0)
    dp = C.POINTER(C.c_double)
    for _ in range(300):
        n = int(rng.integers(2, 10))
        share = rng.uniform(0.05, 1.0, n)
        el = float(rng.uniform(0.2, 15.0))
        if abs(el - 1.0) < 1e-3:
            el = 2.0
        x = rng.uniform(0.0, 30.0, n) * (rng.random(n) > 0.1)
        tfp = float(rng.uniform(0.1, 3.0))
        want = L.fastace_ref_ces_f(tfp, share.ctypes.data_as(dp), el, x.ctypes.data_as(dp), n)
        s, rho = oracle.ces_params(share, el)
        s_ref = np.zeros(n)
        rho_ref = C.c_double()
        L.fastace_ref_ces_params(share.ctypes.data_as(dp), el, n, s_ref.ctypes.data_as(dp), C.byref(rho_ref))
        assert np.array_equal(s, s_ref) and rho == rho_ref.value
        assert oracle.ces_f(tfp, s, rho, x).hex() == want.hex()
        e = rng.uniform(0.05, 1.0, n)
        xx = rng.uniform(0.1, 30.0, n)
        assert oracle.cobb_douglas_f(tfp, e, xx).hex() == L.fastace_ref_cobb_douglas_f(
            tfp, e.ctypes.data_as(dp), xx.ctypes.data_as(dp), n).hex()


def test_double_to_int_is_x86_cvttsd2si(oracle):
    f = oracle.lib.fastace_oracle_double_to_int
    assert f(3.99) == 3 and f(-3.99) == -3 and f(0.0) == 0
    assert f(2147483647.5) == 2147483647
    assert f(2147483648.0) == -2147483648 and f(1e300) == -2147483648
    assert f(float("inf")) == -2147483648 and f(float("nan")) == -2147483648
