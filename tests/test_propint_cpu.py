"""Pair-propagator term table built on the host (pimc_jl_b200/propint.py) against scipy's adaptive quadrature of the integrals of
src/propagator.jl:34-70 (the reference evaluates them with QuadGK, rtol 1e-11), and determine_nnrange (src/system.jl:10-15)."""
import math
import numpy as np
import pytest
from pimc_jl_b200 import propint as P


def quad_terms(r1, r2, g0, tau):
    from scipy.integrate import quad
    from scipy.special import j0, y0

    def D(k):
        return (2 / math.pi) * (P.EULER_GAMMA + math.log(k / 2)) - 4 / g0
    fs = ((lambda k: k * math.exp(-tau * k * k) / (1 + D(k) ** 2) * j0(k * r1) * j0(k * r2), -1.0),
          (lambda k: k * math.exp(-tau * k * k) * D(k) / (1 + D(k) ** 2) * (j0(k * r1) * y0(k * r2) + j0(k * r2) * y0(k * r1)), -1.0),
          (lambda k: k * math.exp(-tau * k * k) / (1 + D(k) ** 2) * y0(k * r1) * y0(k * r2), 1.0))
    edges = np.linspace(0.0, math.sqrt(44 / tau), 120)
    tot = 0.0
    for f, s in fs:
        tot += s * sum(quad(f, a, b, epsabs=1e-15, epsrel=1e-13, limit=400)[0] for a, b in zip(edges[:-1], edges[1:])) / (2 * math.pi)
    return tot


@pytest.mark.parametrize("g0,tau,L", [(2.0, 1 / (0.2 * 256), 12.0), (0.7, 0.05, 6.0)])
def test_table_entries_against_adaptive_quadrature(g0, tau, L):
    p = P.build_prop_int(L, g0, tau, delta=120)
    r = np.linspace(P.R_LO, L, 120)
    scale = np.abs(p["tab"]).max()
    for i, j in [(0, 0), (0, 3), (1, 1), (2, 5), (7, 7), (11, 13), (30, 31), (119, 119)]:
        ref = quad_terms(r[i], r[j], g0, tau)
        assert abs(p["tab"][i, j] - ref) <= 1e-11 * abs(ref) + 1e-15 * scale, (i, j, p["tab"][i, j], ref)
    assert np.array_equal(p["tab"], p["tab"].T)


def test_prop_int_limits_and_nnrange():
    g0, tau, L = 2.0, 1 / (0.2 * 256), 8.0
    p = P.build_prop_int(math.ceil(math.sqrt(2) * L), g0, tau)      # examples/density_SRL_lattice.jl:17
    assert p["tab"].shape == (600, 600) and p["lo"] == 1e-20
    # far apart the pair propagator tends to the free one; inside the core it is suppressed
    assert abs(P.prop_int(p, [3.0, 0.0], [3.0, 0.1], tau) - 1.0) < 1e-6
    assert P.prop_int(p, [0.05, 0.0], [0.05, 0.0], tau) < 0.9
    ra = P.determine_nnrange(p, tau, 1e-20, L)
    assert 0.0 < ra < L and abs(P.prop_int(p, [ra], [ra], tau) - 0.999) < 1e-9
    # bilinear lookup reproduces the nodes
    r = np.linspace(1e-20, p["hi"], 600)
    assert P.terms_lookup(p, r[17], r[40]) == pytest.approx(p["tab"][17, 40], rel=1e-13)


# ---- the library's own host-side construction (csrc/pimc_propint.cu: what the Julia shim binds) ----
def _lib_table(L, g0, tau, delta):
    import ctypes as C
    from pimc_jl_b200 import _lib
    lib = _lib.load()
    tab = np.zeros((delta, delta), order="F")
    lo, hi = C.c_double(), C.c_double()
    assert lib.pimc_build_prop_table(L, g0, tau, delta, tab.ctypes.data_as(_lib.f64p), C.byref(lo), C.byref(hi)) == 0
    return dict(tab=tab, lo=lo.value, hi=hi.value)


@pytest.mark.parametrize("g0,tau,L", [(2.0, 1 / (0.2 * 256), 12.0), (0.7, 0.05, 6.0)])
def test_library_table_against_adaptive_quadrature_and_numpy(g0, tau, L):
    """pimc_build_prop_table (C++ host code, glibc j0/y0) against adaptive quadrature (1e-11, the reference's QuadGK tolerance) and against
    the independent numpy construction (scipy j0/y0) entry by entry"""
    q = _lib_table(L, g0, tau, 120)
    p = P.build_prop_int(L, g0, tau, delta=120)
    assert q["lo"] == 1e-20 and q["hi"] == L
    scale = np.abs(p["tab"]).max()
    assert np.all(np.abs(q["tab"] - p["tab"]) <= 2e-12 * np.abs(p["tab"]) + 1e-15 * scale)
    assert np.array_equal(q["tab"], q["tab"].T)
    r = np.linspace(P.R_LO, L, 120)
    for i, j in [(0, 0), (0, 3), (2, 5), (11, 13), (119, 119)]:
        ref = quad_terms(r[i], r[j], g0, tau)
        assert abs(q["tab"][i, j] - ref) <= 1e-11 * abs(ref) + 1e-15 * scale, (i, j, q["tab"][i, j], ref)


def test_library_prop_int_and_nnrange():
    """pimc_prop_int / pimc_determine_nnrange at the shape of examples/density_SRL_lattice.jl:17 against the numpy versions"""
    import ctypes as C
    from pimc_jl_b200 import _lib
    lib = _lib.load()
    g0, tau, L = 2.0, 1 / (0.2 * 256), 8.0
    q = _lib_table(math.ceil(math.sqrt(2) * L), g0, tau, 600)
    tp = q["tab"].ctypes.data_as(_lib.f64p)
    out = C.c_double()
    rng = np.random.default_rng(3)
    for _ in range(50):
        r1, r2 = rng.uniform(-2, 2, 2), rng.uniform(-2, 2, 2)
        assert lib.pimc_prop_int(tp, 600, q["lo"], q["hi"], r1.ctypes.data_as(_lib.f64p), r2.ctypes.data_as(_lib.f64p), 2, tau, C.byref(out)) == 0
        ref = P.prop_int(q, r1, r2, tau)
        assert abs(out.value - ref) <= 1e-12 * max(1.0, abs(ref))   # exp(d^2 / 4 tau) amplifies the last bit of d^2
    ra = C.c_double()
    assert lib.pimc_determine_nnrange(tp, 600, q["lo"], q["hi"], tau, 1e-20, L, C.byref(ra)) == 0
    assert abs(ra.value - P.determine_nnrange(q, tau, 1e-20, L)) <= 1e-12
    x = np.array([ra.value])
    lib.pimc_prop_int(tp, 600, q["lo"], q["hi"], x.ctypes.data_as(_lib.f64p), x.ctypes.data_as(_lib.f64p), 1, tau, C.byref(out))
    assert abs(out.value - 0.999) < 1e-9
    # no sign change on the bracket: Roots.find_zero would throw (system.jl:14)
    flat = np.zeros((4, 4), order="F")
    assert lib.pimc_determine_nnrange(flat.ctypes.data_as(_lib.f64p), 4, 0.0, 1.0, tau, 1e-20, 1.0, C.byref(ra)) != 0
