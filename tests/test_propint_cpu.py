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
