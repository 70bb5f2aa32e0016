"""Host-side construction of the pair-propagator term table -- the input `tab` of the C ABI.

Reference: `prop_rel_interpolate_terms` / `build_prop_int` (src/propagator.jl:34-86) tabulate, on a 600 x 600 grid over
[1e-20, L]^2,

    A(r1, r2) = -T1 - T2 + T3,    T_i = (1/2pi) * Int_0^inf k exp(-tau k^2) w_i(k) B_i(k r1, k r2) dk
    w_1 = w_3 = t^2/(1+t^2),  w_2 = t/(1+t^2),  t(k) = 1 / ((2/pi)(gamma + ln(k/2)) - 4/g0)
    B_1 = J0 J0,  B_2 = J0(kr1) Y0(kr2) + J0(kr2) Y0(kr1),  B_3 = Y0 Y0

with 3 x 360 000 adaptive Gauss-Kronrod quadratures (QuadGK, rtol 1e-11).  Every integrand is a product
f(k) * u(k r1) * v(k r2), so on a FIXED quadrature rule {k_q, w_q} the whole table is three small matrix products
A = -J' W1 J - (J' W2 Y + Y' W2 J) + Y' W1 Y with J = J0(k_q r_i), Y = Y0(k_q r_i) -- seconds instead of hours.
The rule: 16-point Gauss-Legendre panels, uniform of width <= 10/(2L) (phase advance <= 10 rad per panel for the fastest
oscillation cos(k (r1 + r2))), dyadically graded towards k = 0 where ln k makes the integrand non-analytic, cut at
exp(-tau k^2) < 1e-19.  With D = 1/t the weights are 1/(1+D^2) and D/(1+D^2): smooth through the pole of t(k).
`tests/test_propint_cpu.py` checks entries against scipy's adaptive quad.  (Parity with QuadGK itself is unpinned: Julia
is not available; both converge to the same integrals.)

`determine_nnrange` (src/system.jl:10-15) is the cut-off radius r_a: minimise p(r) - 0.999 on the diagonal r1 = r2 = (r,), then
bisect for its zero between the minimiser and L.
"""
import math
import numpy as np

EULER_GAMMA = 0.5772156649015329
DELTA = 600            # propagator.jl:35
R_LO = 1e-20           # propagator.jl:38-39


def _gl_rule(tau, rmax, order=16, levels=48, phase=10.0):
    """nodes and weights of the composite rule on (0, kmax)"""
    kmax = math.sqrt(44.0 / tau)
    h = min(phase / (2.0 * rmax), kmax / 8.0)
    x, w = np.polynomial.legendre.leggauss(order)
    edges = [0.0] + [h * 2.0 ** (-j) for j in range(levels, 0, -1)]
    npan = int(math.ceil((kmax - h) / h))
    edges += list(h + (kmax - h) * np.arange(0, npan + 1) / npan)[0:]
    edges = np.unique(np.array(edges))
    a, b = edges[:-1], edges[1:]
    k = (0.5 * (b - a))[:, None] * x[None, :] + (0.5 * (b + a))[:, None]
    wk = (0.5 * (b - a))[:, None] * w[None, :]
    return k.ravel(), wk.ravel()


def _weights(k, wk, g0, tau):
    D = (2.0 / math.pi) * (EULER_GAMMA + np.log(k / 2.0)) - 4.0 / g0       # 1 / tk(k)
    base = wk * k * np.exp(-tau * k * k) / (2.0 * math.pi)
    return base / (1.0 + D * D), base * D / (1.0 + D * D)                 # t^2/(1+t^2), t/(1+t^2)


def prop_rel_terms(r1, r2, g0, tau):
    """A(r1_i, r2_j) for arbitrary radius vectors (same fixed rule as the table)"""
    from scipy.special import j0, y0
    r1, r2 = np.atleast_1d(np.asarray(r1, dtype=np.float64)), np.atleast_1d(np.asarray(r2, dtype=np.float64))
    k, wk = _gl_rule(tau, max(r1.max(), r2.max(), 1e-3))
    w1, w2 = _weights(k, wk, g0, tau)
    J1, Y1 = j0(k[:, None] * r1[None, :]), y0(k[:, None] * r1[None, :])
    J2, Y2 = (J1, Y1) if r2 is r1 else (j0(k[:, None] * r2[None, :]), y0(k[:, None] * r2[None, :]))
    t1 = (J1 * w1[:, None]).T @ J2
    t2 = (J1 * w2[:, None]).T @ Y2 + (Y1 * w2[:, None]).T @ J2
    t3 = (Y1 * w1[:, None]).T @ Y2
    return -t1 - t2 + t3


def prop_rel_interpolate_terms(L, g0, tau, delta=DELTA):
    """the sampled table of propagator.jl:34-70: dict(tab=(delta, delta) array [i over r1, j over r2], lo, hi)"""
    r = np.linspace(R_LO, L, delta)
    A = prop_rel_terms(r, r, g0, tau)
    return dict(tab=np.asfortranarray(0.5 * (A + A.T)), lo=R_LO, hi=float(L), g0=float(g0), tau=float(tau))


def build_prop_int(L, g0, tau, delta=DELTA):
    """build_prop_int(L, g0, tau) (propagator.jl:79-89): returns the `propint` argument of System(...; interactions=true)"""
    return prop_rel_interpolate_terms(L, g0, tau, delta)


def terms_lookup(p, x, y):
    """bilinear interpolation `terms(r1_norm, r2_norm)` (Interpolations.jl scale(interpolate(A, BSpline(Linear())), ...))"""
    A, n = p["tab"], p["tab"].shape[0]
    h = (p["hi"] - p["lo"]) / (n - 1)
    tx, ty = (x - p["lo"]) / h, (y - p["lo"]) / h
    ix, iy = min(max(int(math.floor(tx)), 0), n - 2), min(max(int(math.floor(ty)), 0), n - 2)
    fx, fy = tx - ix, ty - iy
    c0 = (1 - fx) * A[ix, iy] + fx * A[ix + 1, iy]
    c1 = (1 - fx) * A[ix, iy + 1] + fx * A[ix + 1, iy + 1]
    return (1 - fy) * c0 + fy * c1


def prop_int(p, r1_rel, r2_rel, tau):
    """prop_int(r1_rel, r2_rel, tau) = 1 + terms(|r1|, |r2|) / prop_rel0(r1, r2, tau) (propagator.jl:73-86)"""
    r1, r2 = np.atleast_1d(np.asarray(r1_rel, dtype=np.float64)), np.atleast_1d(np.asarray(r2_rel, dtype=np.float64))
    d = r1 - r2
    rel0 = math.exp(-float(d @ d) / (4 * tau)) / (4 * math.pi * tau)
    return 1 + terms_lookup(p, float(np.linalg.norm(r1)), float(np.linalg.norm(r2))) / rel0


def determine_nnrange(p, tau, a, b):
    """determine_nnrange(propint, tau, a, b) (system.jl:10-15): zero of p([r],[r]) - 0.999 right of its minimiser"""
    f = lambda r: prop_int(p, [r], [r], tau) - 0.999
    # minimiser: coarse scan + golden section (Optim's Nelder-Mead from x0 = a lands in the same basin: f has one minimum)
    xs = np.linspace(max(a, p["lo"]), b, 2001)
    fs = np.array([f(x) for x in xs])
    i = int(np.argmin(fs))
    lo, hi = xs[max(i - 1, 0)], xs[min(i + 1, len(xs) - 1)]
    gr = (math.sqrt(5) - 1) / 2
    for _ in range(80):
        c, d = hi - gr * (hi - lo), lo + gr * (hi - lo)
        if f(c) < f(d):
            hi = d
        else:
            lo = c
    rmin = 0.5 * (lo + hi)
    fa, fb = f(rmin), f(b)
    if not (fa < 0.0 <= fb or fb < 0.0 <= fa):
        raise ValueError("determine_nnrange: no sign change of propint - 0.999 on (r_min, b) (Roots.find_zero would throw)")
    lo, hi = rmin, b
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if (f(mid) < 0.0) == (fa < 0.0):
            lo = mid
        else:
            hi = mid
        if hi - lo <= 1e-15 * max(1.0, abs(hi)):
            break
    return 0.5 * (lo + hi)
