#!/usr/bin/env python
"""Generates tests/golden/pimc_golden.npz -- golden input/output vectors for the hot path.

The reference (pure Julia, unseeded RNG) cannot be executed in this image and ships no golden vectors, so the vectors are made
by the PURE-PYTHON restatements in tests/test_oracle_cpu.py (written from the Julia source, independent of the C oracle and of
the CUDA code): levy! with supplied Gaussians (src/updates/helper.jl:118-139), teleport / distance (src/propagator.jl:6-32),
Energy and Density (src/measurement.jl:45-122), lattice potential (examples/tools/potentialtools.jl:1-39).  Deterministic
(numpy default_rng seeds).  Consumers: tests/test_golden_cpu.py (oracle) and tests/test_gpu_parity.py::test_golden_vectors (CUDA).

    python tests/golden/make_golden.py
"""
import math
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from test_oracle_cpu import teleport_py, distance_py, levy_py, energy_py  # noqa: E402


def lattice_py(r, angles, scale, depth, sgn):     # potentialtools.jl:1-16,38-39: sgn * depth * normalized_intensity(coord, ang, scale)
    s = c = 0.0
    for a in angles:
        rr = r[0] * math.sin(a) + r[1] * math.cos(a)
        s += math.sin(2 * math.pi * rr * scale)
        c += math.cos(2 * math.pi * rr * scale)
    s /= len(angles)
    c /= len(angles)
    return sgn * depth * (s * s + c * c)


def density_py(r, L, nbins, shift=True):          # measurement.jl:45-55 (shift = as shipped: floor-bin 0 dropped)
    N, dim, M = r.shape
    binw = (2 * L) / nbins
    dens = np.zeros((nbins,) * dim)
    for n in range(N):
        for m in range(M):
            ib = np.floor((r[n, :, m] + L) / binw).astype(int)
            if shift:
                if np.all(ib > 0) and np.all(ib < nbins + 1):
                    dens[tuple(ib - 1)] += 1
            elif np.all(ib >= 0) and np.all(ib < nbins):
                dens[tuple(ib)] += 1
    return dens


def main():
    rng = np.random.default_rng(20261017)
    g = {}
    # --- teleport / distance ---
    L = np.array([4.0, 100.0, 0.37, 16.0])
    x = np.concatenate([rng.uniform(-5, 5, (4, 60)) * L[:, None], np.array([[1, -1, 0, 3, -3, 1e-300, -1e-17]]) * L[:, None]], axis=1)
    y = rng.uniform(-1, 1, x.shape) * L[:, None]
    g["tp_L"], g["tp_x"], g["tp_y"] = L, x, y
    g["tp_teleport"] = np.array([[teleport_py(v, l) for v in row] for row, l in zip(x, L)])
    g["tp_distance"] = np.array([[distance_py(a, b, l) for a, b in zip(ra, rb)] for ra, rb, l in zip(x, y, L)])
    # --- levy! with supplied Gaussians ---
    cases = [(3, 2, 4.0, 1.0, 0.01), (12, 2, 100.0, 0.5, 0.2), (100, 2, 4.0, 1.0, 0.01), (6, 2, 0.5, 2.0, 0.3), (21, 2, 16.0, 1.0, 1 / 128), (9, 1, 4.0, 1.0, 0.01)]
    g["levy_cases"] = np.array(cases, dtype=np.float64)
    for i, (rows, dim, l, lam, tau) in enumerate(cases):
        r = np.zeros((rows, dim))
        r[0], r[-1] = rng.uniform(-l, l, dim), rng.uniform(-l, l, dim)
        if i % 2 == 0:
            r[0, 0], r[-1, 0] = 0.95 * l, -0.95 * l      # endpoints across the periodic boundary (helper.jl:120-125)
        xi = rng.standard_normal((rows - 2, dim))
        g[f"levy{i}_r"], g[f"levy{i}_xi"], g[f"levy{i}_out"] = r, xi, levy_py(r, tau, l, lam, xi)
    # --- Energy / Density on fixed worldlines, harmonic trap and l25 lattice ---
    N, dim, M, l, T, lam = 5, 2, 9, 3.0, 0.8, 0.5
    tau = (1 / T) / M
    r = rng.uniform(-l, l, (N, dim, M))
    nxt = np.array([2, 3, 1, 4, 5], dtype=np.int64)      # a 3-cycle and two identity cycles
    g["en_r"], g["en_next"], g["en_par"] = r, nxt, np.array([l, T, lam])
    g["en_harmonic"] = np.array(energy_py(r, nxt, l, tau, lam, lambda q: 0.5 * (q[0] ** 2 + q[1] ** 2), lambda q: q))
    ang = [2.214297435588181, 0.9272952180016122, -0.6435011087932844, 0.6435011087932844, -2.498091544796509, 3.141592653589793,
           2.498091544796509, 0, 1.5707963267948966, -2.2142974355881813, -1.5707963267948968, -0.9272952180016123]
    g["lat_angles"] = np.array(ang)
    g["en_lattice"] = np.array(energy_py(r, nxt, l, tau, lam, lambda q: lattice_py(q, ang, 1.0, 6.0, -1.0), lambda q: 0 * q))
    pts = rng.uniform(-l, l, (40, 2))
    g["lat_pts"], g["lat_V"] = pts, np.array([lattice_py(p, ang, 1.0, 6.0, -1.0) for p in pts])
    g["dens_shift"], g["dens_fixed"] = density_py(r, l, 10, True), density_py(r, l, 10, False)
    # --- closed-form finite-M harmonic energies (formula in tests/test_oracle_cpu.py::test_closed_form_constants) ---
    g["exact_continuum_T1"] = np.array(2.1639534137386534)      # examples/energy_2d_harmonically_trapped_bose_gas.jl:28
    np.savez_compressed(os.path.join(HERE, "pimc_golden.npz"), **g)
    print("wrote", os.path.join(HERE, "pimc_golden.npz"), len(g), "arrays")


if __name__ == "__main__":
    main()
