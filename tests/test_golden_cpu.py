"""The oracle against the committed golden vectors (tests/golden/pimc_golden.npz, made by tests/golden/make_golden.py from
pure-Python restatements of the Julia source)."""
import ctypes as C
import os
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pimc_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_teleport_distance_golden(oracle, gold):
    ob = oracle
    for row, l in enumerate(gold["tp_L"]):
        for i, (x, y) in enumerate(zip(gold["tp_x"][row], gold["tp_y"][row])):
            assert ob.lib().ora_teleport(x, l) == gold["tp_teleport"][row, i]
            assert ob.lib().ora_distance(x, y, l) == gold["tp_distance"][row, i]


def test_levy_golden(oracle, gold):
    ob = oracle
    for i, (rows, dim, l, lam, tau) in enumerate(gold["levy_cases"]):
        rows, dim = int(rows), int(dim)
        cm = np.ascontiguousarray(gold[f"levy{i}_r"].T)
        ob.lib().ora_levy(ob._p(cm), rows, dim, tau, l, lam, ob._p(np.ascontiguousarray(gold[f"levy{i}_xi"])))
        assert np.array_equal(cm.T, gold[f"levy{i}_out"]), i      # bit for bit


def test_energy_density_lattice_golden(oracle, gold):
    ob = oracle
    l, T, lam = gold["en_par"]
    r, nxt = gold["en_r"], gold["en_next"]
    N, dim, M = r.shape
    for pot, key in ((ob.make_potential("harmonic", "identity"), "en_harmonic"),
                     (ob.make_potential("lattice", "zero", depth=6.0, scale=1.0, sgn=-1.0, angles=list(gold["lat_angles"])), "en_lattice")):
        s = ob.System(pot, dim=dim, M=M, N=N, L=l, T=T, lam=lam, seed=3)
        s.set_paths(r, nxt)
        E, Ev, _ = s.energy_now()
        scale = dim * N / (2 * s.tau)             # the two leading terms of E cancel at this scale
        assert abs(E - gold[key][0]) <= 1e-12 * scale and abs(Ev - gold[key][1]) <= 1e-12 * max(1.0, abs(gold[key][1]))
    for compat, key in ((ob.COMPAT_ALL, "dens_shift"), (0, "dens_fixed")):
        s = ob.System(ob.make_potential("harmonic", "identity"), dim=dim, M=M, N=N, L=l, T=T, lam=lam, seed=3, compat=compat)
        s.set_paths(r, nxt)
        d = ob.Density(s, 10)
        d.measure(s)
        dens, nd, _ = d.read()
        assert np.array_equal(dens, gold[key]) and nd == M
    p = ob.make_potential("lattice", "zero", depth=6.0, scale=1.0, sgn=-1.0, angles=list(gold["lat_angles"]))
    for pt, v in zip(gold["lat_pts"], gold["lat_V"]):
        assert ob.lib().ora_potential_eval(C.byref(p), ob._p(np.array(pt)), 2) == pytest.approx(v, rel=1e-12, abs=1e-13)
