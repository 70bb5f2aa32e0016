"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for bridges / positions / link cache / counters; <= 1e-12 relative for reduced sums (order differs)."""
import ctypes as C
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import pimc_jl_b200 as pj
from pimc_jl_b200 import engine as eng
from pimc_jl_b200 import _lib as L

RTOL = 1e-12
L25 = [2.214297435588181, 0.9272952180016122, -0.6435011087932844, 0.6435011087932844, -2.498091544796509, 3.141592653589793,
       2.498091544796509, 0, 1.5707963267948966, -2.2142974355881813, -1.5707963267948968, -0.9272952180016123]


def pots(ob):
    """matching (oracle potential, engine potential) pairs"""
    out = {}
    for name, kw in {
        "zero": dict(kind="zero", dv="identity"),
        "harmonic": dict(kind="harmonic", dv="identity"),
        "sin2": dict(kind="sin2_1d", dv="zero", depth=8.0, scale=0.5),
        "lattice": dict(kind="lattice", dv="zero", depth=6.0, scale=1.0, sgn=-1.0, angles=L25),
        "lattice_grad": dict(kind="lattice", dv="gradient", depth=6.0, scale=1.0, sgn=-1.0, angles=L25, helical=True),
    }.items():
        out[name] = (ob.make_potential(**kw), pj.make_potential(**kw))
    return out


def close(a, b, rtol=RTOL, atol=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol)


def test_primitives_bit_exact(oracle):
    ob = oracle
    rng = np.random.default_rng(1)
    for Lbox in (4.0, 100.0, 16.0, 0.7):
        x1 = rng.uniform(-3 * Lbox, 3 * Lbox, 4000)
        x2 = rng.uniform(-3 * Lbox, 3 * Lbox, 4000)
        x1[:6] = [Lbox, -Lbox, 0.0, -0.0, 1e-20, 3 * Lbox]
        ref_d = np.array([ob.lib().ora_distance(a, b, Lbox) for a, b in zip(x1, x2)])
        ref_t = np.array([ob.lib().ora_teleport(a, Lbox) for a in x1])
        assert np.array_equal(eng.distance(x1, x2, Lbox), ref_d)
        assert np.array_equal(eng.teleport(x1, Lbox), ref_t)
        t = eng.teleport(x1, Lbox)
        assert np.all(t >= -Lbox) and np.all(t <= Lbox)
    r1, r2 = rng.uniform(-4, 4, (500, 2)), rng.uniform(-4, 4, (500, 2))
    ref = np.array([ob.lib().ora_lnK(ob._p(a.copy()), ob._p(b.copy()), 2, 0.01, 0.5, 4.0) for a, b in zip(r1, r2)])
    assert np.array_equal(eng.lnK(r1, r2, 0.01, 0.5, 4.0), ref)


def test_potentials(oracle):
    ob = oracle
    rng = np.random.default_rng(2)
    r = rng.uniform(-4, 4, (800, 2))
    for name, (po, pg) in pots(ob).items():
        refV = np.array([ob.lib().ora_potential_eval(C.byref(po), ob._p(x.copy()), 2) for x in r])
        refdV = np.zeros_like(r)
        for i, x in enumerate(r):
            g = np.zeros(2)
            ob.lib().ora_potential_grad(C.byref(po), ob._p(x.copy()), 2, ob._p(g))
            refdV[i] = g
        V, dV = eng.potential_eval(r, pg)
        if name in ("zero", "harmonic"):
            assert np.array_equal(V, refV), name
            assert np.array_equal(dV, refdV), name
        else:
            assert close(V, refV, 1e-12, 1e-13), name
            assert close(dV, refdV, 1e-11, 1e-11), name
        ref_lnV = np.array([ob.lib().ora_lnV(ob._p(a.copy()), ob._p(b.copy()), 2, 0.02, C.byref(po)) for a, b in zip(r[:-1], r[1:])])
        assert close(eng.lnV(r[:-1], r[1:], 0.02, pg), ref_lnV, 1e-12, 1e-13), name


def test_rng_stream_bit_exact(oracle):
    ob = oracle
    g = eng.gauss_pairs(0x5EEDB200, 7, 12345678901, 3, 2, 5, 1, 2000)
    ref = np.zeros_like(g)
    a, b = C.c_double(), C.c_double()
    for i in range(2000):
        ob.lib().ora_gauss_pair(0x5EEDB200, 7, 12345678901, 3, 2, 5, 1 + i, C.byref(a), C.byref(b))
        ref[i] = (a.value, b.value)
    assert np.array_equal(g, ref)
    big = eng.gauss_pairs(1, 0, 0, 0, 2, 0, 0, 16000).ravel()
    assert abs(big.mean()) < 0.03 and abs(big.var() - 1) < 0.03


@pytest.mark.parametrize("dim", [1, 2])
def test_levy_bridge_bit_exact(oracle, dim):
    """Given the same Gaussian draws levy! is reproduced bit for bit (north-star check 2)."""
    ob = oracle
    rng = np.random.default_rng(3)
    for rows, Lbox, lam, tau in [(3, 4.0, 1.0, 0.01), (7, 100.0, 0.5, 0.2), (22, 4.0, 1.0, 0.01), (130, 16.0, 1.0, 1.0 / 128), (5, 0.5, 1.0, 0.3)]:
        nb = 64
        r = np.zeros((nb, dim, rows))
        r[:, :, 0] = rng.uniform(-Lbox, Lbox, (nb, dim))
        r[:, :, -1] = rng.uniform(-Lbox, Lbox, (nb, dim))
        r[0, :, 0], r[0, :, -1] = 0.9 * Lbox, -0.9 * Lbox  # forces the boundary shift
        r[1, :, 0], r[1, :, -1] = 0.0, -0.9 * Lbox        # sign(0) = 0
        xi = rng.standard_normal((nb, rows - 2, dim))
        out = eng.levy_bridge(r, tau, Lbox, lam, xi)
        for b in range(nb):
            ref = r[b].copy()
            ob.lib().ora_levy(ob._p(ref), rows, dim, tau, Lbox, lam, ob._p(np.ascontiguousarray(xi[b])))
            assert np.array_equal(out[b], ref), (rows, b)


CONFIGS = [
    dict(pot="harmonic", dim=2, M=5, N=1, L=100.0, T=1.0, lam=0.5, Ncycle=10),     # C1 as shipped
    dict(pot="zero", dim=2, M=16, N=6, L=4.0, T=1.0, lam=1.0, Ncycle=2),           # C2-like, small
    dict(pot="sin2", dim=1, M=12, N=3, L=4.0, T=1.0, lam=1.0, Ncycle=3),           # 1-D test system
    dict(pot="lattice", dim=2, M=20, N=4, L=8.0, T=0.2, lam=1.0 / np.pi ** 2, Ncycle=3),  # C4-like
]


def make_pair(ob, cfg, chains=3, seed=11, **extra):
    po, pg = pots(ob)[cfg["pot"]]
    kw = dict(dim=cfg["dim"], M=cfg["M"], N=cfg["N"], T=cfg["T"], lam=cfg["lam"], Ncycle=cfg["Ncycle"], seed=seed)
    kw.update(extra)
    e = pj.Engine(pg, chains=chains, L_=cfg["L"], **kw)
    os_ = [ob.System(po, L=cfg["L"], chain=c, **kw) for c in range(chains)]
    return e, os_


@pytest.mark.parametrize("cfg", CONFIGS)
def test_init_world_and_estimators(oracle, cfg):
    ob = oracle
    e, os_ = make_pair(ob, cfg)
    r, V, bins, nxt = e.paths()
    E, Ev, parts = e.energy_now()
    ac, ar = e.action()
    for c, s in enumerate(os_):
        ro, Vo, bo, no = s.paths()
        assert np.array_equal(r[c], ro)
        exact = cfg["pot"] in ("zero", "harmonic")
        assert np.array_equal(V[c], Vo) if exact else close(V[c], Vo, 1e-12, 1e-14)
        assert np.array_equal(bins[c], bo) and np.array_equal(nxt[c], no)
        Eo, Evo, po_ = s.energy_now()
        # E is a difference of two large terms: compare the parts relatively and E against the scale of its terms
        assert close(parts[c], po_, 1e-12, 1e-13)
        scale = cfg["dim"] * cfg["N"] / (2 * s.tau)
        assert abs(E[c] - Eo) <= 1e-12 * scale and abs(Ev[c] - Evo) <= 1e-12 * max(1.0, abs(Evo))
        assert close(ac[c], ob.lib().ora_action_links(s.h), 1e-12, 1e-13)
        assert close(ar[c], ob.lib().ora_action_links_recomputed(s.h), 1e-12, 1e-13)


def _sync_paths(e, os_, exact_v=True):
    r, V, bins, nxt = e.paths()
    for c, s in enumerate(os_):
        ro, Vo, bo, no = s.paths()
        assert np.array_equal(nxt[c], no), f"chain {c}: permutation differs"
        assert np.array_equal(r[c], ro), f"chain {c}: positions differ"
        assert np.array_equal(bins[c], bo), f"chain {c}: bins differ"
        if exact_v:
            assert np.array_equal(V[c], Vo), f"chain {c}: link cache differs"
        else:
            assert close(V[c], Vo, 1e-12, 1e-14)


@pytest.mark.parametrize("cfg", CONFIGS[:2] + CONFIGS[3:])
def test_explicit_moves_delta_u(oracle, cfg):
    """On identical configurations w_initial / w_updated match the oracle to 1e-12 and the committed paths bit for bit."""
    ob = oracle
    e, os_ = make_pair(ob, cfg, chains=2, seed=5)
    rng = np.random.default_rng(9)
    M, N, dim = cfg["M"], cfg["N"], cfg["dim"]
    exact = cfg["pot"] in ("zero", "harmonic")
    for trial in range(40):
        c = trial % 2
        s = os_[c]
        n = int(rng.integers(1, N + 1)); j0 = int(rng.integers(1, M + 1)); m = int(rng.integers(2, M - 1))
        u = float(rng.uniform())
        kind = trial % 3 if N > 1 else (trial % 2) * 2
        if kind == 0:
            xi = rng.standard_normal((m - 1, dim))
            wi, wu = C.c_double(), C.c_double()
            rp = np.zeros((dim, m + 1))
            acc_o = ob.lib().ora_reshape_linear_explicit(s.h, n, j0, m, ob._p(xi), u, 1, C.byref(wi), C.byref(wu), ob._p(rp))
            acc, gwi, gwu, grp = e.reshape_linear_explicit(c, n, j0, m, xi, u)
            assert acc == acc_o
            assert np.array_equal(grp, rp)
        elif kind == 1:
            n2 = int(rng.integers(1, N + 1))
            if n2 == n:
                n2 = n % N + 1
            xi1, xi2 = rng.standard_normal((m - 1, dim)), rng.standard_normal((m - 1, dim))
            wi, wu = C.c_double(), C.c_double()
            acc_o = ob.lib().ora_reshape_swap_explicit(s.h, n, n2, j0, m, ob._p(xi1), ob._p(xi2), u, 1, C.byref(wi), C.byref(wu))
            acc, gwi, gwu = e.reshape_swap_explicit(c, n, n2, j0, m, xi1, xi2, u)
            assert acc == acc_o
        else:
            d = rng.uniform(-1, 1, 2)
            wi, wu = C.c_double(), C.c_double()
            acc_o = ob.lib().ora_com_explicit(s.h, n, 1, ob._p(d), u, 1, C.byref(wi), C.byref(wu))
            acc, gwi, gwu = e.com_explicit(c, n, d, u, polymer=True)
            assert acc == acc_o
        assert close(gwi, wi.value, 1e-12, 1e-13) and close(gwu, wu.value, 1e-12, 1e-13), (trial, kind, gwi, wi.value, gwu, wu.value)
        _sync_paths(e, os_, exact)


def _mk_updates(ob, e, s_list, spec):
    ge = [(every, e.update_create(kind, v0)) for every, kind, v0 in spec]
    oo = [[(every, ob.Update(s, kind, v0)) for every, kind, v0 in spec] for s in s_list]
    return ge, oo


@pytest.mark.parametrize("sched,impl", [(L.SCHED_FAITHFUL, 0), (L.SCHED_FAITHFUL, -1), (L.SCHED_SWEEP, 1), (L.SCHED_SWEEP, 2)],
                         ids=["faithful", "faithful-one-thread", "sweep-persistent", "sweep-batched"])
@pytest.mark.parametrize("cfg", CONFIGS + [dict(pot="harmonic", dim=2, M=40, N=70, L=6.0, T=0.5, lam=0.5, Ncycle=4)], ids=lambda c: f"{c['pot']}-N{c['N']}-M{c['M']}")
def test_run_trajectory_bit_exact(oracle, cfg, sched, impl):
    """Same seed, same schedule: the GPU Markov chains follow the oracle's chains bit for bit, including the adaptive
    step / slice variables, acceptance windows, and the measured energies / density histograms."""
    ob = oracle
    e, os_ = make_pair(ob, cfg, chains=3, seed=21)
    if impl < 0:
        e.set_option(L.OPT_FAITHFUL_IMPL, 1)  # one thread per proposal (pimc_moves.cuh) instead of the warp-cooperative bodies
    else:
        e.set_option(L.OPT_SWEEP_IMPL, impl)
    spec = [(2, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 6)]
    if cfg["N"] > 1:
        spec += [(1, L.UPD_RESHAPE_SWAP, 5), (3, L.UPD_POLYMER_COM, 0.7)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    exact = cfg["pot"] in ("zero", "harmonic")
    # thermalisation leg (no measurements), then a measured leg; two calls exercise counter persistence
    n1, n2 = (300, 400) if sched == L.SCHED_FAITHFUL else ((60, 90) if cfg["N"] < 50 else (30, 40))
    e.run(n1, ge, sched=sched)
    for s, ups in zip(os_, oo):
        s.run(n1, ups, sched=sched)
    _sync_paths(e, os_, exact)
    en_id, de_id = e.energy_create(1000), e.density_create(40)
    oen = [ob.Energy(1000) for _ in os_]
    ode = [ob.Density(s, 40) for s in os_]
    st = e.run(n2, ge, energies=[en_id], densities=[de_id], sched=sched)
    for s, ups, en, de in zip(os_, oo, oen, ode):
        s.run(n2, ups, energies=[en], densities=[de], sched=sched)
    _sync_paths(e, os_, exact)
    tot_bm = 0
    for (_, uid), k in zip(ge, range(len(spec))):
        for c in range(3):
            g, o = e.update_get(uid, c), oo[c][k][1].get()
            assert g["tries"] == o["tries"] and g["tries_var"] == o["tries_var"] and g["accepted"] == o["accepted"], (k, c, g, o)
            assert g["bead_moves"] == o["bead_moves"] and g["var"] == o["var"], (k, c, g, o)
            assert (np.isnan(g["acc_window"]) and np.isnan(o["acc_window"])) or g["acc_window"] == o["acc_window"]
            tot_bm += o["bead_moves"]
    scale = cfg["dim"] * cfg["N"] / (2 * os_[0].tau)
    dsum = None
    for c in range(3):
        E, Ev, n = e.energy_read(en_id, c)
        Eo, Evo = oen[c].read()
        assert n == len(Eo) == n2 // cfg["Ncycle"]
        assert np.all(np.abs(E - Eo) <= 1e-12 * scale) and np.all(np.abs(Ev - Evo) <= 1e-12 * np.maximum(1.0, np.abs(Evo)))
        d, nd, b = ode[c].read()
        dsum = d if dsum is None else dsum + d
    Em, Evm, n = e.energy_read(en_id, -1)
    assert np.all(np.abs(Em - np.mean([oen[c].read()[0] for c in range(3)], axis=0)) <= 1e-12 * scale)
    dg, ndg, bg = e.density_read(de_id, 40)
    assert np.array_equal(dg, dsum) and ndg == 3 * (n2 // cfg["Ncycle"]) * cfg["M"] and bg == ode[0].read()[2]
    assert st["measurements"] == n2 // cfg["Ncycle"] and st["iterations"] == n2


def test_density_compat_and_intended(oracle):
    ob = oracle
    cfg = CONFIGS[1]
    for compat in (L.COMPAT_ALL, 0):  # 0 also clears B14 (stale link) -- no interactions here so B3/B4 are moot
        e, os_ = make_pair(ob, cfg, chains=2, seed=3, compat=compat)
        did = e.density_create(16)
        e.density_measure(did)
        tot = None
        for s in os_:
            d = ob.Density(s, 16)
            d.measure(s)
            tot = d.read()[0] if tot is None else tot + d.read()[0]
        dg, nd, _ = e.density_read(did, 16)
        assert np.array_equal(dg, tot) and nd == 2 * cfg["M"]
        if compat == 0:
            assert dg.sum() == 2 * cfg["N"] * cfg["M"]  # intended mode counts every bead
            spec = [(1, L.UPD_RESHAPE_SWAP, 5), (1, L.UPD_RESHAPE_LINEAR, 5)]
            ge, oo = _mk_updates(ob, e, os_, spec)
            e.run(300, ge)
            for s, ups in zip(os_, oo):
                s.run(300, ups)
            _sync_paths(e, os_)


def synthetic_table(n=64, hi=12.0):
    x = np.linspace(1e-3, hi, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    return -0.004 * np.exp(-0.5 * (X + Y)) * (1 + 0.3 * np.cos(X - Y)), 1e-3, hi


@pytest.mark.parametrize("fimpl", [0, 1], ids=["warp", "one-thread"])
@pytest.mark.parametrize("compat", [L.COMPAT_ALL, L.COMPAT_PAIR_BYVALUE | L.COMPAT_DENSITY_SHIFT, 0], ids=["as-shipped", "swap-fixes", "intended"])
def test_interacting_faithful_cell_list(oracle, compat, fimpl):
    """Hard core a > 0, pair action through the lnU table, cell list queries: faithful schedule against the oracle."""
    ob = oracle
    tab, lo, hi = synthetic_table()
    cfg = dict(pot="harmonic", dim=2, M=10, N=8, L=4.0, T=0.5, lam=0.5, Ncycle=2)
    e, os_ = make_pair(ob, cfg, chains=2, seed=77, interactions=True, g=1.6, r_a=1.0, tab=tab, tab_lo=lo, tab_hi=hi, compat=compat)
    e.set_option(L.OPT_FAITHFUL_IMPL, fimpl)
    assert e.a > 0 and e.a == os_[0].a and e.nbins == os_[0].nbins == 8
    _sync_paths(e, os_)
    rng = np.random.default_rng(4)
    for t in range(30):
        c = t % 2
        s = os_[c]
        r = rng.uniform(-4, 4, 2)
        j = int(rng.integers(1, cfg["M"] + 1)); exc = int(rng.integers(1, cfg["N"] + 1))
        ex = np.array([exc], dtype=np.int64)
        nn_o = ob.lib().ora_find_nn(s.h, ob._p(r.copy()), j, ob._pi(ex), 1)
        assert e.find_nn(c, r, j, exc) == nn_o
        out = np.zeros(64, dtype=np.int64)
        cnt = ob.lib().ora_find_nns_pos(s.h, ob._p(r.copy()), j, ob._pi(ex), 1, ob._pi(out))
        assert sorted(e.find_nns(c, r, j, exc).tolist()) == sorted(out[:cnt].tolist())
    spec = [(1, L.UPD_SINGLE_COM, 0.5), (1, L.UPD_RESHAPE_LINEAR, 4), (1, L.UPD_RESHAPE_SWAP, 4), (2, L.UPD_POLYMER_COM, 0.5)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    e.run(400, ge)
    for s, ups in zip(os_, oo):
        s.run(400, ups)
    r, V, bins, nxt = e.paths()
    for c, s in enumerate(os_):
        ro, Vo, bo, no = s.paths()
        assert np.array_equal(nxt[c], no) and np.array_equal(r[c], ro) and np.array_equal(V[c], Vo) and np.array_equal(bins[c], bo)
    # explicit Delta-U hooks on the interacting configuration (pair sums through the cell list vs the oracle's lists)
    for t in range(12):
        c = t % 2
        n = int(rng.integers(1, cfg["N"] + 1)); j0 = int(rng.integers(1, cfg["M"] + 1)); m = int(rng.integers(2, cfg["M"] - 1)); u = float(rng.uniform())
        xi = 0.05 * rng.standard_normal((m - 1, 2))
        wi, wu = C.c_double(), C.c_double()
        acc_o = ob.lib().ora_reshape_linear_explicit(os_[c].h, n, j0, m, ob._p(xi), u, 0, C.byref(wi), C.byref(wu), None)
        acc, gwi, gwu, _ = e.reshape_linear_explicit(c, n, j0, m, xi, u, commit=False)
        assert acc == acc_o and close(gwi, wi.value, 1e-11, 1e-12) and close(gwu, wu.value, 1e-11, 1e-12), (t, gwi, wi.value, gwu, wu.value)
    if fimpl == 1:   # the sweep of interacting worldlines exists only with the cooperative proposals
        with pytest.raises(pj.PimcError):
            e.run(10, ge, sched=L.SCHED_SWEEP)


@pytest.mark.parametrize("isw", [0, 2], ids=["sequential-kernel", "optimistic-kernels"])
@pytest.mark.parametrize("compat", [L.COMPAT_ALL, 0], ids=["as-shipped", "intended"])
def test_interacting_sweep_sequential(oracle, compat, isw):
    """Sweep schedule of INTERACTING worldlines (hard core, lnU table, cell list): every worldline proposes once per iteration, strictly
    in order; held to the oracle's ORA_SCHED_SWEEP_SEQ bit for bit in both executions -- inside the persistent kernel (default) and by the
    optimistic-parallel kernels of pimc_isweep.cuh (PIMC_OPT_ISWEEP = 2; they apply to the as-shipped compat mode, the intended mode stays sequential)."""
    ob = oracle
    tab, lo, hi = synthetic_table()
    cfg = dict(pot="harmonic", dim=2, M=12, N=9, L=3.0, T=0.5, lam=0.5, Ncycle=3)
    e, os_ = make_pair(ob, cfg, chains=3, seed=31, interactions=True, g=3.0, r_a=1.0, tab=tab, tab_lo=lo, tab_hi=hi, compat=compat)
    e.set_option(L.OPT_ISWEEP, isw)
    assert e.a > 0
    spec = [(2, L.UPD_SINGLE_COM, 0.4), (1, L.UPD_RESHAPE_LINEAR, 6), (2, L.UPD_RESHAPE_SWAP, 6), (3, L.UPD_POLYMER_COM, 0.3)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    en_id, de_id = e.energy_create(400), e.density_create(16)
    oen = [ob.Energy(400) for _ in os_]
    ode = [ob.Density(s, 16) for s in os_]
    st = e.run(120, ge, energies=[en_id], densities=[de_id], sched=L.SCHED_SWEEP)
    for s, ups, en, de in zip(os_, oo, oen, ode):
        s.run(120, ups, energies=[en], densities=[de], sched=ob.SCHED_SWEEP_SEQ)
    _sync_paths(e, os_)
    tot = 0
    for (_, uid), k in zip(ge, range(len(spec))):
        for c in range(3):
            gq, o = e.update_get(uid, c), oo[c][k][1].get()
            assert gq["tries"] == o["tries"] and gq["tries_var"] == o["tries_var"] and gq["accepted"] == o["accepted"], (k, c, gq, o)
            assert gq["var"] == o["var"] and gq["bead_moves"] == o["bead_moves"], (k, c, gq, o)
            tot += o["bead_moves"]
    assert st["bead_moves"] == tot and st["proposals"] > 120 * 3
    scale = cfg["dim"] * cfg["N"] / (2 * os_[0].tau)
    for c in range(3):
        E, Ev, n = e.energy_read(en_id, c)
        Eo, Evo = oen[c].read()
        assert n == len(Eo) == 40 and np.all(np.abs(E - Eo) <= 1e-12 * scale)
    dg, _, _ = e.density_read(de_id, 16)
    assert np.array_equal(dg, sum(d.read()[0] for d in ode))


@pytest.mark.parametrize("compat", [L.COMPAT_ALL, 0], ids=["as-shipped", "intended"])
@pytest.mark.parametrize("cfg,g,n_it", [
    (dict(pot="harmonic", dim=2, M=40, N=24, L=3.0, T=0.5, lam=0.5, Ncycle=3), 3.5, 500),   # a = 0.166: dense, hard-core redraws and failed bridges
    (dict(pot="zero", dim=2, M=48, N=36, L=4.0, T=1.0, lam=1.0, Ncycle=4), 2.4, 400),       # windows longer than a warp (m up to 46)
    (dict(pot="sin2", dim=1, M=12, N=5, L=4.0, T=1.0, lam=1.0, Ncycle=3), 2.0, 300),        # 1-D: three-cell stencil
], ids=["dense-hardcore", "long-windows", "one-dim"])
def test_interacting_faithful_warp_stress(oracle, cfg, g, n_it, compat):
    """The warp-cooperative proposals (pimc_faithful.cuh) on denser systems: speculative hard-core bridges with redraws, pair sums of
    many neighbours per slice, windows longer than 32 slices, cycle-merging swaps; measured Energy and Density included."""
    ob = oracle
    tab, lo, hi = synthetic_table()
    e, os_ = make_pair(ob, cfg, chains=3, seed=123, interactions=True, g=g, r_a=1.0, tab=tab, tab_lo=lo, tab_hi=hi, compat=compat)
    assert e.a > 0 and e.a == os_[0].a
    exact = cfg["pot"] in ("zero", "harmonic")
    _sync_paths(e, os_, exact)
    m0 = cfg["M"] - 2
    spec = [(2, L.UPD_SINGLE_COM, 0.4), (1, L.UPD_RESHAPE_LINEAR, m0), (1, L.UPD_RESHAPE_SWAP, m0), (3, L.UPD_POLYMER_COM, 0.3)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    en_id, de_id = e.energy_create(1000), e.density_create(24)
    oen = [ob.Energy(1000) for _ in os_]
    ode = [ob.Density(s, 24) for s in os_]
    st = e.run(n_it, ge, energies=[en_id], densities=[de_id])
    for s, ups, en, de in zip(os_, oo, oen, ode):
        s.run(n_it, ups, energies=[en], densities=[de])
    _sync_paths(e, os_, exact)
    for (_, uid), k in zip(ge, range(len(spec))):
        for c in range(3):
            gq, o = e.update_get(uid, c), oo[c][k][1].get()
            assert gq["tries"] == o["tries"] and gq["accepted"] == o["accepted"] and gq["var"] == o["var"] and gq["bead_moves"] == o["bead_moves"], (k, c, gq, o)
    scale = cfg["dim"] * cfg["N"] / (2 * os_[0].tau)
    for c in range(3):
        E, Ev, n = e.energy_read(en_id, c)
        Eo, Evo = oen[c].read()
        assert n == len(Eo) and np.all(np.abs(E - Eo) <= 1e-12 * scale)
    dg, ndg, _ = e.density_read(de_id, 24)
    assert np.array_equal(dg, sum(d.read()[0] for d in ode))
    # the one-thread bodies give the same bits from the same start
    e2, _ = make_pair(ob, cfg, chains=3, seed=123, interactions=True, g=g, r_a=1.0, tab=tab, tab_lo=lo, tab_hi=hi, compat=compat)
    e2.set_option(L.OPT_FAITHFUL_IMPL, 1)
    ge2 = [(every, e2.update_create(kind, v0)) for every, kind, v0 in spec]
    e2.run(n_it, ge2, energies=[e2.energy_create(1000)], densities=[e2.density_create(24)])   # measuring runs cap redraws at 1000 (simulation.jl:31-32)
    r1, V1, b1, n1 = e.paths()
    r2, V2, b2, n2 = e2.paths()
    assert np.array_equal(r1, r2) and np.array_equal(V1, V2) and np.array_equal(b1, b2) and np.array_equal(n1, n2)


@pytest.mark.parametrize("impl", [1, 2, 3], ids=["sweep-persistent", "sweep-batched", "sweep-chain-major"])
@pytest.mark.parametrize("cfg,rng_,n_it", [
    (dict(pot="zero", dim=2, M=33, N=300, L=5.0, T=1.0, lam=1.0, Ncycle=3), 10000, 14),      # two super-batches, ragged M (KM = 2)
    (dict(pot="harmonic", dim=2, M=200, N=5, L=6.0, T=0.25, lam=0.5, Ncycle=2), 10000, 30),   # KM = 8 register tiles
    (dict(pot="harmonic", dim=2, M=12, N=9, L=4.0, T=1.0, lam=0.5, Ncycle=2), 70, 120),       # acceptance window wraps / evicts (ballot path)
    (dict(pot="harmonic", dim=2, M=12, N=9, L=4.0, T=1.0, lam=0.5, Ncycle=2), 33, 60),        # tiny window: serial fallback
    (dict(pot="sin2", dim=1, M=64, N=40, L=4.0, T=1.0, lam=1.0, Ncycle=5), 200, 40),          # 1-D, window wrap
], ids=["N300-M33", "N5-M200", "window70", "window33", "1d-window200"])
def test_sweep_edge_cases_bit_exact(oracle, cfg, rng_, n_it, impl):
    """ragged sizes, several super-batches, register-tile variants and acceptance-window wrap-around, both sweep implementations"""
    _run_sweep_case(oracle, cfg, rng_, n_it, impl)


def _run_sweep_case(oracle, cfg, rng_, n_it, impl):
    ob = oracle
    e, os_ = make_pair(ob, cfg, chains=2, seed=99)
    e.set_option(L.OPT_SWEEP_IMPL, impl)
    spec = [(1, L.UPD_SINGLE_COM, 0.8), (1, L.UPD_RESHAPE_LINEAR, 7), (4, L.UPD_POLYMER_COM, 0.4)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    for (_, uid), k in zip(ge, range(len(spec))):
        lim = (2, cfg["M"] - 2, 0.6, 0.8) if spec[k][1] == L.UPD_RESHAPE_LINEAR else (0.1, cfg["L"] / 2, 0.4, 0.6)
        e.update_configure(uid, *lim, adj=10, rng=rng_)
        for c in range(2):
            oo[c][k][1].configure(*lim, adj=10, rng=rng_)
    en_id = e.energy_create(200)
    oen = [ob.Energy(200) for _ in os_]
    e.run(n_it, ge, energies=[en_id], sched=L.SCHED_SWEEP)
    for s, ups, en in zip(os_, oo, oen):
        s.run(n_it, ups, energies=[en], sched=L.SCHED_SWEEP)
    _sync_paths(e, os_, cfg["pot"] in ("zero", "harmonic"))
    for (_, uid), k in zip(ge, range(len(spec))):
        for c in range(2):
            g, o = e.update_get(uid, c), oo[c][k][1].get()
            assert (g["tries"], g["tries_var"], g["accepted"], g["bead_moves"], g["var"]) == (o["tries"], o["tries_var"], o["accepted"], o["bead_moves"], o["var"]), (k, c, g, o)
            assert (np.isnan(g["acc_window"]) and np.isnan(o["acc_window"])) or g["acc_window"] == o["acc_window"], (k, c, g, o)
    scale = cfg["dim"] * cfg["N"] / (2 * os_[0].tau)
    for c in range(2):
        E, Ev, n = e.energy_read(en_id, c)
        Eo, Evo = oen[c].read()
        assert n == len(Eo) and np.all(np.abs(E - Eo) <= 1e-12 * scale)
    Eb, Evb, n = e.energy_read_range(en_id, 1, 2)
    assert len(Eb) == min(2, max(0, n - 1)) and np.array_equal(Eb, e.energy_read(en_id, -1)[0][1:3])


# ---- the BASELINE.json shapes, held to the oracle with the kernels bench.py times (C2, C2 in a trap, C3, C4, C5) ----
BASELINE_SHAPES = {
    "C2": (dict(pot="zero", dim=2, M=128, N=64, L=16.0, T=1.0, lam=1.0, Ncycle=2), [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 20)], "energy", 12),
    "C2-trap": (dict(pot="harmonic", dim=2, M=128, N=64, L=6.0, T=0.5, lam=0.5, Ncycle=2), [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 20)], "energy", 12),
    "C2s": (dict(pot="zero", dim=2, M=128, N=64, L=16.0, T=1.0, lam=1.0, Ncycle=2),
            [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 20), (1, L.UPD_RESHAPE_SWAP, 20)], "energy", 18),
    "C3": (dict(pot="harmonic", dim=2, M=100, N=256, L=16.0, T=0.5, lam=0.5, Ncycle=5),
           [(1, L.UPD_POLYMER_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 20), (1, L.UPD_RESHAPE_SWAP, 20)], "density", 15),
    "C4": (dict(pot="lattice", dim=2, M=256, N=128, L=8.0, T=0.2, lam=1.0 / np.pi ** 2, Ncycle=3),
           [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 5), (20, L.UPD_RESHAPE_SWAP, 20)], "density", 9),
    "C5": (dict(pot="harmonic", dim=2, M=64, N=1024, L=100.0, T=1.0, lam=0.5, Ncycle=10), [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 2)], "energy", 10),
}


def _check_against_oracle(ob, e, ge, spec, cfg, chains, measure, n_it, seed, sched_o=None, en_cap=64, nbins=500, **extra):
    """runs the oracle on `chains` (global ids) with the same seed / schedule and compares positions, link cache, permutation, counters,
    energies (<= 1e-12 dN/2tau) and the density histogram (integer-equal) of those chains"""
    po = pots(ob)[cfg["pot"]][0]
    kw = dict(dim=cfg["dim"], M=cfg["M"], N=cfg["N"], T=cfg["T"], lam=cfg["lam"], Ncycle=cfg["Ncycle"], seed=seed)
    kw.update(extra)
    exact = cfg["pot"] in ("zero", "harmonic")
    scale = cfg["dim"] * cfg["N"] / (2 * e.tau)
    dsum = None
    for c in chains:
        s = ob.System(po, L=cfg["L"], chain=c, **kw)
        ups = [(every, ob.Update(s, kind, v0)) for every, kind, v0 in spec]
        oen, ode = ob.Energy(en_cap), ob.Density(s, nbins)
        s.run(n_it, ups, energies=[oen] if measure == "energy" else [], densities=[ode] if measure == "density" else [],
              sched=ob.SCHED_SWEEP if sched_o is None else sched_o)
        r, V, bins, nxt = e.paths(c, 1)
        ro, Vo, bo, no = s.paths()
        assert np.array_equal(nxt[0], no), f"chain {c}: permutation differs"
        assert np.array_equal(r[0], ro), f"chain {c}: positions differ"
        assert np.array_equal(bins[0], bo), f"chain {c}: bins differ"
        assert np.array_equal(V[0], Vo) if exact else close(V[0], Vo, 1e-12, 1e-14), f"chain {c}: link cache differs"
        for (_, uid), (_, uo) in zip(ge, ups):
            g, o = e.update_get(uid, c), uo.get()
            assert (g["tries"], g["tries_var"], g["accepted"], g["bead_moves"], g["var"]) == (o["tries"], o["tries_var"], o["accepted"], o["bead_moves"], o["var"]), (c, g, o)
        if measure == "energy":
            E, Ev, n = e.energy_read(e._test_en, c)
            Eo, Evo = oen.read()
            assert n == len(Eo) == n_it // cfg["Ncycle"]
            assert np.all(np.abs(E - Eo) <= 1e-12 * scale) and np.all(np.abs(Ev - Evo) <= 1e-12 * np.maximum(1.0, np.abs(Evo))), c
        else:
            d = ode.read()[0]
            dsum = d if dsum is None else dsum + d
    return dsum


@pytest.mark.parametrize("impl", [3, 2], ids=["chain-major", "per-iteration"])
@pytest.mark.parametrize("name", sorted(BASELINE_SHAPES))
def test_baseline_shapes_batched_kernels_vs_oracle(oracle, name, impl):
    """the throughput kernels bench.py times -- k_chain<POT,KM> (chain-major persistent kernel, the default dispatch) and the per-iteration
    k_sweep<POT,KM> / k_swap_iter / k_measure<POT,KM> it shares its device bodies with -- at every BASELINE shape against the oracle:
    bit-exact positions / link cache / permutation / counters, E within 1e-12 dN/2tau, histogram integer-equal."""
    ob = oracle
    cfg, spec, measure, n_it = BASELINE_SHAPES[name]
    chains = 2 if cfg["N"] >= 256 else 3
    pg = pots(ob)[cfg["pot"]][1]
    e = pj.Engine(pg, chains=chains, L_=cfg["L"], dim=cfg["dim"], M=cfg["M"], N=cfg["N"], T=cfg["T"], lam=cfg["lam"], Ncycle=cfg["Ncycle"], seed=2025)
    e.set_option(L.OPT_SWEEP_IMPL, impl)       # regardless of the batch size
    ge = [(every, e.update_create(kind, v0)) for every, kind, v0 in spec]
    e._test_en = e.energy_create(64)
    de = e.density_create(500)
    kw = dict(energies=[e._test_en] if measure == "energy" else [], densities=[de] if measure == "density" else [], sched=L.SCHED_SWEEP)
    n1 = n_it // 2 + 1                         # two calls: the iteration / cadence counters carry over
    st = e.run(n1, ge, **kw)
    st2 = e.run(n_it - n1, ge, **kw)
    # one k_sweep (+ k_swap_iter, + k_measure) launch per iteration, or ONE k_chain launch per call
    assert (st["launches"] >= n1 and st2["launches"] >= n_it - n1) if impl == 2 else (st["launches"] == 1 and st2["launches"] == 1)
    dsum = _check_against_oracle(ob, e, ge, spec, cfg, range(chains), measure, n_it, 2025)
    if measure == "density":
        dg, nd, _ = e.density_read(de, 500)
        assert np.array_equal(dg, dsum) and nd == chains * (n_it // cfg["Ncycle"]) * cfg["M"]


def test_c2_default_dispatch_at_bench_scale_vs_oracle(oracle):
    """C2 with 128 chains = 2^20 beads: the DEFAULT dispatch (PIMC_OPT_SWEEP_IMPL = 0) picks the per-iteration kernels exactly as in bench.py;
    Energy fused into the sweep launch (option) and the separate estimator launch (default) give the oracle's values; spot-checked chains vs oracle."""
    ob = oracle
    cfg, spec, measure, n_it = BASELINE_SHAPES["C2"]
    out = []
    for impl, fuse in ((0, 0), (2, 1), (2, 0)):   # default dispatch (chain-major at this batch size); per-iteration kernels with and without the fused Energy
        e = pj.Engine(pots(ob)["zero"][1], chains=128, L_=cfg["L"], dim=2, M=128, N=64, T=1.0, lam=1.0, Ncycle=2, seed=7)
        e.set_option(L.OPT_SWEEP_IMPL, impl)
        e.set_option(L.OPT_FUSE_ENERGY, fuse)
        ge = [(every, e.update_create(kind, v0)) for every, kind, v0 in spec]
        e._test_en = e.energy_create(64)
        st = e.run(n_it, ge, energies=[e._test_en], sched=L.SCHED_SWEEP)
        assert st["launches"] == 1 if impl == 0 else st["launches"] >= n_it   # 128 chains fill less than two rounds of CTA slots: auto = chain-major
        _check_against_oracle(ob, e, ge, spec, cfg, [0, 37, 127], "energy", n_it, 7)
        out.append((e.paths(want=("r",))[0], e.energy_read(e._test_en, -1)[0]))
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0])
        assert np.all(np.abs(out[0][1] - o[1]) <= 1e-12 * 2 * 64 / (2 * e.tau))


def _real_table(Lbox, g, T, M):
    from pimc_jl_b200 import propint
    import math
    tau = (1.0 / T) / M
    p = propint.build_prop_int(math.ceil(math.sqrt(2) * Lbox), g, tau)      # examples/density_SRL_lattice.jl:17
    return p, tau


@pytest.mark.parametrize("sched", ["faithful", "sweep", "sweep-optimistic"])
@pytest.mark.parametrize("name", ["C3i", "C4i"])
def test_interacting_baseline_scale_real_table(oracle, name, sched):
    """The interacting BASELINE configurations at full per-chain size with the REAL pair-propagator table (propint.build_prop_int) on both
    sides: C3i (N=256, M=100, L=16, a=0.05, r_a=1 -> 32x32 cells per slice) and C4i (N=128, M=256, l25 lattice, g=2 -> a=exp(-pi),
    r_a from determine_nnrange); reference schedule and sweep schedule (oracle: ORA_SCHED_SWEEP_SEQ)."""
    import math
    from pimc_jl_b200 import propint
    ob = oracle
    if name == "C3i":
        cfg = dict(pot="harmonic", dim=2, M=100, N=256, L=16.0, T=0.5, lam=0.5, Ncycle=5)
        g = -2 * math.pi / math.log(0.05)
        spec = [(1, L.UPD_POLYMER_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 20), (1, L.UPD_RESHAPE_SWAP, 20)]
        p, tau = _real_table(cfg["L"], g, cfg["T"], cfg["M"])
        r_a = 1.0
    else:
        cfg = dict(pot="lattice", dim=2, M=256, N=128, L=8.0, T=0.2, lam=1.0 / np.pi ** 2, Ncycle=3)
        g = 2.0
        spec = [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 5), (20, L.UPD_RESHAPE_SWAP, 20)]
        p, tau = _real_table(cfg["L"], g, cfg["T"], cfg["M"])
        r_a = propint.determine_nnrange(p, tau, 1e-20, cfg["L"])
    ia = dict(interactions=True, g=g, r_a=r_a, tab=p["tab"], tab_lo=p["lo"], tab_hi=p["hi"])
    chains = 2
    e = pj.Engine(pots(ob)[cfg["pot"]][1], chains=chains, L_=cfg["L"], dim=2, M=cfg["M"], N=cfg["N"], T=cfg["T"], lam=cfg["lam"], Ncycle=cfg["Ncycle"], seed=31, **ia)
    assert e.a > 0 and (name != "C3i" or e.nbins == 32)
    if sched == "sweep-optimistic":
        e.set_option(L.OPT_ISWEEP, 2)
    ge = [(every, e.update_create(kind, v0)) for every, kind, v0 in spec]
    de = e.density_create(500)
    n_it = 150 if sched == "faithful" else 6
    e.run(n_it, ge, densities=[de], sched=L.SCHED_FAITHFUL if sched == "faithful" else L.SCHED_SWEEP)
    dsum = _check_against_oracle(ob, e, ge, spec, cfg, range(chains), "density", n_it, 31,
                                 sched_o=ob.SCHED_FAITHFUL if sched == "faithful" else ob.SCHED_SWEEP_SEQ, **ia)
    dg, nd, _ = e.density_read(de, 500)
    assert np.array_equal(dg, dsum)


def test_swap_weights_and_update_nnbins_hooks(oracle):
    """direct tests of two C-ABI hooks: pimc_swap_weights (sampleparticles table, helper.jl:230-260) bit for bit against the oracle on
    permuted worlds, and pimc_update_nnbins (nearest_neighbours.jl:182-196): a full rebuild leaves bins and every neighbour query unchanged."""
    ob = oracle
    tab, lo, hi = synthetic_table()
    cfg = dict(pot="harmonic", dim=2, M=14, N=10, L=4.0, T=0.5, lam=0.5, Ncycle=2)
    e, os_ = make_pair(ob, cfg, chains=2, seed=8, interactions=True, g=1.8, r_a=1.0, tab=tab, tab_lo=lo, tab_hi=hi)
    spec = [(1, L.UPD_RESHAPE_LINEAR, 5), (1, L.UPD_RESHAPE_SWAP, 5), (2, L.UPD_POLYMER_COM, 0.5)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    e.run(600, ge)
    for s, ups in zip(os_, oo):
        s.run(600, ups)
    _sync_paths(e, os_)
    assert any(not np.array_equal(s.paths()[3], np.arange(1, 11)) for s in os_)       # exchange cycles present
    rng = np.random.default_rng(12)
    for t in range(40):
        c = t % 2
        n1 = int(rng.integers(1, 11)); j0 = int(rng.integers(1, 15)); m = int(rng.integers(2, 13))
        w = np.zeros(10)
        ob.lib().ora_swap_weights(os_[c].h, n1, j0, m, ob._p(w))
        assert np.array_equal(e.swap_weights(c, n1, j0, m), w), (t, n1, j0, m)
    q = [(int(rng.integers(0, 2)), rng.uniform(-4, 4, 2), int(rng.integers(1, 15)), int(rng.integers(1, 11))) for _ in range(40)]
    before = [(e.find_nn(c, r, j, x), sorted(e.find_nns(c, r, j, x).tolist())) for c, r, j, x in q]
    b0 = e.paths(want=("bins",))[2]
    e.update_nnbins()
    for s in os_:
        ob.lib().ora_update_nnbins(s.h)
    assert np.array_equal(e.paths(want=("bins",))[2], b0)
    after = [(e.find_nn(c, r, j, x), sorted(e.find_nns(c, r, j, x).tolist())) for c, r, j, x in q]
    assert before == after
    for (c, r, j, x), (nn, nns) in zip(q, after):
        ex = np.array([x], dtype=np.int64)
        assert nn == ob.lib().ora_find_nn(os_[c].h, ob._p(r.copy()), j, ob._pi(ex), 1)
        out = np.zeros(64, dtype=np.int64)
        cnt = ob.lib().ora_find_nns_pos(os_[c].h, ob._p(r.copy()), j, ob._pi(ex), 1, ob._pi(out))
        assert nns == sorted(out[:cnt].tolist())
    # the rebuilt lists carry on: further moves stay on the oracle's trajectory
    e.run(100, ge)
    for s, ups in zip(os_, oo):
        s.run(100, ups)
    _sync_paths(e, os_)


def test_energy_objects_count_their_own_samples(oracle):
    """An Energy object appends at ITS OWN first free slot (measurement.jl:119-120), not at the System's N_MC: thermalise with a Density
    only, then add an Energy; use two Energy objects in different runs.  Both schedules / kernels."""
    ob = oracle
    cfg = CONFIGS[1]
    for impl, sched in ((1, L.SCHED_FAITHFUL), (2, L.SCHED_SWEEP)):
        e, os_ = make_pair(ob, cfg, chains=2, seed=5)
        e.set_option(L.OPT_SWEEP_IMPL, impl)
        spec = [(1, L.UPD_SINGLE_COM, 1.0), (1, L.UPD_RESHAPE_LINEAR, 6)]
        ge, oo = _mk_updates(ob, e, os_, spec)
        d, e1, e2 = e.density_create(16), e.energy_create(10), e.energy_create(10)
        od = [ob.Density(s, 16) for s in os_]
        o1, o2 = [ob.Energy(10) for _ in os_], [ob.Energy(10) for _ in os_]
        so = ob.SCHED_FAITHFUL if sched == L.SCHED_FAITHFUL else ob.SCHED_SWEEP
        e.run(12, ge, densities=[d], sched=sched)                      # 6 measurements, no Energy
        e.run(8, ge, energies=[e1], sched=sched)                       # 4 samples into e1 at indices 0..3
        e.run(6, ge, energies=[e2, e1], sched=sched)                   # 3 samples: e2 at 0..2, e1 at 4..6
        for c, (s, ups) in enumerate(zip(os_, oo)):
            s.run(12, ups, densities=[od[c]], sched=so)
            s.run(8, ups, energies=[o1[c]], sched=so)
            s.run(6, ups, energies=[o2[c], o1[c]], sched=so)
        scale = cfg["dim"] * cfg["N"] / (2 * os_[0].tau)
        for c in range(2):
            for gid, oe, cnt in ((e1, o1[c], 7), (e2, o2[c], 3)):
                E, Ev, n = e.energy_read(gid, c)
                Eo, _ = oe.read()
                assert n == cnt == len(Eo) and len(E) == cnt and np.all(np.abs(E - Eo) <= 1e-12 * scale), (impl, c, n, E, Eo)
        with pytest.raises(pj.PimcError):
            e.run(8, ge, energies=[e1], sched=sched)                   # 7 + 4 > 10: the reference errors on a full vector


def test_golden_vectors():
    """The CUDA path against the committed golden vectors (tests/golden/pimc_golden.npz: pure-Python restatements of levy!, teleport,
    distance, Energy, Density and the lattice potential, written from the Julia source) -- no oracle involved."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pimc_golden.npz"))
    for row, l in enumerate(gold["tp_L"]):
        assert np.array_equal(eng.teleport(gold["tp_x"][row], l), gold["tp_teleport"][row])
        assert np.array_equal(eng.distance(gold["tp_x"][row], gold["tp_y"][row], l), gold["tp_distance"][row])
    for i, (rows, dim, l, lam, tau) in enumerate(gold["levy_cases"]):
        out = eng.levy_bridge(np.ascontiguousarray(gold[f"levy{i}_r"].T)[None], tau, l, lam, gold[f"levy{i}_xi"][None])
        assert np.array_equal(out[0].T, gold[f"levy{i}_out"]), i          # bit for bit given the same Gaussian draws
    l, T, lam = gold["en_par"]
    r, nxt = gold["en_r"], gold["en_next"]
    N, dim, M = r.shape
    lat = dict(kind="lattice", dv="zero", depth=6.0, scale=1.0, sgn=-1.0, angles=list(gold["lat_angles"]))
    for pot, key in ((dict(kind="harmonic", dv="identity"), "en_harmonic"), (lat, "en_lattice")):
        e = pj.Engine(pj.make_potential(**pot), dim=dim, M=M, N=N, chains=2, L_=l, T=T, lam=lam, seed=3)
        e.set_paths(np.stack([r, r]), np.stack([nxt, nxt]))
        E, Ev, _ = e.energy_now()
        scale = dim * N / (2 * e.tau)
        assert np.all(np.abs(E - gold[key][0]) <= 1e-12 * scale) and np.all(np.abs(Ev - gold[key][1]) <= 1e-12 * max(1.0, abs(gold[key][1])))
    for compat, key in ((L.COMPAT_ALL, "dens_shift"), (0, "dens_fixed")):
        e = pj.Engine(pj.make_potential("harmonic", "identity"), dim=dim, M=M, N=N, chains=1, L_=l, T=T, lam=lam, seed=3, compat=compat)
        e.set_paths(r[None], nxt[None])
        d = e.density_create(10)
        e.density_measure(d)
        dens, nd, _ = e.density_read(d, 10)
        assert np.array_equal(dens, gold[key]) and nd == M
    V, _ = eng.potential_eval(gold["lat_pts"], pj.make_potential(**lat))
    assert close(V, gold["lat_V"], 1e-12, 1e-13)


@pytest.mark.parametrize("cfg,nb,rmax", [
    (dict(pot="harmonic", dim=2, M=20, N=13, L=3.0, T=0.6, lam=0.5, Ncycle=3), 40, 3.0),
    (dict(pot="zero", dim=2, M=33, N=70, L=4.0, T=1.0, lam=1.0, Ncycle=2), 9000, 5.0),     # > 8192 bins: global-atomic path; odd M: ragged last tile
    (dict(pot="sin2", dim=1, M=16, N=9, L=2.0, T=1.0, lam=0.5, Ncycle=4), 16, 1.5),
], ids=["trap2d", "free2d-many-bins", "1d"])
def test_paircorr_and_winding_vs_oracle(oracle, cfg, nb, rmax):
    """the estimators the reference lists as TODO (measurement.jl:125-127): g(r) pair counts integer for integer and winding numbers against
    the oracle's definitions, (i) as functors on the current configuration, (ii) inside run! with swaps (pimc_run_ex: same cadence as
    Energy / Density, which are measured alongside and must not change), (iii) on a hand-made winding worldline."""
    ob = oracle
    chains = 3
    e, os_ = make_pair(ob, cfg, chains=chains, seed=77)
    spec = [(1, L.UPD_SINGLE_COM, 0.5), (1, L.UPD_RESHAPE_LINEAR, 6), (1, L.UPD_RESHAPE_SWAP, 6)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    pc, wi, en = e.paircorr_create(nb, rmax), e.winding_create(64), e.energy_create(64)
    opc = [ob.PairCorrelation(s, nb, rmax) for s in os_]
    oen = [ob.Energy(64) for _ in os_]
    # (i) functor calls
    e.paircorr_measure(pc)
    for g, s in zip(opc, os_):
        g.measure(s)
    h, nd, b = e.paircorr_read(pc, nb)
    assert np.array_equal(h, sum(g.read()[0] for g in opc)) and nd == chains * cfg["M"] and b == rmax / nb and h.sum() > 0
    assert np.allclose(e.winding_now(), np.stack([ob.winding_now(s) for s in os_]), atol=1e-12)
    # (ii) inside run!, 41 iterations: the cadence counter carries over between the two calls
    nC = cfg["Ncycle"]
    for n_it in (23, 18):
        e.run(n_it, ge, energies=[en], paircorrs=[pc], windings=[wi], sched=L.SCHED_SWEEP)
    Wo = [[] for _ in os_]
    for c, (s, ups, g, eo) in enumerate(zip(os_, oo, opc, oen)):
        done = 0
        while done < 41:
            seg = min(nC - s.scalars()["Nctr"], 41 - done)
            s.run(seg, ups, energies=[eo], sched=ob.SCHED_SWEEP)
            done += seg
            if s.scalars()["Nctr"] == 0:
                g.measure(s)
                Wo[c].append(ob.winding_now(s))
    _sync_paths(e, os_, exact_v=cfg["pot"] in ("zero", "harmonic"))
    h, nd, _ = e.paircorr_read(pc, nb)
    nm = 41 // nC
    assert np.array_equal(h, sum(g.read()[0] for g in opc)) and nd == chains * cfg["M"] * (1 + nm)
    scale = cfg["dim"] * cfg["N"] / (2 * os_[0].tau)
    for c in range(chains):
        Wg, n = e.winding_read(wi, c)
        assert n == nm and np.allclose(Wg, np.array(Wo[c]), atol=1e-9) and np.array_equal(np.rint(Wg), np.rint(np.array(Wo[c])))
        E, _, ne = e.energy_read(en, c)
        assert ne == nm and np.all(np.abs(E - oen[c].read()[0]) <= 1e-12 * scale)
    W2, n = e.winding_read(wi, -1)
    assert n == nm and np.allclose(W2, np.mean([(np.array(w) ** 2).sum(axis=1) for w in Wo], axis=0), atol=1e-9)
    # (iii) one worldline wrapped once around x
    r, _, _, nxt = e.paths()
    M, Lb = cfg["M"], cfg["L"]
    r[1, 2, 0, :] = -Lb + (np.arange(M) + 0.5) * (2 * Lb / M)
    e.set_paths(r, nxt)
    os_[1].set_paths(r[1], nxt[1])
    W = e.winding_now()
    assert np.allclose(W[1], ob.winding_now(os_[1]), atol=1e-9) and np.rint(W[1, 0]) == 1 + np.rint(Wo[1][-1][0])
    with pytest.raises(pj.PimcError):
        e.run(64 * nC, ge, windings=[wi], sched=L.SCHED_SWEEP)   # more samples than the pre-sized series holds: refused like Energy


@pytest.mark.parametrize("cfg,kmax", [
    (dict(pot="harmonic", dim=2, M=20, N=13, L=3.0, T=0.6, lam=0.5, Ncycle=3), 4),
    (dict(pot="zero", dim=2, M=33, N=70, L=4.0, T=1.0, lam=1.0, Ncycle=2), 6),     # largest kmax, odd M: ragged last tile, N > 64: three passes per warp
    (dict(pot="sin2", dim=1, M=16, N=9, L=2.0, T=1.0, lam=0.5, Ncycle=4), 3),
], ids=["trap2d", "free2d-kmax6", "1d"])
def test_structure_factor_and_compressibility_vs_oracle(oracle, cfg, kmax):
    """`#TODO Compressibilty` (measurement.jl:127): static structure factor sums |rho_k|^2 on the box's wave vectors against the oracle's
    definition (1e-9 relative: the sums run in a different order, sincospi vs libm), (i) as a functor on the current configuration,
    (ii) inside run! with swaps (pimc_run_ex, the cadence of Energy, which is measured alongside and must not change), (iii) the
    compressibility read-out, (iv) a checkpoint taken mid-run carries the accumulators."""
    ob = oracle
    chains = 3
    e, os_ = make_pair(ob, cfg, chains=chains, seed=91)
    spec = [(1, L.UPD_SINGLE_COM, 0.5), (1, L.UPD_RESHAPE_LINEAR, 6), (1, L.UPD_RESHAPE_SWAP, 6)]
    ge, oo = _mk_updates(ob, e, os_, spec)
    sk, en = e.structure_create(kmax), e.energy_create(64)
    oen = [ob.Energy(64) for _ in os_]
    N, M, dim = cfg["N"], cfg["M"], cfg["dim"]
    a, b = np.meshgrid(np.arange(kmax + 1), np.arange(-kmax, kmax + 1), indexing="ij")
    half = ((a > 0) | (b > 0)) if dim > 1 else ((b == 0) & (a > 0))

    def same(got, want):
        return np.all(np.abs(got - want) <= 1e-9 * np.maximum(1.0, np.abs(want))) and np.all(got[~half] == 0) and np.all(got[half] > 0)
    # (i) functor call
    e.structure_measure(sk)
    ref = sum(ob.structure_now(s, kmax) for s in os_)
    got, nd = e.structure_read(sk, kmax)
    assert nd == chains * M and same(got, ref)
    # independent numpy evaluation of one entry: k = (pi / L)(1, -1) (2-D) or (1) (1-D), chain 0
    r0 = e.paths(0, 1, want=("r",))[0][0]                       # [N][dim][M]
    ph = np.pi / cfg["L"] * (r0[:, 0, :] - (r0[:, 1, :] if dim > 1 else 0.0))
    one = (np.abs(np.exp(1j * ph).sum(axis=0)) ** 2).sum()
    assert abs(ob.structure_now(os_[0], kmax)[1, kmax - (1 if dim > 1 else 0)] - one) <= 1e-9 * one
    # (ii) inside run!
    nC = cfg["Ncycle"]
    blob = None
    for n_it in (23, 18):
        e.run(n_it, ge, energies=[en], structures=[sk], sched=L.SCHED_SWEEP)
        if blob is None:
            blob = e.get_state()
            mid = e.structure_read(sk, kmax)
    for s, ups, eo in zip(os_, oo, oen):
        done = 0
        while done < 41:
            seg = min(nC - s.scalars()["Nctr"], 41 - done)
            s.run(seg, ups, energies=[eo], sched=ob.SCHED_SWEEP)
            done += seg
            if s.scalars()["Nctr"] == 0:
                ref = ref + ob.structure_now(s, kmax)
    _sync_paths(e, os_, exact_v=cfg["pot"] in ("zero", "harmonic"))
    got, nd = e.structure_read(sk, kmax)
    nm = 41 // nC
    assert nd == chains * M * (1 + nm) and same(got, ref)
    scale = dim * N / (2 * os_[0].tau)
    for c in range(chains):
        E, _, ne = e.energy_read(en, c)
        assert ne == nm and np.all(np.abs(E - oen[c].read()[0]) <= 1e-12 * scale)
    # (iii) kappa_T = beta S(k_min) / rho on the smallest shell
    kappa, s0 = e.compressibility(sk)
    Sk = got / (nd * N)
    s0_ref = 0.5 * (Sk[1, kmax] + Sk[0, kmax + 1]) if dim > 1 else Sk[1, kmax]
    beta, rho = 1.0 / cfg["T"], N / (2 * cfg["L"]) ** dim
    assert abs(s0 - s0_ref) <= 1e-12 * s0_ref and abs(kappa - beta * s0_ref / rho) <= 1e-12 * kappa
    # (iv) the state blob carries the accumulators: restore the mid-run checkpoint, finish the run, same sums bit for bit
    e.set_state(blob)
    g2, n2 = e.structure_read(sk, kmax)
    assert n2 == mid[1] and np.array_equal(g2, mid[0])
    e.run(18, ge, energies=[en], structures=[sk], sched=L.SCHED_SWEEP)
    g3, n3 = e.structure_read(sk, kmax)
    assert n3 == nd and np.array_equal(g3, got)
    with pytest.raises(pj.PimcError):
        e.structure_create(7)                                    # kmax beyond the kernel's table
