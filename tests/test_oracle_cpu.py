"""CPU tests (no GPU): the oracle against independent numpy / mpmath re-derivations of the reference formulas,
against the constants the reference's own example and tests hold, and against closed-form finite-M answers."""
import ctypes as C
import math
import numpy as np
import pytest


# ---------- independent second opinions (pure Python / numpy, written from the Julia source, not from the C) ----------
def teleport_py(x, L):                      # src/propagator.jl:30-32
    return (x + L) - math.floor(x / (2 * L) + 0.5) * (2 * L) - L


def distance_py(a, b, L):                   # src/propagator.jl:6-9
    dx = abs(a - b)
    return min((2 * L) - dx, dx)


def levy_py(r, tau, L, lam, xi):            # src/updates/helper.jl:118-139 ; r rows x dim (row index first)
    r = r.copy()
    rows, dim = r.shape
    for k in range(dim):
        if abs(r[0, k] - r[-1, k]) > L:
            r[-1, k] += np.sign(r[0, k]) * (2 * L)
    m = rows - 2
    for j in range(1, m + 1):
        alpha = (m + 1 - j) / (m + 2 - j)
        r[j, :] = alpha * r[j - 1, :] + (1 - alpha) * r[-1, :] + xi[j - 1, :] * math.sqrt(2 * lam * alpha * tau)
    for j in range(rows):
        for k in range(dim):
            r[j, k] = teleport_py(r[j, k], L)
    return r


def energy_py(r, nxt, L, tau, lam, V, dV):  # src/measurement.jl:92-122 ; r[n][dim][M]
    N, dim, M = r.shape
    link = pot = vkin = 0.0
    for i in range(N):
        for j in range(M):
            inext = nxt[i] - 1 if j == M - 1 else i
            jn = (j + 1) % M
            a, b = r[i, :, j], r[inext, :, jn]
            dr = np.array([distance_py(a[k], b[k], L) for k in range(dim)])
            link += float(dr @ dr)
            pot += V(a) + V(b)
            vkin += float(a @ dV(a))
    E = dim * N / (2 * tau) - 1 / (4 * lam * tau ** 2 * M) * link + 1 / (2 * M) * pot
    Ev = 1 / (2 * M) * vkin + 1 / (2 * M) * pot
    return E, Ev


def test_teleport_distance_against_python(oracle):
    ob = oracle
    rng = np.random.default_rng(0)
    for L in (4.0, 100.0, 0.37):
        for x in np.concatenate([rng.uniform(-5 * L, 5 * L, 500), [L, -L, 0.0, 3 * L, -3 * L, 1e-300, -1e-17]]):
            assert ob.lib().ora_teleport(x, L) == teleport_py(x, L)
            t = ob.lib().ora_teleport(x, L)
            assert -L <= t <= L
            y = rng.uniform(-L, L)
            assert ob.lib().ora_distance(x, y, L) == distance_py(x, y, L)


@pytest.mark.parametrize("dim", [1, 2])
def test_levy_against_numpy(oracle, dim):
    ob = oracle
    rng = np.random.default_rng(1)
    for rows, L, lam, tau in [(3, 4.0, 1.0, 0.01), (12, 100.0, 0.5, 0.2), (100, 4.0, 1.0, 0.01), (6, 0.5, 2.0, 0.3)]:
        for trial in range(20):
            r = np.zeros((rows, dim))
            r[0], r[-1] = rng.uniform(-L, L, dim), rng.uniform(-L, L, dim)
            if trial == 0:
                r[0], r[-1] = 0.95 * L, -0.95 * L
            xi = rng.standard_normal((rows - 2, dim))
            ref = levy_py(r, tau, L, lam, xi)
            cm = np.ascontiguousarray(r.T)  # oracle layout: column-major rows x dim
            ob.lib().ora_levy(ob._p(cm), rows, dim, tau, L, lam, ob._p(np.ascontiguousarray(xi)))
            assert np.array_equal(cm.T, ref)
            # bridge sanity: endpoints fixed (mod box), interior finite
            assert np.all(np.abs(cm) <= L)


def test_bridge_statistics(oracle):
    """Levy bridge between equal endpoints: bead t of an m-link bridge has variance 2*lam*tau*t*(m-t)/m (free particle)."""
    ob = oracle
    rng = np.random.default_rng(2)
    rows, L, lam, tau, nb = 9, 1e6, 0.7, 0.13, 20000
    out = np.zeros((nb, rows))
    for b in range(nb):
        cm = np.zeros((1, rows))
        xi = rng.standard_normal((rows - 2, 1))
        ob.lib().ora_levy(ob._p(cm), rows, 1, tau, L, lam, ob._p(xi))
        out[b] = cm[0]
    m = rows - 1
    for t in range(1, m):
        var = 2 * lam * tau * t * (m - t) / m
        assert abs(out[:, t].var() / var - 1) < 0.05


def test_energy_density_against_numpy(oracle):
    ob = oracle
    rng = np.random.default_rng(3)
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=7, N=4, L=3.0, T=0.8, lam=0.5, seed=5)
    r = rng.uniform(-3, 3, (4, 2, 7))
    nxt = np.array([2, 1, 3, 4], dtype=np.int64)  # one exchange cycle of length 2
    s.set_paths(r, nxt)
    E, Ev, parts = s.energy_now()
    Er, Evr = energy_py(r, nxt, 3.0, s.tau, 0.5, lambda x: 0.5 * (x[0] ** 2 + x[1] ** 2), lambda x: x)
    assert abs(E - Er) <= 1e-12 * abs(4 * 2 / (2 * s.tau)) and abs(Ev - Evr) <= 1e-12 * abs(Evr)
    # Density (src/measurement.jl:45-55): compat (shifted, floor-bin 0 dropped) vs numpy
    d = ob.Density(s, 10)
    d.measure(s)
    dens, nd, binw = d.read()
    ref = np.zeros((10, 10))
    for n in range(4):
        for m in range(7):
            ib = np.floor((r[n, :, m] + 3.0) / binw).astype(int)
            if np.all(ib > 0) and np.all(ib < 11):
                ref[ib[0] - 1, ib[1] - 1] += 1
    assert np.array_equal(dens, ref) and nd == 7 and binw == 0.6


def test_lattice_intensity_reference_kat(oracle):
    """test/testpotential.jl:26-30: the 3-beam lattice intensity is 1.0 at the origin and at two lattice peaks."""
    ob = oracle
    p = ob.make_potential("lattice", depth=1.0, scale=1.0, sgn=1.0, angles=[2 * math.pi * k / 3 for k in range(3)])
    for pt in ([0.0, 0.0], [2 / math.sqrt(3), 0.0], [0.0, 2 / 3]):
        v = ob.lib().ora_potential_eval(C.byref(p), ob._p(np.array(pt)), 2)
        assert v == pytest.approx(1.0, rel=1e-12)


def test_system_constructor_smoke_reference_tests(oracle):
    """test/testsystem.jl:8-35: default System(v1d) has N == 2; the 2-D lattice system has N == 5."""
    ob = oracle
    s = ob.System(ob.make_potential("sin2_1d", depth=8.0, scale=0.5), dim=1)
    assert s.N == 2 and s.M == 100 and s.nbins == 8 and s.a == 0.0
    ang = [2 * math.pi * k / 4 for k in range(4)]
    s = ob.System(ob.make_potential("lattice", depth=8.0, scale=0.5, sgn=1.0, angles=ang), dim=2, M=100, N=5, L=4.0, T=1.0)
    assert s.N == 5
    r, V, bins, nxt = s.paths()
    assert np.all(np.abs(r) <= 4.0) and np.array_equal(nxt, np.arange(1, 6))
    assert np.array_equal(r[:, :, 0], r[:, :, -1])  # closed ring: last slice sits on the first (system.jl:53-54)


def test_periodic_bounds_reference_test(oracle):
    """test/testsystem.jl:37-56: after 10 000 mixed updates on V = 0 no bead has left the box."""
    ob = oracle
    s = ob.System(ob.make_potential("zero"), seed=42)
    ups = [(2, ob.Update(s, ob.UPD_SINGLE_COM, 3.0)), (1, ob.Update(s, ob.UPD_RESHAPE_LINEAR, 20)), (1, ob.Update(s, ob.UPD_RESHAPE_SWAP, 20))]
    s.run(10000, ups)
    r, V, bins, nxt = s.paths()
    assert np.all(r <= s.L) and np.all(r >= -s.L)
    assert sorted(nxt.tolist()) == [1, 2]  # still a permutation
    assert np.allclose(V, 0.0)
    for _, u in ups:
        g = u.get()
        assert g["tries"] > 1000 and 0 <= g["acc_window"] <= 1


def test_density_normalisation_reference_test(oracle):
    """test/testmeasurements.jl:1-33: sum(dens)/ndata ~ N within 1e-2 (the tolerance absorbs the dropped floor-bin 0)."""
    ob = oracle
    s = ob.System(ob.make_potential("sin2_1d", depth=8.0, scale=0.5), dim=1, seed=9)
    ups = [(2, ob.Update(s, ob.UPD_SINGLE_COM, 3.0)), (1, ob.Update(s, ob.UPD_RESHAPE_LINEAR, 20)), (1, ob.Update(s, ob.UPD_RESHAPE_SWAP, 20))]
    d = ob.Density(s, 500)
    s.run(10000, ups, densities=[d])
    dens, nd, _ = d.read()
    assert nd == (10000 // 10) * 100
    assert abs(dens.sum() / nd - s.N) < 1e-2


def harmonic_energy_finite_M(M, beta=1.0, dim=2, omega=1.0):
    """closed form for the primitive action: Z_M = [2 sinh(M theta/2)]^-dim, cosh(theta) = 1 + (omega tau)^2/2."""
    import mpmath as mp
    mp.mp.dps = 30

    def lnZ(b):
        tau = b / M
        th = mp.acosh(1 + (omega * tau) ** 2 / 2)
        return -dim * mp.log(2 * mp.sinh(M * th / 2))
    return float(-mp.diff(lnZ, beta))


def test_closed_form_constants():
    assert harmonic_energy_finite_M(5) == pytest.approx(2.156259612426946, rel=1e-10)
    assert harmonic_energy_finite_M(10) == pytest.approx(2.162019287996358, rel=1e-10)
    assert 1 / math.tanh(0.5) == pytest.approx(2.1639534137386534, rel=1e-14)  # examples/energy_2d_harmonically_trapped_bose_gas.jl:28


def _chain_means(ob, chains, therm, n, seed0, M=5):
    Es, Evs = [], []
    for c in range(chains):
        s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=1, L=100.0, T=1.0, lam=0.5, Ncycle=10, seed=seed0, chain=c)
        ups = [(1, ob.Update(s, ob.UPD_SINGLE_COM, 1.0)), (1, ob.Update(s, ob.UPD_RESHAPE_LINEAR, 2))]
        s.run(therm, ups)
        e = ob.Energy(n // 10 + 1)
        s.run(n, ups, energies=[e])
        E, Ev = e.read()
        Es.append(E.mean())
        Evs.append(Ev.mean())
    return np.array(Es), np.array(Evs)


def test_c1_energy_z_test(oracle):
    """C1 as shipped (examples/energy_2d_harmonically_trapped_bose_gas.jl): sampled <E_thermo>, <E_virial> agree with the
    finite-M closed form 2.1562596 within |z| < 4 (independent chains give the error bar)."""
    ob = oracle
    Es, Evs = _chain_means(ob, 24, 30000, 150000, seed0=123)
    target = 2.156259612426946
    for x in (Es, Evs):
        z = (x.mean() - target) / (x.std(ddof=1) / math.sqrt(len(x)))
        assert abs(z) < 4, (x.mean(), z)


def test_adjust_rules(oracle):
    ob = oracle
    L = ob.lib()
    assert L.ora_adjust_step(1.0, 0.1, 2.0, 0.4, 0.6, 0.3) == 0.9
    assert L.ora_adjust_step(1.0, 0.1, 2.0, 0.4, 0.6, 0.7) == 1.1
    assert L.ora_adjust_step(3.0, 0.1, 2.0, 0.4, 0.6, 0.5) == 2.0
    assert L.ora_adjust_step(3.0, 0.1, 2.0, 0.4, 0.6, float("nan")) == 2.0  # empty window: only the clamps act
    assert L.ora_adjust_slices(5, 2, 8, 0.6, 0.8, 0.5) == 4
    assert L.ora_adjust_slices(8, 2, 8, 0.6, 0.8, 0.9) == 8
    assert L.ora_adjust_slices(2, 2, 8, 0.6, 0.8, 0.1) == 2
    assert L.ora_metropolis(1.0, 0.999) == 1 and L.ora_metropolis(0.5, 0.6) == 0 and L.ora_metropolis(0.5, 0.4) == 1
    assert L.ora_metropolis(float("nan"), 0.0) == 0


def test_cycles_and_bins(oracle):
    ob = oracle
    s = ob.System(ob.make_potential("zero"), dim=2, M=6, N=5, L=4.0, seed=2)
    r = s.paths()[0]
    s.set_paths(r, np.array([3, 2, 5, 4, 1], dtype=np.int64))  # cycle 1->3->5->1, fixed points 2, 4
    cyc = np.zeros(6, dtype=np.int64)
    n = ob.lib().ora_subcycle(s.h, 1, ob._pi(cyc))
    assert n == 3 and cyc[:3].tolist() == [1, 3, 5]
    assert ob.lib().ora_cycle_findprev(s.h, 1) == 5
    pol = np.array([1, 3, 5], dtype=np.int64)
    got = [ob.lib().ora_pcycle(j, ob._pi(pol), 3, 6) for j in (1, 6, 7, 12, 13, 18, 19)]
    assert got == [1, 1, 3, 3, 5, 5, 1]  # helper.jl:113-115
    nb = np.zeros(9, dtype=np.int64)
    ob.lib().ora_bin_neighbors(1, 8, 2, ob._pi(nb))
    assert nb.tolist() == [1, 16, 9, 10, 8, 2, 64, 57, 58]  # nearest_neighbours.jl:55-65 (periodic 3x3 stencil)
    assert ob.lib().ora_bin(ob._p(np.array([-4.0, -4.0])), 2, 8, 4.0) == 1
    assert ob.lib().ora_bin(ob._p(np.array([3.99, -4.0])), 2, 8, 4.0) == 8
    assert ob.lib().ora_bin(ob._p(np.array([-4.0, 3.99])), 2, 8, 4.0) == 57
    # every bead is filed in the cell its stored bin names
    r, V, bins, nxt = s.paths()
    for j in (1, 4):
        for n_ in range(1, 6):
            out = np.zeros(16, dtype=np.int64)
            k = ob.lib().ora_nn_cell(s.h, j, int(bins[n_ - 1, j - 1]), ob._pi(out))
            assert n_ in out[:k].tolist()


def test_find_nn_brute_force(oracle):
    ob = oracle
    s = ob.System(ob.make_potential("zero"), dim=2, M=4, N=40, L=4.0, r_a=1.0, seed=8)
    r = s.paths()[0]
    rng = np.random.default_rng(5)
    for t in range(50):
        q = rng.uniform(-4, 4, 2)
        j = int(rng.integers(1, 5))
        exc = np.array([int(rng.integers(1, 41))], dtype=np.int64)
        out = np.zeros(400, dtype=np.int64)
        cnt = ob.lib().ora_find_nns_pos(s.h, ob._p(q.copy()), j, ob._pi(exc), 1, ob._pi(out))
        d = np.sqrt(sum(np.minimum(np.abs(r[:, k, j - 1] - q[k]), 8 - np.abs(r[:, k, j - 1] - q[k])) ** 2 for k in range(2)))
        brute = sorted(int(i + 1) for i in np.nonzero(d <= 1.0)[0] if i + 1 != exc[0])
        assert sorted(out[:cnt].tolist()) == brute  # cell width == cutoff: the stencil holds every particle within range
        nn = ob.lib().ora_find_nn(s.h, ob._p(q.copy()), j, ob._pi(exc), 1)
        if brute:
            dm = d.copy()
            dm[exc[0] - 1] = np.inf
            assert nn == int(np.argmin(dm)) + 1


def test_swap_detailed_state(oracle):
    """an accepted swap exchanges `next`, rewrites the bridged beads and swaps the tails (reshape.jl:250-278); the link cache
    stays consistent with the positions."""
    ob = oracle
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=8, N=3, L=4.0, T=1.0, lam=0.5, seed=4)
    r0, V0, _, n0 = s.paths()
    rng = np.random.default_rng(6)
    xi1, xi2 = 0.01 * rng.standard_normal((2, 2)), 0.01 * rng.standard_normal((2, 2))
    wi, wu = C.c_double(), C.c_double()
    acc = ob.lib().ora_reshape_swap_explicit(s.h, 1, 2, 2, 3, ob._p(xi1), ob._p(xi2), 0.0, 1, C.byref(wi), C.byref(wu))
    assert acc == 1  # u = 0 accepts whenever delta > 0
    r1, V1, _, n1 = s.paths()
    assert n1.tolist() == [2, 1, 3]
    assert np.array_equal(r1[0, :, 5:], r0[1, :, 5:]) and np.array_equal(r1[1, :, 5:], r0[0, :, 5:])  # tails j_m+1..M swapped
    assert np.array_equal(r1[0, :, :2], r0[0, :, :2]) and np.array_equal(r1[2], r0[2])
    assert abs(ob.lib().ora_action_links(s.h) - ob.lib().ora_action_links_recomputed(s.h)) < 1e-12


def test_sweep_equals_sequential_definition(oracle):
    """the sweep schedule is executed sequentially by the oracle; total proposals per iteration equal N for the staging move."""
    ob = oracle
    s = ob.System(ob.make_potential("zero"), dim=2, M=16, N=7, L=4.0, seed=3)
    u = ob.Update(s, ob.UPD_RESHAPE_LINEAR, 5)
    s.run(11, [(1, u)], sched=ob.SCHED_SWEEP)
    g = u.get()
    assert g["tries"] == 77 and g["accepted"] == 77 and g["var"] == 12.0  # free particles: always accepted; m grows by one whenever tries crosses a multiple of adj (7 of 11 sweeps)
    with pytest.raises(RuntimeError):
        s2 = ob.System(ob.make_potential("zero"), dim=2, M=8, N=3, L=4.0, interactions=True, g=1.5, r_a=1.0)
        s2.run(1, [(1, ob.Update(s2, ob.UPD_RESHAPE_LINEAR, 3))], sched=ob.SCHED_SWEEP)


def _hardcore_system(ob, chain, seed, sched_tag, N=4, M=8):
    return ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=N, L=3.0, T=1.0, lam=0.5, Ncycle=4, seed=seed, chain=chain,
                     interactions=True, g=3.5, r_a=1.0)   # a = exp(-2 pi / 3.5) = 0.166, lnU == 0 (no table): pure hard core


def test_interacting_sequential_sweep_invariants_and_z_test(oracle):
    """ORA_SCHED_SWEEP_SEQ (the definition the GPU sweep for interacting worldlines will be held to): every
    worldline proposes once per iteration, strictly in order.  Invariants after a run: hard core respected, beads in the box,
    every bead filed in the cell its position names; and the sampled energy agrees with the reference schedule's (|z| < 4)."""
    ob = oracle
    N, M = 4, 8
    spec = [(1, ob.UPD_SINGLE_COM, 0.5), (1, ob.UPD_RESHAPE_LINEAR, 4)]
    means = {}
    for tag, sched, therm, n in (("faithful", ob.SCHED_FAITHFUL, 8000, 120000), ("sweep", ob.SCHED_SWEEP_SEQ, 2000, 30000)):
        Es = []
        for c in range(16):
            s = _hardcore_system(ob, c, 77, tag, N, M)
            ups = [(every, ob.Update(s, kind, v0)) for every, kind, v0 in spec]
            s.run(therm, ups, sched=sched)
            e = ob.Energy(n // 4 + 1)
            s.run(n, ups, energies=[e], sched=sched)
            Es.append(e.read()[0].mean())
            if c == 0:
                r, V, bins, nxt = s.paths()
                assert np.all(np.abs(r) <= 3.0)
                for j in range(M):
                    for a_ in range(N):
                        assert bins[a_, j] == ob.lib().ora_bin(ob._p(np.ascontiguousarray(r[a_, :, j])), 2, s.nbins, 3.0)
                        out = np.zeros(16, dtype=np.int64)
                        k = ob.lib().ora_nn_cell(s.h, j + 1, int(bins[a_, j]), ob._pi(out))
                        assert out[:k].tolist().count(a_ + 1) == 1
                        for b_ in range(a_ + 1, N):
                            d = np.abs(r[a_, :, j] - r[b_, :, j]); d = np.minimum(d, 6.0 - d)
                            assert math.hypot(*d) >= s.a
                assert abs(ob.lib().ora_action_links(s.h) - ob.lib().ora_action_links_recomputed(s.h)) < 1e-10
                if tag == "sweep":
                    g = ups[1][1].get()
                    assert g["tries"] % N == 0 and g["tries"] > 0   # N proposals per picked iteration
        means[tag] = np.array(Es)
    a, b = means["faithful"], means["sweep"]
    z = (a.mean() - b.mean()) / math.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    assert abs(z) < 4, (a.mean(), b.mean(), z)
    assert a.mean() > 8.6 and b.mean() > 8.6   # the hard core is felt: the same gas without it has <E> = 8.48(3)


# ---------- second opinions for the move functors (numpy, written from src/updates/*.jl) ----------
def _lnV_py(a, b, tau, V):
    return -0.5 * tau * (V(a) + V(b))


def _harm(r):
    return 0.5 * float(np.sum(np.asarray(r) ** 2))


def test_reshape_linear_delta_u_against_numpy(oracle):
    """ReshapeLinear (src/updates/reshape.jl:56-87): proposal = levy bridge between the fixed beads, w_initial = cached links,
    w_updated = lnV of the proposed links, commit writes rows 1..m and the m link-cache entries (incl. wrap to the next particle)."""
    ob = oracle
    rng = np.random.default_rng(11)
    M, N, L, lam = 9, 3, 3.0, 0.5
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=N, L=L, T=0.7, lam=lam, seed=21)
    r0 = rng.uniform(-L, L, (N, 2, M))
    s.set_paths(r0, np.array([2, 3, 1], dtype=np.int64))       # one 3-cycle: segments may wrap onto the next particle
    nxt = [2, 3, 1]
    for trial in range(60):
        r, V, _, _ = s.paths()
        n, j0, m = int(rng.integers(1, N + 1)), int(rng.integers(1, M + 1)), int(rng.integers(2, M - 1))
        xi = rng.standard_normal((m - 1, 2))
        u = float(rng.uniform())
        own = lambda j: (n if j <= M else nxt[n - 1], (j - 1) % M)            # (particle, 0-based slice) of unwrapped slice j
        rp = np.zeros((m + 1, 2))
        rp[0] = r[n - 1, :, j0 - 1]
        pe, je = own(j0 + m)
        rp[-1] = r[pe - 1, :, je]
        bridge = levy_py(rp, s.tau, L, lam, xi)
        w_i = sum(V[own(j)[0] - 1, own(j)[1]] for j in range(j0, j0 + m))
        Vp = [_lnV_py(bridge[k], bridge[k + 1], s.tau, _harm) for k in range(m)]
        w_u = sum(Vp)
        acc_py = (math.exp(w_u - w_i) >= 1.0) or (math.exp(w_u - w_i) > u)
        wi, wu = C.c_double(), C.c_double()
        rpo = np.zeros((2, m + 1))
        acc = ob.lib().ora_reshape_linear_explicit(s.h, n, j0, m, ob._p(xi), u, 1, C.byref(wi), C.byref(wu), ob._p(rpo))
        assert np.array_equal(rpo.T, bridge)
        assert abs(wi.value - w_i) <= 1e-12 * max(1, abs(w_i)) and abs(wu.value - w_u) <= 1e-12 * max(1, abs(w_u))
        assert bool(acc) == acc_py
        r2, V2, _, _ = s.paths()
        exp_r, exp_V = r.copy(), V.copy()
        if acc:
            for k in range(m):
                p, sl = own(j0 + k)
                exp_r[p - 1, :, sl] = bridge[k]
                exp_V[p - 1, sl] = Vp[k]
        assert np.array_equal(r2, exp_r) and np.allclose(V2, exp_V, rtol=1e-14, atol=0)


def test_com_delta_u_against_numpy(oracle):
    """Single/PolymerCenterOfMass (src/updates/com.jl:47-100,168-220): every bead of the cycle shifted by d and wrapped."""
    ob = oracle
    rng = np.random.default_rng(12)
    M, N, L = 7, 4, 3.0
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=N, L=L, T=1.0, lam=0.5, seed=3)
    s.set_paths(rng.uniform(-L, L, (N, 2, M)), np.array([2, 1, 3, 4], dtype=np.int64))
    cycles = {1: [1, 2], 2: [2, 1], 3: [3], 4: [4]}
    nxt = [2, 1, 3, 4]
    for trial in range(40):
        r, V, _, _ = s.paths()
        n = int(rng.integers(1, N + 1))
        d = rng.uniform(-1.5, 1.5, 2)
        u = float(rng.uniform())
        pol = cycles[n]
        newr = {p: np.array([[teleport_py(r[p - 1, k, j] + d[k], L) for k in range(2)] for j in range(M)]) for p in pol}
        w_i = sum(V[p - 1].sum() for p in pol)
        Vp = {p: [_lnV_py(newr[p][j], newr[p][j + 1] if j < M - 1 else newr[nxt[p - 1]][0], s.tau, _harm) for j in range(M)] for p in pol}
        w_u = sum(sum(Vp[p]) for p in pol)
        wi, wu = C.c_double(), C.c_double()
        acc = ob.lib().ora_com_explicit(s.h, n, 1, ob._p(d.copy()), u, 1, C.byref(wi), C.byref(wu))
        assert abs(wi.value - w_i) <= 1e-12 * max(1, abs(w_i)) and abs(wu.value - w_u) <= 1e-12 * max(1, abs(w_u))
        assert bool(acc) == ((math.exp(w_u - w_i) >= 1.0) or (math.exp(w_u - w_i) > u))
        r2, V2, _, _ = s.paths()
        if acc:
            for p in pol:
                assert np.array_equal(r2[p - 1].T, newr[p]) and np.allclose(V2[p - 1], Vp[p], rtol=1e-14, atol=0)
        else:
            assert np.array_equal(r2, r)


def test_swap_weights_against_numpy(oracle):
    """sampleparticles (src/updates/helper.jl:224-267): table exp(lnK(n1 -> end_i) + lnK(i -> end_n1)) over all i."""
    ob = oracle
    rng = np.random.default_rng(13)
    M, N, L, lam = 8, 5, 3.0, 0.7
    s = ob.System(ob.make_potential("zero"), dim=2, M=M, N=N, L=L, T=0.9, lam=lam, seed=8)
    r = rng.uniform(-L, L, (N, 2, M))
    nxt = np.array([2, 3, 1, 5, 4], dtype=np.int64)
    s.set_paths(r, nxt)
    for n1, j0, m in [(1, 2, 3), (4, 7, 4), (3, 8, 2), (5, 5, 6)]:
        w = np.zeros(N)
        ob.lib().ora_swap_weights(s.h, n1, j0, m, ob._p(w))
        jm = (j0 + m - 1) % M
        endp = lambda i: (nxt[i - 1] if j0 + m > M else i)
        lnk = lambda a, b: -sum(distance_py(a[k], b[k], L) ** 2 for k in range(2)) / (4 * (m * s.tau) * lam)
        ref = [math.exp(lnk(r[n1 - 1, :, j0 - 1], r[endp(i) - 1, :, jm]) + lnk(r[i - 1, :, j0 - 1], r[endp(n1) - 1, :, jm])) for i in range(1, N + 1)]
        assert np.allclose(w, ref, rtol=1e-13, atol=0)


def test_paircorr_and_winding_estimators_against_numpy(oracle):
    ob = oracle
    """the two estimators the reference lists as TODO (measurement.jl:125-127) as the oracle defines them, against direct numpy evaluations:
    g(r) pair counts (minimum-image distances, all pairs, all slices) and the winding number (an integer; non-zero for a worldline that
    wraps the box)"""
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=12, N=7, L=2.0, T=0.7, lam=0.5, seed=3)
    g = ob.PairCorrelation(s, 25, 2.5)
    g.measure(s)
    g.measure(s)
    hist, nd, b = g.read()
    r = s.paths()[0]                                        # [N][dim][M]
    ref = np.zeros(25)
    for m in range(12):
        for i in range(7):
            for j in range(i + 1, 7):
                d = np.abs(r[i, :, m] - r[j, :, m])
                d = np.minimum(2 * 2.0 - d, d)
                ib = math.floor(math.sqrt(d[0] * d[0] + d[1] * d[1]) / (2.5 / 25))
                if ib < 25:
                    ref[ib] += 1
    assert nd == 24 and b == 2.5 / 25 and np.array_equal(hist, 2 * ref) and hist.sum() > 0
    assert np.allclose(ob.winding_now(s), 0.0, atol=1e-12)  # init_world closes every ring inside the box
    # a worldline that winds once around x: beads advance by 2L / M per slice and wrap
    M, L_ = 12, 2.0
    r2 = r.copy()
    x = -L_ + (np.arange(M) + 0.5) * (2 * L_ / M)
    r2[3, 0, :] = x
    r2[3, 1, :] = 0.25
    s.set_paths(r2, s.paths()[3])
    W = ob.winding_now(s)
    assert abs(W[0] - 1.0) < 1e-12 and abs(W[1]) < 1e-12


@pytest.mark.parametrize("dim", [1, 2])
def test_structure_factor_against_numpy(oracle, dim):
    """`#TODO Compressibilty` (measurement.jl:127) as the oracle defines it -- sums over the slices of |rho_k|^2 on the wave vectors of the
    periodic box -- against a direct numpy evaluation, plus two closed forms: a perfect lattice of N = n^2 sites at spacing 2L / n scatters
    only at the reciprocal-lattice vectors (|rho_k|^2 = N^2 at (a, b) = (n, 0), zero at the other k), and one particle gives |rho_k|^2 = 1."""
    ob = oracle
    kmax, M, N, L_ = 4, 6, 9, 1.5
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=dim, M=M, N=N, L=L_, T=0.7, lam=0.5, seed=11)
    S = ob.structure_now(s, kmax)
    r = s.paths()[0]                                        # [N][dim][M]
    for a in range(kmax + 1):
        for b in range(-kmax, kmax + 1):
            indep = (a > 0 or b > 0) if dim == 2 else (b == 0 and a > 0)
            ph = np.pi / L_ * (a * r[:, 0, :] + (b * r[:, 1, :] if dim == 2 else 0.0))
            want = (np.abs(np.exp(1j * ph).sum(axis=0)) ** 2).sum() if indep else 0.0
            assert abs(S[a, b + kmax] - want) <= 1e-12 * max(1.0, want)
    if dim == 2:   # 3 x 3 square lattice, every slice the same
        g = -L_ + (np.arange(3) + 0.25) * (2 * L_ / 3)
        r2 = np.stack([np.stack([np.full(M, x), np.full(M, y)]) for x in g for y in g])
        s.set_paths(r2, s.paths()[3])
        S = ob.structure_now(s, kmax)
        assert abs(S[3, kmax] - M * N * N) < 1e-9 and abs(S[0, kmax + 3] - M * N * N) < 1e-9 and abs(S[3, kmax + 3] - M * N * N) < 1e-9
        assert abs(S[1, kmax]) < 1e-9 and abs(S[2, kmax + 1]) < 1e-9 and abs(S[4, kmax - 2]) < 1e-9
    s1 = ob.System(ob.make_potential("zero", "identity"), dim=dim, M=M, N=1, L=L_, T=0.7, lam=0.5, seed=5)
    S1 = ob.structure_now(s1, 2)
    nv = 2 * 2 + 2 * 2 * 2 if dim == 2 else 2              # independent vectors: kmax + kmax (2 kmax + 1) in 2-D, kmax in 1-D
    assert np.count_nonzero(S1) == nv and np.allclose(S1[S1 != 0], M, atol=1e-12)


def test_structure_factor_ideal_gas_limit(oracle):
    """physical pin of the estimator: init_world (system.jl:36-78) places every worldline uniformly and independently in the periodic box, so
    for V = 0 the density modes are uncorrelated and <|rho_k|^2> = N at every k != 0, i.e. S(k) = 1 and kappa_T = beta / rho (ideal gas).
    Mean over 300 independent systems of S on the smallest shell; its standard error is ~ 1 / sqrt(300) (|rho_k|^2 / N is exponential-like)."""
    ob = oracle
    N, M, L_, kmax, nsys = 6, 4, 2.0, 2, 300
    acc = np.zeros((kmax + 1, 2 * kmax + 1))
    for seed in range(nsys):
        s = ob.System(ob.make_potential("zero", "identity"), dim=2, M=M, N=N, L=L_, T=1.0, lam=1.0, seed=1000 + seed)
        acc += ob.structure_now(s, kmax)
    S = acc / (nsys * M * N)
    shell = np.array([S[1, kmax], S[0, kmax + 1], S[1, kmax + 1], S[1, kmax - 1]])
    assert np.all(np.abs(shell - 1.0) < 0.25), shell
    assert abs(S[(np.arange(kmax + 1)[:, None] > 0) | (np.arange(-kmax, kmax + 1)[None, :] > 0)].mean() - 1.0) < 0.08
