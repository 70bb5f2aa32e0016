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


def test_reshape_swap_delta_u_and_commit_against_numpy(oracle):
    """ReshapeSwapLinear for independent worldlines (src/updates/reshape.jl:123-283), restated from the Julia source: crossed end points
    (r1 ends on fpcycle2(j_m), r2 on fpcycle1(j_m), :148-151), w_initial = cached links of both strands (:165), w_updated = sum(V1) + sum(V2)
    (:244), and on acceptance the commit in the reference's order -- exchange `next`, RE-COMPUTE both cycles, write rows 2..m+1 and links 1..m
    through the new cycles (:252-268), exchange the tails j_m+1..M of r and V when j_m < M (:269-275; the cached link AT j_m stays where it was:
    B14).  Permutations with 1-, 2- and 3-cycles, windows that wrap onto the next member of the cycle."""
    ob = oracle
    rng = np.random.default_rng(17)
    M, N, L, lam = 7, 4, 3.0, 0.5
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=N, L=L, T=0.7, lam=lam, seed=23)
    r_start = rng.uniform(-L, L, (N, 2, M))
    s.set_paths(r_start, np.array([2, 1, 4, 3], dtype=np.int64))
    mod1 = lambda j: (j - 1) % M + 1

    def cycle(nxt, n):                                   # subcycle (helper.jl:64-85): the cycle of n, starting at n
        pol, p = [n], nxt[n - 1]
        while p != n:
            pol.append(p); p = nxt[p - 1]
        return pol
    naccept = nwrap = ntail = 0
    for trial in range(120):
        r, V, _, nxt = s.paths()
        nxt = [int(x) for x in nxt]
        n1 = int(rng.integers(1, N + 1)); n2 = int(rng.integers(1, N)); n2 += n2 >= n1
        j0, m = int(rng.integers(1, M + 1)), int(rng.integers(2, M - 1))
        jm = j0 + m
        xi1, xi2 = 0.3 * rng.standard_normal((m - 1, 2)), 0.3 * rng.standard_normal((m - 1, 2))
        u = float(rng.uniform())
        pol1, pol2 = cycle(nxt, n1), cycle(nxt, n2)
        fp = lambda pol, j: pol[(1 + (j - 1) // M - 1) % len(pol)]          # pcycle (helper.jl:269-271)
        r1, r2 = np.zeros((m + 1, 2)), np.zeros((m + 1, 2))
        r1[0], r2[0] = r[n1 - 1, :, j0 - 1], r[n2 - 1, :, j0 - 1]
        r1[m], r2[m] = r[fp(pol2, jm) - 1, :, mod1(jm) - 1], r[fp(pol1, jm) - 1, :, mod1(jm) - 1]
        b1, b2 = levy_py(r1, s.tau, L, lam, xi1), levy_py(r2, s.tau, L, lam, xi2)
        w_i = 0.0
        for j in range(j0, jm):
            w_i += V[fp(pol1, j) - 1, mod1(j) - 1] + V[fp(pol2, j) - 1, mod1(j) - 1]
        V1 = [_lnV_py(b1[k], b1[k + 1], s.tau, _harm) for k in range(m)]
        V2 = [_lnV_py(b2[k], b2[k + 1], s.tau, _harm) for k in range(m)]
        w_u = sum(V1) + sum(V2)
        delta = math.exp(w_u - w_i)
        acc_py = delta >= 1.0 or delta > u
        wi, wu = C.c_double(), C.c_double()
        acc = ob.lib().ora_reshape_swap_explicit(s.h, n1, n2, j0, m, ob._p(xi1), ob._p(xi2), u, 1, C.byref(wi), C.byref(wu))
        assert abs(wi.value - w_i) <= 1e-12 * max(1, abs(w_i)) and abs(wu.value - w_u) <= 1e-12 * max(1, abs(w_u)), (trial, wi.value, w_i, wu.value, w_u)
        assert bool(acc) == acc_py, trial
        exp_r, exp_V, exp_n = r.copy(), V.copy(), list(nxt)
        if acc_py:
            naccept += 1; nwrap += jm - 1 > M; ntail += jm < M
            exp_n[n1 - 1], exp_n[n2 - 1] = nxt[n2 - 1], nxt[n1 - 1]
            q1, q2 = cycle(exp_n, n1), cycle(exp_n, n2)                    # the closures see the re-computed cycles (:254-257)
            for j in range(2, m + 2):
                exp_r[fp(q1, j0 + j - 1) - 1, :, mod1(j0 + j - 1) - 1] = b1[j - 1]
                exp_r[fp(q2, j0 + j - 1) - 1, :, mod1(j0 + j - 1) - 1] = b2[j - 1]
            for j in range(1, m + 1):
                exp_V[fp(q1, j0 + j - 1) - 1, mod1(j0 + j - 1) - 1] = V1[j - 1]
                exp_V[fp(q2, j0 + j - 1) - 1, mod1(j0 + j - 1) - 1] = V2[j - 1]
            if jm < M:
                for j in range(jm + 1, M + 1):
                    exp_r[n1 - 1, :, j - 1], exp_r[n2 - 1, :, j - 1] = exp_r[n2 - 1, :, j - 1].copy(), exp_r[n1 - 1, :, j - 1].copy()
                    exp_V[n1 - 1, j - 1], exp_V[n2 - 1, j - 1] = exp_V[n2 - 1, j - 1], exp_V[n1 - 1, j - 1]
        r2_, V2_, _, n2_ = s.paths()
        assert [int(x) for x in n2_] == exp_n, trial
        assert np.array_equal(r2_, exp_r), trial
        assert np.allclose(V2_, exp_V, rtol=1e-14, atol=0), trial
    assert naccept > 20 and nwrap > 3 and ntail > 3, (naccept, nwrap, ntail)


def test_interacting_swap_pair_action_against_numpy(oracle):
    """The one place where the pair action enters the moves as shipped: ReshapeSwapLinear on an interacting System (reshape.jl:163-243),
    restated from the Julia source in numpy -- find_nns (stencil of the bead's cell, periodic Euclidean distance <= one cell width,
    nearest_neighbours.jl:72-154), next() (helper.jl:286-300), lnU = p < 0 ? -mu : log(p), p = 1 + terms(|r|, |r'|) / prop_rel0(r, r', tau)
    (system.jl:25-28, propagator.jl:73-86) with the scaled linear B-spline of the term table, the component-wise |distance| vectors of
    propagator.jl:6-9 -- and with BOTH the old-pair and the new-pair sums added to w_initial (B4, reshape.jl:182-239).  Proposals are not
    committed, so the permutation (a 3-cycle and a fixed point) and the cell lists stay as set."""
    ob = oracle
    rng = np.random.default_rng(29)
    M, N, L, lam, T, mu = 6, 4, 3.0, 0.5, 0.8, 0.3
    n_tab, lo, hi = 24, 1e-3, 9.0
    x = np.linspace(lo, hi, n_tab)
    X, Y = np.meshgrid(x, x, indexing="ij")
    tab = -0.6 * np.exp(-0.5 * (X + Y)) * (1 + 0.3 * np.cos(X - 2 * Y))         # not symmetric: pins the (|r|, |r'|) argument order
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=N, L=L, T=T, lam=lam, mu=mu, interactions=True, g=0.3, r_a=1.0,
                  tab=tab, tab_lo=lo, tab_hi=hi, seed=31)
    r0 = rng.uniform(-1.2, 1.2, (N, 2, M))                                      # a dense cloud: every bead has neighbours within a cell width
    nxt = [2, 3, 1, 4]
    s.set_paths(r0, np.array(nxt, dtype=np.int64))
    r, V, bins, _ = s.paths()
    tau, nb = s.tau, 6
    w_cell = 2 * L / nb
    mod1 = lambda j: (j - 1) % M + 1
    dist = lambda a, b: np.array([distance_py(a[k], b[k], L) for k in range(2)])

    def bin_py(p):
        ix, iy = (int(math.floor((p[k] + L) / w_cell)) for k in range(2))
        return ix, iy
    assert all(bins[n, j] == bin_py(r[n, :, j])[0] + nb * bin_py(r[n, :, j])[1] + 1 for n in range(N) for j in range(M))

    def find_nns_py(pos, sl, exc):                       # particles (1-based) within one cell width of pos at slice sl
        ix, iy = bin_py(pos)
        cells = {((ix + dx) % nb, (iy + dy) % nb) for dx in (-1, 0, 1) for dy in (-1, 0, 1)}
        out = []
        for n in range(1, N + 1):
            if n in exc or bin_py(r[n - 1, :, sl - 1]) not in cells:
                continue
            if math.sqrt(float(np.sum(dist(r[n - 1, :, sl - 1], pos) ** 2))) <= w_cell:
                out.append(n)
        return out
    next_py = lambda n, sl: ((nxt[n - 1] if sl == M else n), mod1(sl + 1))

    def terms_py(a, b):                                  # scale(interpolate(A, BSpline(Linear())), A_r1, A_r2)
        h = (hi - lo) / (n_tab - 1)
        ta, tb = (a - lo) / h, (b - lo) / h
        ia, ib = min(max(int(math.floor(ta)), 0), n_tab - 2), min(max(int(math.floor(tb)), 0), n_tab - 2)
        fa, fb = ta - ia, tb - ib
        return (1 - fb) * ((1 - fa) * tab[ia, ib] + fa * tab[ia + 1, ib]) + fb * ((1 - fa) * tab[ia, ib + 1] + fa * tab[ia + 1, ib + 1])

    def lnU_py(r1, r2):
        p = 1 + terms_py(float(np.linalg.norm(r1)), float(np.linalg.norm(r2))) / (math.exp(-float((r1 - r2) @ (r1 - r2)) / (4 * tau)) / (4 * math.pi * tau))
        return -mu if p < 0.0 else math.log(p)

    def cycle(n):
        pol, p = [n], nxt[n - 1]
        while p != n:
            pol.append(p); p = nxt[p - 1]
        return pol
    npair = nneg = 0
    for trial in range(80):
        n1 = int(rng.integers(1, N + 1)); n2 = int(rng.integers(1, N)); n2 += n2 >= n1
        j0, m = int(rng.integers(1, M + 1)), int(rng.integers(2, M - 1))
        jm = j0 + m
        xi1, xi2 = 0.2 * rng.standard_normal((m - 1, 2)), 0.2 * rng.standard_normal((m - 1, 2))
        pol1, pol2 = cycle(n1), cycle(n2)
        fp = lambda pol, j: pol[((j - 1) // M) % len(pol)]
        r1, r2 = np.zeros((m + 1, 2)), np.zeros((m + 1, 2))
        r1[0], r2[0] = r[n1 - 1, :, j0 - 1], r[n2 - 1, :, j0 - 1]
        r1[m], r2[m] = r[fp(pol2, jm) - 1, :, mod1(jm) - 1], r[fp(pol1, jm) - 1, :, mod1(jm) - 1]
        b1, b2 = levy_py(r1, tau, L, lam, xi1), levy_py(r2, tau, L, lam, xi2)     # a = 8e-10: hardspherelevy! never redraws
        w_i = 0.0
        for j in range(j0, jm):
            sl = mod1(j)
            for pol in (pol1, pol2):
                p = fp(pol, j)
                w_i += V[p - 1, sl - 1]
            for pol in (pol1, pol2):                       # old pairs of both strands
                p = fp(pol, j)
                pn, pj = next_py(p, sl)
                for nn in find_nns_py(r[p - 1, :, sl - 1], sl, [p]):
                    qn, qj = next_py(nn, sl)
                    u_ = lnU_py(dist(r[nn - 1, :, sl - 1], r[p - 1, :, sl - 1]), dist(r[qn - 1, :, qj - 1], r[pn - 1, :, pj - 1]))
                    w_i += u_; npair += 1; nneg += u_ == -mu
        V1 = [_lnV_py(b1[k], b1[k + 1], tau, _harm) for k in range(m)]
        V2 = [_lnV_py(b2[k], b2[k + 1], tau, _harm) for k in range(m)]
        for j in range(j0, jm):                            # new pairs: added to w_initial as well (B4)
            sl, k = mod1(j), j - j0
            exc = [fp(pol1, j), fp(pol2, j)]
            for b in (b1, b2):
                for nn in find_nns_py(b[k], sl, exc):
                    qn, qj = next_py(nn, sl)
                    w_i += lnU_py(dist(r[nn - 1, :, sl - 1], b[k]), dist(r[qn - 1, :, qj - 1], b[k + 1])); npair += 1
        w_u = sum(V1) + sum(V2)
        wi, wu = C.c_double(), C.c_double()
        u = float(rng.uniform())
        acc = ob.lib().ora_reshape_swap_explicit(s.h, n1, n2, j0, m, ob._p(xi1), ob._p(xi2), u, 0, C.byref(wi), C.byref(wu))
        assert abs(wi.value - w_i) <= 1e-11 * max(1, abs(w_i)) and abs(wu.value - w_u) <= 1e-12 * max(1, abs(w_u)), (trial, wi.value, w_i, wu.value, w_u)
        d = math.exp(w_u - w_i)
        assert bool(acc) == (d >= 1.0 or d > u), trial
    assert npair > 500 and 0 < nneg < npair, (npair, nneg)


def test_hardcore_bridge_redraws_against_numpy(oracle):
    """hardspherelevy! (src/updates/helper.jl:141-181) inside ReshapeLinear on a dense hard-core System, restated from the Julia source: every
    interior bead is redrawn (fresh Gaussians: the addressed draw (bead, retry) of the RNG spec) while the nearest other particle at that slice
    -- find_nn over the 3 x 3 cell stencil of the TELEPORTED candidate, the moved worldline excluded (nearest_neighbours.jl:156-179) -- lies
    closer than s.a; the candidate is stored un-teleported until the final wrap; more than s.ctr tries give up the proposal.  V = 0 and the
    pair action does not enter ReshapeLinear as shipped, so a proposal is accepted iff its bridge could be laid; the committed rows must equal
    the restated bridge bit for bit."""
    ob = oracle
    rng = np.random.default_rng(41)
    M, N, L, lam, seed = 8, 7, 1.5, 0.5, 77
    s = ob.System(ob.make_potential("zero", "identity"), dim=2, M=M, N=N, L=L, T=1.0, lam=lam, interactions=True, g=6.0, r_a=0.5, seed=seed,
                  tab=np.zeros((4, 4)), tab_lo=1e-3, tab_hi=6.0)            # terms = 0: lnU = log(1) = 0, only the hard core acts
    a = s.scalars()["a"]
    assert abs(a - math.exp(-2 * math.pi / 6.0)) < 1e-15                    # a = interactions ? exp(-2 pi / g) : 0, system.jl:151
    u = ob.Update(s, ob.UPD_RESHAPE_LINEAR, M - 2)
    ob.lib().ora_set_ctr(s.h, 6)                                             # few tries: proposals that give up are exercised too
    nb = int(math.floor(2 * L / 0.5)); w_cell = 2 * L / nb
    mod1 = lambda j: (j - 1) % M + 1
    bin_py = lambda p: tuple(min(max(int(math.floor((p[k] + L) / w_cell)), 0), nb - 1) for k in range(2))

    def gauss(it, bead, retry):
        g0, g1 = C.c_double(), C.c_double()
        ob.lib().ora_gauss_pair(seed, 0, it, 0, 2, retry, bead, C.byref(g0), C.byref(g1))   # kind 2 = PIMC_K_BRIDGE, slot 0
        return np.array([g0.value, g1.value])
    nretry = nfail = nacc = 0
    for it in range(1, 161):
        ob.lib().ora_set_iter(s.h, it)
        r, _, bins, nxt = s.paths()
        n, j0 = int(rng.integers(1, N + 1)), int(rng.integers(1, M + 1))
        bm0 = u.get()["bead_moves"]
        acc = ob.lib().ora_update_call(s.h, u.h, 0, n, j0)
        m = u.get()["bead_moves"] - bm0 + 1
        own = lambda j: (n if j <= M else int(nxt[n - 1]), mod1(j))
        rp = np.zeros((m + 1, 2))
        rp[0] = r[n - 1, :, j0 - 1]
        pe, je = own(j0 + m)
        rp[m] = r[pe - 1, :, je - 1]
        for k in range(2):
            if abs(rp[0, k] - rp[m, k]) > L:
                rp[m, k] += np.sign(rp[0, k]) * (2 * L)
        ok = True
        mi = m - 1                                                            # interior beads (rows - 2)
        for j in range(1, mi + 1):
            alpha = (mi + 1 - j) / (mi + 2 - j)
            sl = mod1(j0 + j)
            ctr, placed = 0, False
            while True:
                ctr += 1
                if ctr > 6:
                    break
                rp[j] = alpha * rp[j - 1] + (1 - alpha) * rp[m] + gauss(it, j, ctr - 1) * math.sqrt(2 * lam * alpha * s.tau)
                tp = np.array([teleport_py(rp[j, k], L) for k in range(2)])
                bx, by = bin_py(tp)
                cells = {((bx + dx) % nb, (by + dy) % nb) for dx in (-1, 0, 1) for dy in (-1, 0, 1)}
                best = None
                for q in range(1, N + 1):
                    if q == n:
                        continue
                    b = int(bins[q - 1, sl - 1]) - 1
                    if (b % nb, b // nb) not in cells:
                        continue
                    d = math.sqrt(sum(distance_py(tp[k], r[q - 1, k, sl - 1], L) ** 2 for k in range(2)))
                    best = d if best is None else min(best, d)
                if best is not None and best < a:
                    nretry += 1
                    continue
                placed = True
                break
            if not placed:
                ok = False
                break
        r2 = s.paths()[0]
        if not ok:
            nfail += 1
            assert acc == 0 and np.array_equal(r2, r), it
            continue
        nacc += 1
        exp_r = r.copy()
        for k in range(m):                                                    # rows 1..m: the first end point is rewritten with its wrapped self
            p, sl = own(j0 + k)
            exp_r[p - 1, :, sl - 1] = [teleport_py(rp[k, 0], L), teleport_py(rp[k, 1], L)]
        assert acc == 1 and np.array_equal(r2, exp_r), it
    assert nretry > 30 and nfail > 0 and nacc > 60, (nretry, nfail, nacc)


def test_hardcore_centre_of_mass_retries_against_numpy(oracle):
    """move_polymer! (src/updates/helper.jl:368-395) inside SingleCenterOfMass (com.jl:136-224) on a dense hard-core System, restated from the
    Julia source: a uniform displacement d = maxd * 2 * (rand(dim) - 0.5) is redrawn (the addressed draw `retry` of the RNG spec) while any
    displaced, wrapped bead has its nearest other particle of the same slice closer than s.a; more than s.ctr tries give up.  V = 0: a
    proposal whose displacement could be placed is accepted, and the committed worldline must equal teleport(r + d) bit for bit."""
    ob = oracle
    rng = np.random.default_rng(43)
    M, N, L, seed = 6, 7, 1.5, 99
    s = ob.System(ob.make_potential("zero", "identity"), dim=2, M=M, N=N, L=L, T=1.0, lam=0.5, interactions=True, g=6.0, r_a=0.5, seed=seed,
                  tab=np.zeros((4, 4)), tab_lo=1e-3, tab_hi=6.0)
    a = s.scalars()["a"]
    u = ob.Update(s, ob.UPD_SINGLE_COM, 0.6)
    ob.lib().ora_set_ctr(s.h, 4)
    nb = int(math.floor(2 * L / 0.5)); w_cell = 2 * L / nb
    bin_py = lambda p: tuple(min(max(int(math.floor((p[k] + L) / w_cell)), 0), nb - 1) for k in range(2))
    nretry = nfail = nacc = 0
    for it in range(1, 201):
        ob.lib().ora_set_iter(s.h, it)
        r, _, bins, nxt = s.paths()
        n = int(rng.integers(1, N + 1))
        maxd = u.get()["var"]
        acc = ob.lib().ora_update_call(s.h, u.h, 0, n, 0)
        new = None
        for ctr in range(1, 5):
            u0, u1 = C.c_double(), C.c_double()
            ob.lib().ora_uniform_pair(seed, 0, it, 0, 4, ctr - 1, 0, C.byref(u0), C.byref(u1))      # kind 4 = PIMC_K_COM
            d = np.array([maxd * 2 * (u0.value - 0.5), maxd * 2 * (u1.value - 0.5)])
            cand = np.zeros((2, M)); hit = False
            for j in range(1, M + 1):
                c = np.array([teleport_py(r[n - 1, k, j - 1] + d[k], L) for k in range(2)])
                cand[:, j - 1] = c
                bx, by = bin_py(c)
                cells = {((bx + dx) % nb, (by + dy) % nb) for dx in (-1, 0, 1) for dy in (-1, 0, 1)}
                best = None
                for q in range(1, N + 1):
                    b = int(bins[q - 1, j - 1]) - 1
                    if q == n or (b % nb, b // nb) not in cells:
                        continue
                    dq = math.sqrt(sum(distance_py(c[k], r[q - 1, k, j - 1], L) ** 2 for k in range(2)))
                    best = dq if best is None else min(best, dq)
                if best is not None and best < a:
                    hit = True
                    break
            if not hit:
                new = cand
                break
            nretry += 1
        r2 = s.paths()[0]
        if new is None:
            nfail += 1
            assert acc == 0 and np.array_equal(r2, r), it
        else:
            nacc += 1
            exp_r = r.copy(); exp_r[n - 1] = new
            assert acc == 1 and np.array_equal(r2, exp_r), it
    assert nretry > 30 and nfail > 0 and nacc > 60, (nretry, nfail, nacc)


def test_init_world_against_numpy(oracle):
    """init_world (src/system.jl:36-78) restated from the Julia source, with the addressed draws of the RNG spec (start point: kind INIT0,
    retry = attempt; ring Gaussians: kind INIT, retry = index of the levy! call): uniform start, closed ring by levy! between two copies of the
    start point, and the hard-core rejection loop as written -- a hit re-bridges r IN PLACE and the scan over slices and earlier particles
    simply continues on the new ring before the whole attempt is repeated.  Positions bit for bit, link cache to 1e-14."""
    ob = oracle
    M, N, L, lam, seed, T = 6, 8, 1.5, 0.5, 1234, 1.0
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=M, N=N, L=L, T=T, lam=lam, interactions=True, g=6.0, r_a=0.5, seed=seed,
                  tab=np.zeros((4, 4)), tab_lo=1e-3, tab_hi=6.0)
    a, tau = s.scalars()["a"], s.tau
    r_o, V_o, _, nxt = s.paths()

    def start(slot, attempt):
        u0, u1 = C.c_double(), C.c_double()
        ob.lib().ora_uniform_pair(seed, 0, 0, slot, 7, attempt, 0, C.byref(u0), C.byref(u1))       # kind 7 = PIMC_K_INIT0
        return np.array([2 * L * (u0.value - 0.5), 2 * L * (u1.value - 0.5)])

    def ring(r, slot, call):
        xi = np.zeros((M - 2, 2))
        for t in range(1, M - 1):
            g0, g1 = C.c_double(), C.c_double()
            ob.lib().ora_gauss_pair(seed, 0, 0, slot, 6, call, t, C.byref(g0), C.byref(g1))        # kind 6 = PIMC_K_INIT
            xi[t - 1] = g0.value, g1.value
        return levy_py(r, tau, L, lam, xi)
    world, nrebridge = [], 0
    for n in range(1, N + 1):
        slot, calls, ctr = n - 1, 0, 0
        r = np.zeros((M, 2))
        passed = n != 1
        while passed:
            passed = False
            ctr += 1
            assert ctr <= 10000
            r[0] = start(slot, ctr - 1); r[-1] = r[0]
            r = ring(r, slot, calls); calls += 1
            for m in range(M):
                for i in range(n - 1):
                    d = math.sqrt(sum(distance_py(world[i][m, k], r[m, k], L) ** 2 for k in range(2)))
                    if d < a:
                        r = ring(r, slot, calls); calls += 1
                        nrebridge += 1
                        passed = True
        if n == 1:
            r[0] = start(slot, 0); r[-1] = r[0]
            r = ring(r, slot, calls)
        world.append(r)
    mine = np.stack([w.T for w in world])                                       # [N][dim][M]
    assert nrebridge > 3
    assert np.array_equal(mine, r_o) and list(nxt) == list(range(1, N + 1))
    V_mine = np.array([[_lnV_py(world[n][m], world[n][(m + 1) % M], tau, _harm) for m in range(M)] for n in range(N)])
    assert np.allclose(V_o, V_mine, rtol=1e-14, atol=0)
    for n in range(N):                                                          # the result respects the hard core at every slice
        for i in range(n):
            for m in range(M):
                assert math.sqrt(sum(distance_py(world[i][m, k], world[n][m, k], L) ** 2 for k in range(2))) >= a


def test_run_update_choice_and_measurement_cadence(oracle):
    """run! (src/simulation.jl:29-42) and measurement_Z_sector (src/measurement.jl:1-17) restated: per iteration one update drawn with
    weights 1 ./ every by StatsBase's `sample(wv)` walk (t = rand() * sum(w); advance while the running sum is < t) from the chain-level
    uniform of the RNG spec; every Ncycle-th iteration of runs WITH Zmeasurements takes a measurement, the counter carrying over between
    calls; s.ctr = 1000 with measurements, 10000 without."""
    ob = oracle
    seed, Ncycle = 4321, 3
    s = ob.System(ob.make_potential("harmonic", "identity"), dim=2, M=6, N=3, L=4.0, T=1.0, lam=0.5, Ncycle=Ncycle, seed=seed)
    every = [1, 2, 4]
    ups = [(every[0], ob.Update(s, ob.UPD_SINGLE_COM, 0.5)), (every[1], ob.Update(s, ob.UPD_RESHAPE_LINEAR, 3)), (every[2], ob.Update(s, ob.UPD_RESHAPE_SWAP, 3))]
    en = ob.Energy(200)
    w = [1.0 / e for e in every]

    def picks(it0, n):
        cnt = [0, 0, 0]
        for it in range(it0, it0 + n):
            u0, u1 = C.c_double(), C.c_double()
            ob.lib().ora_uniform_pair(seed, 0, it, 0xFFFF, 0, 0, 0, C.byref(u0), C.byref(u1))   # slot PIMC_SLOT_CHAIN, kind PIMC_K_ITER
            t, i, cw = u0.value * sum(w), 0, w[0]
            while cw < t and i < 2:
                i += 1; cw += w[i]
            cnt[i] += 1
        return cnt
    it0 = s.scalars()["iter"]
    s.run(40, ups)                                            # thermalisation: no measurements, no cadence
    sc = s.scalars()
    assert sc["ctr"] == 10000 and sc["N_MC"] == 0 and sc["Nctr"] == 0 and sc["iter"] == it0 + 40
    assert [u.get()["tries"] for _, u in ups] == picks(it0, 40)
    s.run(47, ups, energies=[en])
    sc = s.scalars()
    assert sc["ctr"] == 1000 and sc["N_MC"] == 47 // Ncycle and sc["Nctr"] == 47 % Ncycle
    s.run(10, ups, energies=[en])                             # 2 carried over + 10 = 4 more measurements
    sc = s.scalars()
    assert sc["N_MC"] == 57 // Ncycle and sc["Nctr"] == 57 % Ncycle and len(en.read()[0]) == 57 // Ncycle
    tot = picks(it0, 97)
    assert [u.get()["tries"] for _, u in ups] == tot and sum(tot) == 97 and min(tot) > 5
