"""Engine: numpy-facing wrapper of one libpimc_b200 handle (C chains of one reference `System` on one GPU)."""
import ctypes as C
import numpy as np
from . import _lib as L


def _p(a):
    return a.ctypes.data_as(L.f64p)


def _pi(a):
    return a.ctypes.data_as(L.i64p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def comm_unique_id():
    """128 bytes of an ncclUniqueId (pimc_comm_get_unique_id); rank 0 creates it and ships it to the other ranks"""
    buf = (C.c_ubyte * 128)()
    L.check(L.load().pimc_comm_get_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


class Engine:
    def __init__(self, pot=None, dim=2, M=100, N=2, chains=1, chain_offset=0, mu=0.0, L_=4.0, T=1.0, lam=1.0,
                 interactions=False, g=0.0, r_a=0.0, Ncycle=10, compat=L.COMPAT_ALL, init=True, seed=0x5EEDB200,
                 tab=None, tab_lo=0.0, tab_hi=1.0, device=-1):
        self.lib = L.load()
        c = L.Config()
        c.dim, c.M, c.N, c.chains, c.chain_offset = dim, M, N, chains, chain_offset
        c.mu, c.lam, c.L, c.T = mu, lam, L_, T
        c.interactions, c.g, c.r_a, c.Ncycle, c.compat, c.init = int(interactions), g, r_a, Ncycle, compat, int(init)
        c.seed, c.device = seed, device
        c.pot = pot if pot is not None else L.make_potential()
        self._tab = None
        if tab is not None:
            self._tab = np.asfortranarray(tab, dtype=np.float64)
            c.tab = _p(self._tab)
            c.tab_n, c.tab_lo, c.tab_hi = self._tab.shape[0], tab_lo, tab_hi
        self.cfg = c
        self.dim, self.M, self.N, self.C, self.L = dim, M, N, chains, L_
        h = C.c_void_p()
        L.check(self.lib.pimc_create(C.byref(c), C.byref(h)))
        self.h = h
        sc = self.scalars()
        self.beta, self.tau, self.a, self.nbins = sc["beta"], sc["tau"], sc["a"], sc["nbins"]

    def close(self):
        if getattr(self, "h", None):
            self.lib.pimc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        return L.check(rc, self.h)

    # ---- state ----
    def scalars(self):
        out = np.zeros(5)
        io = np.zeros(5, dtype=np.int64)
        self._ck(self.lib.pimc_get_scalars(self.h, _p(out), _pi(io)))
        return dict(beta=out[0], tau=out[1], vol=out[2], a=out[3], r_a=out[4], nbins=int(io[0]), N_MC=int(io[1]),
                    Nctr=int(io[2]), ctr=int(io[3]), iter=int(io[4]))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.pimc_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, option, value):
        self._ck(self.lib.pimc_set_option(self.h, option, value))

    def set_iter(self, it):
        self._ck(self.lib.pimc_set_iter(self.h, it))

    def comm_init(self, nranks, rank, unique_id):
        """attach an NCCL communicator (pimc_comm_init): estimator read-outs with chain = -1 / density_read become global and collective"""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.pimc_comm_init(self.h, nranks, rank, C.cast(buf, C.c_void_p)))

    def comm_info(self):
        nr, rk, ver, tot = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        self._ck(self.lib.pimc_comm_info(self.h, C.byref(nr), C.byref(rk), C.byref(tot), C.byref(ver)))
        return dict(nranks=nr.value, rank=rk.value, chains_total=tot.value, nccl_version=ver.value)

    def get_state(self):
        """the complete chain state as one uint8 array (pimc_get_state): resumes bit for bit through set_state"""
        n = C.c_int64()
        self._ck(self.lib.pimc_state_size(self.h, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        self._ck(self.lib.pimc_get_state(self.h, buf.ctypes.data_as(C.c_void_p), n.value))
        return buf

    def set_state(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self._ck(self.lib.pimc_set_state(self.h, buf.ctypes.data_as(C.c_void_p), buf.size))

    def paths(self, chain0=0, nchains=None, want=("r", "V", "bins", "next")):
        nc = self.C - chain0 if nchains is None else nchains
        r = np.zeros((nc, self.N, self.dim, self.M)) if "r" in want else None
        V = np.zeros((nc, self.N, self.M)) if "V" in want else None
        bins = np.zeros((nc, self.N, self.M), dtype=np.int64) if "bins" in want else None
        nxt = np.zeros((nc, self.N), dtype=np.int64) if "next" in want else None
        self._ck(self.lib.pimc_get_paths(self.h, chain0, nc, _p(r) if r is not None else None, _p(V) if V is not None else None,
                                         _pi(bins) if bins is not None else None, _pi(nxt) if nxt is not None else None))
        return r, V, bins, nxt

    def get_r_into(self, out, chain0=0):
        """device -> caller's (possibly pinned) host buffer, positions only"""
        self._ck(self.lib.pimc_get_paths(self.h, chain0, out.shape[0], _p(out), None, None, None))

    def set_paths(self, r, nxt=None, chain0=0):
        r = _f(r)
        nc = r.shape[0]
        assert r.shape == (nc, self.N, self.dim, self.M), r.shape
        n_ptr = None
        if nxt is not None:
            nxt = np.ascontiguousarray(nxt, dtype=np.int64)
            n_ptr = _pi(nxt)
        self._ck(self.lib.pimc_set_paths(self.h, chain0, nc, _p(r), n_ptr))

    # ---- estimators ----
    def energy_now(self):
        E, Ev, parts = np.zeros(self.C), np.zeros(self.C), np.zeros((self.C, 3))
        self._ck(self.lib.pimc_energy_now(self.h, _p(E), _p(Ev), _p(parts)))
        return E, Ev, parts

    def action(self):
        a, b = np.zeros(self.C), np.zeros(self.C)
        self._ck(self.lib.pimc_action(self.h, _p(a), _p(b)))
        return a, b

    # ---- neighbour search ----
    def find_nn(self, chain, r, slice_, exception=0):
        r = _f(r)
        out = C.c_int64()
        self._ck(self.lib.pimc_find_nn(self.h, chain, _p(r), slice_, exception, C.byref(out)))
        return out.value

    def find_nns(self, chain, r, slice_, exception=0):
        r = _f(r)
        cap = 9 * self.N + 8
        out = np.zeros(cap, dtype=np.int64)
        cnt = C.c_int64()
        self._ck(self.lib.pimc_find_nns(self.h, chain, _p(r), slice_, exception, _pi(out), cap, C.byref(cnt)))
        return out[:cnt.value].copy()

    def update_nnbins(self):
        self._ck(self.lib.pimc_update_nnbins(self.h))

    # ---- explicit moves ----
    def reshape_linear_explicit(self, chain, n, j0, m, xi, u, commit=True):
        xi = _f(xi)
        wi, wu, acc = C.c_double(), C.c_double(), C.c_int32()
        rp = np.zeros((self.dim, m + 1))
        self._ck(self.lib.pimc_reshape_linear_explicit(self.h, chain, n, j0, m, _p(xi), u, int(commit), C.byref(wi), C.byref(wu), _p(rp), C.byref(acc)))
        return acc.value, wi.value, wu.value, rp

    def reshape_swap_explicit(self, chain, n1, n2, j0, m, xi1, xi2, u, commit=True):
        xi1, xi2 = _f(xi1), _f(xi2)
        wi, wu, acc = C.c_double(), C.c_double(), C.c_int32()
        self._ck(self.lib.pimc_reshape_swap_explicit(self.h, chain, n1, n2, j0, m, _p(xi1), _p(xi2), u, int(commit), C.byref(wi), C.byref(wu), C.byref(acc)))
        return acc.value, wi.value, wu.value

    def com_explicit(self, chain, n, d, u, polymer=False, commit=True):
        d = _f(np.resize(np.asarray(d, dtype=np.float64), 2))
        wi, wu, acc = C.c_double(), C.c_double(), C.c_int32()
        self._ck(self.lib.pimc_com_explicit(self.h, chain, n, int(polymer), _p(d), u, int(commit), C.byref(wi), C.byref(wu), C.byref(acc)))
        return acc.value, wi.value, wu.value

    def swap_weights(self, chain, n1, j0, m):
        w = np.zeros(self.N)
        self._ck(self.lib.pimc_swap_weights(self.h, chain, n1, j0, m, _p(w)))
        return w

    # ---- objects ----
    def update_create(self, kind, var0):
        i = C.c_int32()
        self._ck(self.lib.pimc_update_create(self.h, kind, float(var0), C.byref(i)))
        return i.value

    def update_configure(self, uid, vmin, vmax, minacc, maxacc, adj=10, rng=10000):
        self._ck(self.lib.pimc_update_configure(self.h, uid, float(vmin), float(vmax), minacc, maxacc, adj, rng))

    def update_get(self, uid, chain=-1):
        var, acc = C.c_double(), C.c_double()
        tries, tv, a, bm = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.pimc_update_get(self.h, uid, chain, C.byref(var), C.byref(tries), C.byref(tv), C.byref(acc), C.byref(a), C.byref(bm)))
        return dict(var=var.value, tries=tries.value, tries_var=tv.value, acc_window=acc.value, accepted=a.value, bead_moves=bm.value)

    def energy_create(self, cap):
        i = C.c_int32()
        self._ck(self.lib.pimc_energy_create(self.h, cap, C.byref(i)))
        return i.value

    def energy_read(self, eid, chain=-1, cap=None):
        n = C.c_int64()
        self._ck(self.lib.pimc_energy_read(self.h, eid, chain, None, None, 0, C.byref(n)))
        m = n.value if cap is None else min(n.value, cap)
        E, Ev = np.zeros(max(m, 1)), np.zeros(max(m, 1))
        self._ck(self.lib.pimc_energy_read(self.h, eid, chain, _p(E), _p(Ev), m, C.byref(n)))
        return E[:m], Ev[:m], n.value

    def energy_read_range(self, eid, start, count, chain=-1):
        """measurement block [start, start+count): returns (E, Ev, total measurements taken)"""
        n = C.c_int64()
        E, Ev = np.zeros(max(count, 1)), np.zeros(max(count, 1))
        self._ck(self.lib.pimc_energy_read_range(self.h, eid, chain, start, count, _p(E), _p(Ev), C.byref(n)))
        m = max(0, min(count, n.value - start))
        return E[:m], Ev[:m], n.value

    def energy_stats(self, eid):
        out = np.zeros((self.C, 5))
        self._ck(self.lib.pimc_energy_stats(self.h, eid, _p(out)))
        return out

    def density_create(self, nbins):
        i = C.c_int32()
        self._ck(self.lib.pimc_density_create(self.h, nbins, C.byref(i)))
        return i.value

    def density_measure(self, did):
        self._ck(self.lib.pimc_density_measure(self.h, did))

    def density_read(self, did, nbins):
        shape = (nbins,) * self.dim
        dens = np.zeros(int(np.prod(shape)))
        nd, b = C.c_int64(), C.c_double()
        self._ck(self.lib.pimc_density_read(self.h, did, _p(dens), C.byref(nd), C.byref(b)))
        return dens.reshape(shape, order="F"), nd.value, b.value

    def run(self, n, updates, energies=(), densities=(), sched=L.SCHED_FAITHFUL, paircorrs=(), windings=(), structures=()):
        """updates: [(every, update_id)] ; returns RunStats as dict"""
        nu = len(updates)
        ids = (C.c_int32 * nu)(*[u for _, u in updates])
        ev = (C.c_int64 * nu)(*[e for e, _ in updates])
        en = (C.c_int32 * max(1, len(energies)))(*energies)
        de = (C.c_int32 * max(1, len(densities)))(*densities)
        st = L.RunStats()
        if not paircorrs and not windings and not structures:
            self._ck(self.lib.pimc_run(self.h, n, ids, ev, nu, en, len(energies), de, len(densities), sched, C.byref(st)))
        else:
            pc = (C.c_int32 * max(1, len(paircorrs)))(*paircorrs)
            wi = (C.c_int32 * max(1, len(windings)))(*windings)
            sk = (C.c_int32 * max(1, len(structures)))(*structures)
            z = L.Measurements(C.cast(en, L.i32p), len(energies), C.cast(de, L.i32p), len(densities), C.cast(pc, L.i32p), len(paircorrs),
                               C.cast(wi, L.i32p), len(windings), C.cast(sk, L.i32p), len(structures))
            self._ck(self.lib.pimc_run_ex(self.h, n, ids, ev, nu, C.byref(z), sched, C.byref(st)))
        return {k: getattr(st, k) for k, _ in L.RunStats._fields_}

    # ---- estimators the reference lists as TODO (measurement.jl:125-127) ----
    def paircorr_create(self, nbins, rmax):
        i = C.c_int32()
        self._ck(self.lib.pimc_paircorr_create(self.h, nbins, float(rmax), C.byref(i)))
        return i.value

    def paircorr_measure(self, pid):
        self._ck(self.lib.pimc_paircorr_measure(self.h, pid))

    def paircorr_read(self, pid, nbins):
        hist, nd, b = np.zeros(nbins), C.c_int64(), C.c_double()
        self._ck(self.lib.pimc_paircorr_read(self.h, pid, _p(hist), C.byref(nd), C.byref(b)))
        return hist, nd.value, b.value

    def winding_create(self, cap):
        i = C.c_int32()
        self._ck(self.lib.pimc_winding_create(self.h, cap, C.byref(i)))
        return i.value

    def winding_now(self):
        W = np.zeros((self.C, self.dim))
        self._ck(self.lib.pimc_winding_now(self.h, _p(W)))
        return W

    def winding_read(self, wid, chain=-1):
        n = C.c_int64()
        self._ck(self.lib.pimc_winding_read(self.h, wid, chain, None, 0, C.byref(n)))
        out = np.zeros((max(1, n.value), self.dim)) if chain >= 0 else np.zeros(max(1, n.value))
        self._ck(self.lib.pimc_winding_read(self.h, wid, chain, _p(out), n.value, C.byref(n)))
        return out[:n.value], n.value

    def structure_create(self, kmax):
        i = C.c_int32()
        self._ck(self.lib.pimc_structure_create(self.h, int(kmax), C.byref(i)))
        return i.value

    def structure_measure(self, sid):
        self._ck(self.lib.pimc_structure_measure(self.h, sid))

    def structure_read(self, sid, kmax):
        """sums of |rho_k|^2 [kmax + 1][2 kmax + 1] (entry (a, b + kmax) for k = (pi / L)(a, b)) and ndata; S(k) = sums / (ndata * N)"""
        sums, nd, km = np.zeros((kmax + 1, 2 * kmax + 1)), C.c_int64(), C.c_int32()
        self._ck(self.lib.pimc_structure_read(self.h, sid, _p(sums), C.byref(nd), C.byref(km)))
        assert km.value == kmax
        return sums, nd.value

    def compressibility(self, sid):
        """(kappa_T, S(k_min)) from the smallest shell of the box"""
        k, s0 = C.c_double(), C.c_double()
        self._ck(self.lib.pimc_compressibility(self.h, sid, C.byref(k), C.byref(s0)))
        return k.value, s0.value


# ---- stateless device hooks ----
def distance(x1, x2, L_):
    x1, x2 = _f(x1), _f(x2)
    out = np.zeros_like(x1)
    L.check(L.load().pimc_distance(x1.size, _p(x1), _p(x2), L_, _p(out)))
    return out


def teleport(x, L_):
    x = _f(x)
    out = np.zeros_like(x)
    L.check(L.load().pimc_teleport(x.size, _p(x), L_, _p(out)))
    return out


def lnK(r1, r2, tau, lam, L_):
    r1, r2 = _f(r1), _f(r2)
    n, dim = r1.shape
    out = np.zeros(n)
    L.check(L.load().pimc_lnK(n, _p(r1), _p(r2), dim, tau, lam, L_, _p(out)))
    return out


def lnV(r1, r2, tau, pot):
    r1, r2 = _f(r1), _f(r2)
    n, dim = r1.shape
    out = np.zeros(n)
    L.check(L.load().pimc_lnV(n, _p(r1), _p(r2), dim, tau, C.byref(pot), _p(out)))
    return out


def potential_eval(r, pot):
    r = _f(r)
    n, dim = r.shape
    V, dV = np.zeros(n), np.zeros((n, dim))
    L.check(L.load().pimc_potential_eval(n, _p(r), dim, C.byref(pot), _p(V), _p(dV)))
    return V, dV


def levy_bridge(r, tau, L_, lam, xi):
    """r: (nb, dim, rows) column-major bridges (endpoints set); xi: (nb, rows-2, dim). Returns the bridged copy."""
    r = _f(r).copy()
    xi = _f(xi)
    nb, dim, rows = r.shape
    L.check(L.load().pimc_levy_bridge(_p(r), rows, dim, tau, L_, lam, _p(xi), nb))
    return r


def gauss_pairs(seed, chain, it, slot, kind, retry, bead0, n):
    g = np.zeros((n, 2))
    L.check(L.load().pimc_gauss_pairs(seed, chain, it, slot, kind, retry, bead0, n, _p(g)))
    return g
