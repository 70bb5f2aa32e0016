"""ctypes binding of the CPU oracle (oracle/libpimc_oracle.so).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_ROOT, "oracle", "libpimc_oracle.so")

POT_ZERO, POT_HARMONIC, POT_SIN2_1D, POT_LATTICE = 0, 1, 2, 3
DV_ZERO, DV_IDENTITY, DV_GRADIENT = 0, 1, 2
UPD_RESHAPE_LINEAR, UPD_RESHAPE_SWAP, UPD_SINGLE_COM, UPD_POLYMER_COM = 0, 1, 2, 3
SCHED_FAITHFUL, SCHED_SWEEP, SCHED_SWEEP_SEQ = 0, 1, 2
COMPAT_PAIR_BYVALUE, COMPAT_SWAP_SIGN, COMPAT_DENSITY_SHIFT, COMPAT_SWAP_STALE_LINK, COMPAT_ALL = 1, 2, 4, 8, 15
MAX_ANGLES = 32

f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)


class Potential(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dv_kind", C.c_int32), ("k", C.c_double), ("depth", C.c_double),
                ("scale", C.c_double), ("sgn", C.c_double), ("nang", C.c_int32), ("helical", C.c_int32),
                ("ang", C.c_double * MAX_ANGLES)]


class Config(C.Structure):
    _fields_ = [("dim", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("mu", C.c_double), ("lam", C.c_double),
                ("L", C.c_double), ("T", C.c_double), ("interactions", C.c_int32), ("g", C.c_double),
                ("r_a", C.c_double), ("Ncycle", C.c_int32), ("compat", C.c_int32), ("init", C.c_int32),
                ("seed", C.c_uint64), ("chain", C.c_uint32), ("pot", Potential), ("tab", f64p),
                ("tab_n", C.c_int32), ("tab_lo", C.c_double), ("tab_hi", C.c_double)]


def build(force=False):
    src = os.path.join(_ROOT, "oracle", "pimc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        d, i64, vp = C.c_double, C.c_int64, C.c_void_p
        sig = {
            "ora_distance": (d, [d, d, d]), "ora_teleport": (d, [d, d]),
            "ora_lnK": (d, [f64p, f64p, C.c_int, d, d, d]), "ora_prop_0": (d, [f64p, f64p, C.c_int, d, d, d]),
            "ora_potential_eval": (d, [C.POINTER(Potential), f64p, C.c_int]),
            "ora_potential_grad": (None, [C.POINTER(Potential), f64p, C.c_int, f64p]),
            "ora_lnV": (d, [f64p, f64p, C.c_int, d, C.POINTER(Potential)]),
            "ora_levy": (None, [f64p, C.c_int, C.c_int, d, d, d, f64p]),
            "ora_metropolis": (C.c_int, [d, d]),
            "ora_bin": (i64, [f64p, C.c_int, i64, d]),
            "ora_bin_neighbors": (None, [i64, i64, C.c_int, i64p]),
            "ora_pcycle": (i64, [i64, i64p, i64, i64]),
            "ora_adjust_step": (d, [d, d, d, d, d, d]),
            "ora_adjust_slices": (i64, [i64, i64, i64, d, d, d]),
            "ora_prop_rel0": (d, [f64p, f64p, C.c_int, d]),
            "ora_gauss_pair": (None, [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f64p, f64p]),
            "ora_uniform_pair": (None, [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f64p, f64p]),
            "ora_create": (vp, [C.POINTER(Config)]), "ora_destroy": (None, [vp]), "ora_last_error": (C.c_char_p, []),
            "ora_get_paths": (None, [vp, f64p, f64p, i64p, i64p]), "ora_set_paths": (None, [vp, f64p, i64p]),
            "ora_get_scalars": (None, [vp, f64p, i64p]), "ora_set_iter": (None, [vp, C.c_uint64]), "ora_set_ctr": (None, [vp, i64]),
            "ora_update_nnbins": (None, [vp]), "ora_subcycle": (i64, [vp, i64, i64p]), "ora_cycle_findprev": (i64, [vp, i64]),
            "ora_find_nn": (i64, [vp, f64p, i64, i64p, C.c_int]),
            "ora_find_nns_pos": (i64, [vp, f64p, i64, i64p, C.c_int, i64p]),
            "ora_find_nns_idx": (i64, [vp, i64, i64, i64p, C.c_int, i64p]),
            "ora_nn_cell": (i64, [vp, i64, i64, i64p]),
            "ora_lnU": (d, [vp, f64p, f64p]),
            "ora_action_links": (d, [vp]), "ora_action_links_recomputed": (d, [vp]), "ora_action_pairs": (d, [vp]),
            "ora_update_create": (vp, [vp, C.c_int, d]),
            "ora_update_configure": (None, [vp, d, d, d, d, i64, i64]),
            "ora_update_destroy": (None, [vp]),
            "ora_update_get": (None, [vp, f64p, i64p, i64p, f64p, i64p, i64p]),
            "ora_update_call": (C.c_int, [vp, vp, C.c_uint32, i64, i64]),
            "ora_reshape_linear_explicit": (C.c_int, [vp, i64, i64, i64, f64p, d, C.c_int, f64p, f64p, f64p]),
            "ora_reshape_swap_explicit": (C.c_int, [vp, i64, i64, i64, i64, f64p, f64p, d, C.c_int, f64p, f64p]),
            "ora_com_explicit": (C.c_int, [vp, i64, C.c_int, f64p, d, C.c_int, f64p, f64p]),
            "ora_swap_weights": (None, [vp, i64, i64, i64, f64p]),
            "ora_energy_create": (vp, [i64]), "ora_energy_destroy": (None, [vp]),
            "ora_energy_read": (i64, [vp, f64p, f64p, i64]), "ora_energy_now": (None, [vp, f64p, f64p, f64p]),
            "ora_density_create": (vp, [vp, i64]), "ora_density_destroy": (None, [vp]),
            "ora_density_measure": (None, [vp, vp]), "ora_density_read": (i64, [vp, f64p, f64p]),
            "ora_paircorr_create": (vp, [vp, i64, d]), "ora_paircorr_destroy": (None, [vp]),
            "ora_paircorr_measure": (None, [vp, vp]), "ora_paircorr_read": (i64, [vp, f64p, f64p]),
            "ora_winding_now": (None, [vp, f64p]),
            "ora_structure_now": (None, [vp, C.c_int, f64p]),
            "ora_run": (C.c_int, [vp, i64, C.POINTER(vp), i64p, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(f64p)


def _pi(a):
    return a.ctypes.data_as(i64p)


def make_potential(kind="zero", dv="zero", k=1.0, depth=0.0, scale=1.0, sgn=1.0, angles=(), helical=False):
    p = Potential()
    p.kind = {"zero": 0, "harmonic": 1, "sin2_1d": 2, "lattice": 3}[kind]
    p.dv_kind = {"zero": 0, "identity": 1, "gradient": 2}[dv]
    p.k, p.depth, p.scale, p.sgn = k, depth, scale, sgn
    p.nang, p.helical = len(angles), int(helical)
    for i, a in enumerate(angles):
        p.ang[i] = a
    return p


class Update:
    def __init__(self, system, kind, var0):
        self.kind = kind
        self.h = lib().ora_update_create(system.h, kind, float(var0))

    def configure(self, vmin, vmax, minacc, maxacc, adj=10, rng=10000):
        lib().ora_update_configure(self.h, float(vmin), float(vmax), minacc, maxacc, adj, rng)

    def get(self):
        var, acc = C.c_double(), C.c_double()
        tries, tv, a, bm = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        lib().ora_update_get(self.h, C.byref(var), C.byref(tries), C.byref(tv), C.byref(acc), C.byref(a), C.byref(bm))
        return dict(var=var.value, tries=tries.value, tries_var=tv.value, acc_window=acc.value, accepted=a.value, bead_moves=bm.value)

    def __del__(self):
        try:
            lib().ora_update_destroy(self.h)
        except Exception:
            pass


class Energy:
    def __init__(self, cap=20000):
        self.cap = cap
        self.h = lib().ora_energy_create(cap)

    def read(self):
        E, Ev = np.zeros(self.cap), np.zeros(self.cap)
        n = lib().ora_energy_read(self.h, _p(E), _p(Ev), self.cap)
        n = min(n, self.cap)
        return E[:n], Ev[:n]

    def __del__(self):
        try:
            lib().ora_energy_destroy(self.h)
        except Exception:
            pass


class Density:
    def __init__(self, system, nbins=500):
        self.nbins, self.dim = nbins, system.dim
        self.h = lib().ora_density_create(system.h, nbins)

    def measure(self, system):
        lib().ora_density_measure(self.h, system.h)

    def read(self):
        shape = (self.nbins,) * self.dim
        dens = np.zeros(int(np.prod(shape)))
        b = C.c_double()
        nd = lib().ora_density_read(self.h, _p(dens), C.byref(b))
        return dens.reshape(shape, order="F"), nd, b.value

    def __del__(self):
        try:
            lib().ora_density_destroy(self.h)
        except Exception:
            pass


class PairCorrelation:
    def __init__(self, system, nbins, rmax):
        self.nbins = nbins
        self.h = lib().ora_paircorr_create(system.h, nbins, float(rmax))

    def measure(self, system):
        lib().ora_paircorr_measure(self.h, system.h)

    def read(self):
        hist, b = np.zeros(self.nbins), C.c_double()
        nd = lib().ora_paircorr_read(self.h, _p(hist), C.byref(b))
        return hist, nd, b.value

    def __del__(self):
        try:
            lib().ora_paircorr_destroy(self.h)
        except Exception:
            pass


def winding_now(system):
    W = np.zeros(system.dim)
    lib().ora_winding_now(system.h, _p(W))
    return W


def structure_now(system, kmax):
    """sum over the slices of |rho_k|^2 of the current configuration, [kmax + 1][2 kmax + 1] (a, b + kmax)"""
    S = np.zeros((kmax + 1, 2 * kmax + 1))
    lib().ora_structure_now(system.h, int(kmax), _p(S))
    return S


class System:
    """One reference `System` (src/system.jl:93-168) held by the oracle."""

    def __init__(self, pot=None, dim=2, M=100, N=2, mu=0.0, L=4.0, T=1.0, lam=1.0, interactions=False, g=0.0,
                 r_a=0.0, Ncycle=10, compat=COMPAT_ALL, init=True, seed=0x5EEDB200, chain=0, tab=None, tab_lo=0.0, tab_hi=1.0):
        c = Config()
        c.dim, c.M, c.N, c.mu, c.lam, c.L, c.T = dim, M, N, mu, lam, L, T
        c.interactions, c.g, c.r_a, c.Ncycle, c.compat, c.init = int(interactions), g, r_a, Ncycle, compat, int(init)
        c.seed, c.chain = seed, chain
        c.pot = pot if pot is not None else make_potential()
        self._tab = None
        if tab is not None:
            self._tab = np.asfortranarray(tab, dtype=np.float64)
            c.tab = _p(self._tab)
            c.tab_n, c.tab_lo, c.tab_hi = self._tab.shape[0], tab_lo, tab_hi
        self.cfg = c
        self.dim, self.M, self.N, self.L = dim, M, N, L
        self.h = lib().ora_create(C.byref(c))
        if not self.h:
            raise RuntimeError(lib().ora_last_error().decode())
        sc = self.scalars()
        self.beta, self.tau, self.a, self.nbins = sc["beta"], sc["tau"], sc["a"], sc["nbins"]

    def __del__(self):
        try:
            lib().ora_destroy(self.h)
        except Exception:
            pass

    def scalars(self):
        out = np.zeros(5)
        io = np.zeros(5, dtype=np.int64)
        lib().ora_get_scalars(self.h, _p(out), _pi(io))
        return dict(beta=out[0], tau=out[1], vol=out[2], a=out[3], r_a=out[4], nbins=int(io[0]), N_MC=int(io[1]),
                    Nctr=int(io[2]), ctr=int(io[3]), iter=int(io[4]))

    def paths(self):
        r = np.zeros((self.N, self.dim, self.M))
        V = np.zeros((self.N, self.M))
        bins = np.zeros((self.N, self.M), dtype=np.int64)
        nxt = np.zeros(self.N, dtype=np.int64)
        lib().ora_get_paths(self.h, _p(r), _p(V), _pi(bins), _pi(nxt))
        return r, V, bins, nxt

    def set_paths(self, r, nxt=None):
        r = np.ascontiguousarray(r, dtype=np.float64)
        assert r.shape == (self.N, self.dim, self.M)
        if nxt is None:
            lib().ora_set_paths(self.h, _p(r), None)
        else:
            nxt = np.ascontiguousarray(nxt, dtype=np.int64)
            lib().ora_set_paths(self.h, _p(r), _pi(nxt))

    def energy_now(self):
        E, Ev = C.c_double(), C.c_double()
        parts = np.zeros(3)
        lib().ora_energy_now(self.h, C.byref(E), C.byref(Ev), _p(parts))
        return E.value, Ev.value, parts

    def run(self, n, updates, energies=(), densities=(), sched=SCHED_FAITHFUL):
        """updates: list of (every, Update)"""
        nu = len(updates)
        U = (C.c_void_p * nu)(*[u.h for _, u in updates])
        ev = np.array([e for e, _ in updates], dtype=np.int64)
        En = (C.c_void_p * max(1, len(energies)))(*[e.h for e in energies])
        De = (C.c_void_p * max(1, len(densities)))(*[d.h for d in densities])
        rc = lib().ora_run(self.h, n, U, _pi(ev), nu, En, len(energies), De, len(densities), sched)
        if rc != 0:
            raise RuntimeError(lib().ora_last_error().decode())
