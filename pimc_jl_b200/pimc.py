"""Host-side mirror of the reference's Julia interface (module `Pimc`, src/Pimc.jl:15-26) over the C ABI.

Same names, argument meaning and error behaviour as the reference, Python spelling for `!` (run_b = run!, levy_b = levy!):

    System(potential; dV, dim, M, N, mu, L, T, lam, interactions, propint, g, r_a, length_measurement_cycle)  src/system.jl:129-167
    SingleCenterOfMass / PolymerCenterOfMass / ReshapeLinear / ReshapeSwapLinear                          src/updates/*.jl
    Energy / Density                                                                                       src/measurement.jl
    run_b(s, n, updates, Zmeasurements=[...])                                                              src/simulation.jl:29-42

A `System` holds `chains` independent replicas (default 1 = the reference); with torch.distributed initialised the chains
are sharded over the ranks (one process per GPU) and estimator read-outs are all-reduced.  No CPU fallback exists.
"""
import math
import os
import numpy as np
from . import _lib as _L
from . import engine as _eng
from .engine import Engine
import ctypes as _C


def build_prop_int(L, g0, tau, delta=600):
    """build_prop_int(L, g0, tau) (src/propagator.jl:79-89): the `propint` argument of System(...; interactions=true) -- the sampled term
    table of prop_rel_interpolate_terms, built by the library's host code (pimc_build_prop_table, csrc/pimc_propint.cu)"""
    tab = np.zeros((delta, delta), order="F")
    lo, hi = _C.c_double(), _C.c_double()
    _L.check(_L.load().pimc_build_prop_table(float(L), float(g0), float(tau), int(delta), tab.ctypes.data_as(_L.f64p), _C.byref(lo), _C.byref(hi)))
    return PropInt(tab=tab, lo=lo.value, hi=hi.value, g0=float(g0), tau=float(tau))


class PropInt(dict):
    """what build_prop_int returns: the table, callable like the reference's closure prop_int(r1_rel, r2_rel, tau)"""

    def __call__(self, r1_rel, r2_rel, tau):
        r1, r2 = np.atleast_1d(np.asarray(r1_rel, dtype=np.float64)), np.atleast_1d(np.asarray(r2_rel, dtype=np.float64))
        out = _C.c_double()
        _L.check(_L.load().pimc_prop_int(self["tab"].ctypes.data_as(_L.f64p), self["tab"].shape[0], self["lo"], self["hi"],
                                         r1.ctypes.data_as(_L.f64p), r2.ctypes.data_as(_L.f64p), len(r1), float(tau), _C.byref(out)))
        return out.value


def determine_nnrange(propint, tau, a, b):
    """determine_nnrange(propint, tau, a, b) (src/system.jl:10-15)"""
    out = _C.c_double()
    tab = np.asfortranarray(propint["tab"], dtype=np.float64)
    rc = _L.load().pimc_determine_nnrange(tab.ctypes.data_as(_L.f64p), tab.shape[0], propint["lo"], propint["hi"], float(tau), float(a), float(b), _C.byref(out))
    if rc != 0:
        raise ValueError("determine_nnrange: no sign change of propint - 0.999 on (r_min, b) (Roots.find_zero would throw)")
    return out.value

L25 = [2.214297435588181, 0.9272952180016122, -0.6435011087932844, 0.6435011087932844, -2.498091544796509, 3.141592653589793,
       2.498091544796509, 0, 1.5707963267948966, -2.2142974355881813, -1.5707963267948968, -0.9272952180016123]
L65 = [0.5191461142465229, 1.695151321341658, -0.12435499454676144, -1.4464413322481353, 2.0899424410414196, 2.62244653934327,
       0.12435499454676144, -1.0516502125483738, -1.695151321341658, 1.446441332248135, -2.0899424410414196, -3.017237659043032,
       -0.5191461142465229, 3.017237659043032, -2.62244653934327, 1.0516502125483738]


# ---- potential descriptors (the closures of examples/*.jl and examples/tools/potentialtools.jl) ----
class PotentialSpec(dict):
    """keyword arguments of _lib.make_potential; callable like the Julia closure (evaluated on the device)."""

    def __call__(self, r):
        r = np.atleast_2d(np.asarray(r, dtype=np.float64))
        return _eng.potential_eval(r, _L.make_potential(**self))[0]


def zero_potential():            # _ -> 0.0
    return PotentialSpec(kind="zero")


def harmonic(k=1.0):             # (r) -> 0.5*(r[1]^2+r[2]^2)
    return PotentialSpec(kind="harmonic", k=k)


def sin2_1d(depth, scale):       # test/testsystem.jl:13
    return PotentialSpec(kind="sin2_1d", depth=depth, scale=scale)


def generate_V(scale, depth, sym, attractive=True):   # examples/tools/potentialtools.jl:25-39
    if sym == "harmonic":
        return harmonic()
    ang = {"l65": L65, "l25": L25, "cubic": [2 * math.pi * k / 4 for k in range(4)]}.get(sym)
    if ang is None:
        raise ValueError("error()")
    return PotentialSpec(kind="lattice", depth=depth, scale=scale, sgn=-1.0 if attractive else 1.0, angles=ang)


def lattice(angles, scale, depth, helical=False):     # potentialtools.jl:18-24
    return PotentialSpec(kind="lattice", depth=depth, scale=scale, sgn=1.0, angles=list(angles), helical=helical)


# ---- multi-rank plumbing (one process per GPU) ----
def _dist():
    try:
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() else None
    except Exception:
        return None


def shard(chains_total, rank, world):
    """contiguous chain ranges; returns (offset, count) of `rank`"""
    base, rem = divmod(chains_total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def allreduce_sum(arr):
    """sum a numpy array over the ranks (NCCL when CUDA tensors are required by the backend, gloo on CPU)"""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return arr
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def chain_mean_over_ranks(local_mean, local_chains):
    """chain-weighted mean of per-rank block means"""
    num = allreduce_sum(np.asarray(local_mean, dtype=np.float64) * local_chains)
    den = allreduce_sum(np.array([float(local_chains)]))
    return num / den[0]


# ---- state ----
class Particle:
    """view of s.world[n] (src/system.jl:1-6): r is M x dim, V and bins length M, next 1-based."""

    def __init__(self, r, V, bins, nxt):
        self.r, self.V, self.bins, self.next = r, V, bins, nxt


class _Counter:
    def __init__(self, upd, which):
        self._u, self._w = upd, which

    @property
    def tries(self):
        return self._u._get()["tries" if self._w == "counter" else "tries_var"]

    @property
    def queue(self):
        return self._u  # acceptance(u.counter_var.queue) -> acceptance() below understands update objects


class _Var:
    def __init__(self, upd):
        self._u = upd

    @property
    def size(self):
        return self._u._get()["var"]

    @property
    def m(self):
        return int(round(self._u._get()["var"]))


class _Update:
    KIND = None

    def __init__(self, s, var0, vmin, vmax, minacc, maxacc, adj, range_):
        self.s = s
        self.id = s.engine.update_create(self.KIND, var0)
        s.engine.update_configure(self.id, vmin, vmax, minacc, maxacc, adj, range_)
        self.counter, self.counter_var, self.var = _Counter(self, "counter"), _Counter(self, "counter_var"), _Var(self)

    def _get(self, chain=0):
        return self.s.engine.update_get(self.id, chain)

    def __call__(self, s):
        """functor call f(s)::Bool on every chain (one faithful iteration of this update); returns chain 0's outcome"""
        before = self._get()["accepted"]
        s.engine.run(1, [(1, self.id)], sched=_L.SCHED_FAITHFUL)
        return self._get()["accepted"] > before


class ReshapeLinear(_Update):      # src/updates/reshape.jl:7-29
    KIND = _L.UPD_RESHAPE_LINEAR

    def __init__(self, s, slices, minslices=2, maxslices=None, minacc=0.6, maxacc=0.8, adj=10, range=10_000):
        maxslices = s.M - 2 if maxslices is None else maxslices
        super().__init__(s, min(slices, maxslices), minslices, maxslices, minacc, maxacc, adj, range)


class ReshapeSwapLinear(ReshapeLinear):  # src/updates/reshape.jl:99-121
    KIND = _L.UPD_RESHAPE_SWAP


class SingleCenterOfMass(_Update):  # src/updates/com.jl:112-133
    KIND = _L.UPD_SINGLE_COM

    def __init__(self, s, step, minstep=1e-1, maxstep=None, minacc=0.4, maxacc=0.6, adj=10, range=10_000):
        super().__init__(s, float(step), minstep, s.L / 2 if maxstep is None else maxstep, minacc, maxacc, adj, range)


class PolymerCenterOfMass(SingleCenterOfMass):  # src/updates/com.jl:7-28 (worms = 0 reading)
    KIND = _L.UPD_POLYMER_COM


def acceptance(q):                  # src/simulation.jl:1
    return q._get()["acc_window"]


class Energy:                       # src/measurement.jl:78-122
    def __init__(self, s, n=20_000):
        self.s, self.n = s, n
        self.id = s.engine.energy_create(n)

    def _read(self):
        E, Ev, cnt = self.s.engine.energy_read(self.id, -1)
        if cnt > self.n:
            raise IndexError("Energy: the pre-sized vector is full (reference: findfirst(ismissing, ...) is nothing)")
        if _dist() is not None and not self.s.library_comm:     # (with the library's communicator the values are global already)
            E = chain_mean_over_ranks(E, self.s.engine.C)
            Ev = chain_mean_over_ranks(Ev, self.s.engine.C)
        return E, Ev

    @property
    def energy(self):
        """Dict N -> series (chain mean per measurement), like mea.energy[N] with the missings skipped"""
        return {self.s.N: self._read()[0]}

    @property
    def energy_virial(self):
        return {self.s.N: self._read()[1]}

    def chain_stats(self):
        """per-chain (n, mean E, mean Ev) -- independent chains give the error bar"""
        a = self.s.engine.energy_stats(self.id)
        return a[:, 0], a[:, 1] / a[:, 0], a[:, 3] / a[:, 0]

    def __call__(self, s):          # functor call outside run!: measure now, every chain; returns (E, Ev) arrays
        return s.engine.energy_now()[:2]


class Density:                      # src/measurement.jl:31-55
    def __init__(self, s, nbins=500):
        self.s, self.nbins = s, nbins
        self.bin = (2 * s.L) / nbins
        self.id = s.engine.density_create(nbins)

    def _read(self):
        d, nd, b = self.s.engine.density_read(self.id, self.nbins)
        if _dist() is not None and not self.s.library_comm:
            d = allreduce_sum(d)
            nd = int(allreduce_sum(np.array([float(nd)]))[0])
        return d, nd

    @property
    def dens(self):
        return self._read()[0]

    @property
    def ndata(self):
        return self._read()[1]

    def __call__(self, s):
        s.engine.density_measure(self.id)


class PairCorrelation:              # `#TODO radial distribution` (src/measurement.jl:125), in the style of Density
    """g(r) from equal-time pair distances: PairCorrelation(s; nbins=200, rmax=s.L); .hist (raw pair counts), .ndata, .g (normalised)"""

    def __init__(self, s, nbins=200, rmax=None):
        self.s, self.nbins, self.rmax = s, nbins, float(s.L if rmax is None else rmax)
        self.bin = self.rmax / nbins
        self.id = s.engine.paircorr_create(nbins, self.rmax)

    def _read(self):
        h, nd, _ = self.s.engine.paircorr_read(self.id, self.nbins)   # global over the ranks when the library communicator is attached
        return h, nd

    @property
    def hist(self):
        return self._read()[0]

    @property
    def ndata(self):
        return self._read()[1]

    @property
    def r(self):
        return (np.arange(self.nbins) + 0.5) * self.bin

    @property
    def g(self):
        h, nd = self._read()
        edges = np.arange(self.nbins + 1) * self.bin
        shell = np.pi * (edges[1:] ** 2 - edges[:-1] ** 2) if self.s.dim == 2 else 2 * self.bin * np.ones(self.nbins)
        ideal = nd * self.s.N * (self.s.N - 1) / 2.0 * shell / self.s.vol
        return h / np.maximum(ideal, 1e-300)

    def __call__(self, s):
        s.engine.paircorr_measure(self.id)


class Winding:                      # `#TODO Superfluid Fraction` (src/measurement.jl:126), in the style of Energy
    """winding number per measurement: .W2 (chain-mean of W^2 per measurement), .series(chain), .superfluid_fraction()"""

    def __init__(self, s, n=20_000):
        self.s, self.n = s, n
        self.id = s.engine.winding_create(n)

    @property
    def W2(self):
        return self.s.engine.winding_read(self.id, -1)[0]

    def series(self, chain=0):
        return self.s.engine.winding_read(self.id, chain)[0]

    def superfluid_fraction(self):
        """rho_s / rho = <W^2> (2L)^2 / (2 dim lambda beta N)  (Pollock & Ceperley, PRB 36, 8343)"""
        w2 = self.W2
        return float(np.mean(w2)) * (2 * self.s.L) ** 2 / (2 * self.s.dim * self.s.lam * self.s.beta * self.s.N) if len(w2) else float("nan")

    def __call__(self, s):
        return s.engine.winding_now()


class StructureFactor:              # `#TODO Compressibilty` (src/measurement.jl:127), in the style of Density
    """static structure factor on the box's wave vectors k = (pi / L)(a, b), a = 0..kmax, |b| <= kmax:
    StructureFactor(s; kmax=4); .S ([kmax + 1][2 kmax + 1], NaN outside the half plane of independent vectors), .k (|k| of every entry),
    .ndata, .compressibility() (kappa_T = beta S(k_min) / rho, the long-wavelength limit on the smallest shell)"""

    def __init__(self, s, kmax=4):
        self.s, self.kmax = s, int(kmax)
        self.id = s.engine.structure_create(self.kmax)

    def _read(self):
        return self.s.engine.structure_read(self.id, self.kmax)   # global over the ranks when the library communicator is attached

    @property
    def ndata(self):
        return self._read()[1]

    @property
    def S(self):
        sums, nd = self._read()
        a, b = np.meshgrid(np.arange(self.kmax + 1), np.arange(-self.kmax, self.kmax + 1), indexing="ij")
        half = ((a > 0) | (b > 0)) if self.s.dim > 1 else ((b == 0) & (a > 0))
        return np.where(half, sums / max(1, nd * self.s.N), np.nan)

    @property
    def k(self):
        a, b = np.meshgrid(np.arange(self.kmax + 1), np.arange(-self.kmax, self.kmax + 1), indexing="ij")
        return np.pi / self.s.L * np.hypot(a, b if self.s.dim > 1 else 0 * b)

    def compressibility(self):
        return self.s.engine.compressibility(self.id)[0]

    def __call__(self, s):
        s.engine.structure_measure(self.id)


class System:                       # src/system.jl:93-168
    def __init__(self, potential, dV="zero", dim=2, M=100, N=2, mu=0.0, L=4.0, T=1.0, lam=1.0, interactions=False, propint=None,
                 g=0.0, r_a=0.0, length_measurement_cycle=10, measure_scheme="c", chains=None, seed=None, compat=_L.COMPAT_ALL,
                 schedule=None, device=-1):
        if isinstance(potential, PotentialSpec):
            spec = dict(potential)
        elif callable(potential):
            raise TypeError("arbitrary closures cannot run inside a CUDA kernel: pass a potential descriptor "
                            "(zero_potential(), harmonic(), sin2_1d(), generate_V(), lattice())")
        else:
            spec = dict(potential)
        if not isinstance(dV, str) or dV not in ("zero", "identity", "gradient"):
            raise TypeError("dV must be 'zero' (the reference's default `zero`), 'identity' (dV = identity, as the example scripts pass) or "
                            "'gradient' (the analytic gradient of the potential descriptor); arbitrary callables cannot run inside a CUDA kernel")
        spec["dv"] = dV
        chains = int(os.environ.get("PIMC_CHAINS", "1")) if chains is None else chains
        seed = int(os.environ.get("PIMC_SEED", str(0x5EEDB200))) if seed is None else seed
        self.schedule = schedule or os.environ.get("PIMC_SCHED", "faithful")
        dist = _dist()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        off, cnt = shard(chains, rank, world)
        tab = tab_lo = tab_hi = None
        if interactions:
            if propint is None or not isinstance(propint, dict):
                raise TypeError("interactions=True needs propint = dict(tab=..., lo=..., hi=...) (the sampled term table of build_prop_int)")
            tab, tab_lo, tab_hi = propint["tab"], propint["lo"], propint["hi"]
            # r_a == 0: init_int (src/system.jl:29-31) takes the cut-off from the propagator itself; pimc_create does (pimc_determine_nnrange)
        self.engine = Engine(_L.make_potential(**spec), dim=dim, M=M, N=N, chains=cnt, chain_offset=off, mu=mu, L_=L, T=T, lam=lam,
                             interactions=interactions, g=g, r_a=r_a, Ncycle=length_measurement_cycle, compat=compat, seed=seed,
                             tab=tab, tab_lo=tab_lo or 0.0, tab_hi=tab_hi or 1.0, device=device)
        e = self.engine
        # multi-GPU: with an NCCL process group the library's own communicator is attached (one rank per process), so that the estimator
        # read-outs below are reduced inside the library (side stream, no host bounce); other backends (gloo on CPU hosts) keep the host reduction
        self.library_comm = False
        if dist is not None and world > 1 and dist.get_backend() == "nccl":
            ids = [_eng.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            e.comm_init(world, rank, ids[0])
            self.library_comm = True
        self.dim, self.M, self.N, self.Ninit, self.mu, self.lam, self.L = dim, M, N, N, mu, lam, L
        self.beta, self.tau, self.vol, self.a, self.nbins = e.beta, e.tau, (2 * L) ** dim, e.a, e.nbins
        self.Ncycle, self.measure_scheme, self.chains = length_measurement_cycle, measure_scheme, chains
        self.V = PotentialSpec(spec)

    @property
    def N_MC(self):
        return {self.N: self.engine.scalars()["N_MC"]}

    @property
    def Nctr(self):
        return {self.N: self.engine.scalars()["Nctr"]}

    @property
    def ctr(self):
        return self.engine.scalars()["ctr"]

    def world_of(self, chain=0):
        r, V, bins, nxt = self.engine.paths(chain, 1)
        return [Particle(np.ascontiguousarray(r[0, n].T), V[0, n], bins[0, n], int(nxt[0, n])) for n in range(self.N)]

    @property
    def world(self):
        return self.world_of(0)

    def lnV(self, x1, x2):
        return _eng.lnV(np.atleast_2d(x1), np.atleast_2d(x2), self.tau, _L.make_potential(**self.V))[0]

    def lnK(self, x1, x2, tau):
        return _eng.lnK(np.atleast_2d(x1), np.atleast_2d(x2), self.lam, tau, self.L)[0]  # (lambda, tau) slot order of system.jl:163


def run_b(s, n, updates, Zmeasurements=()):
    """run!(s, n, updates; Zmeasurements) -- src/simulation.jl:29-42, on every chain."""
    en = [m.id for m in Zmeasurements if isinstance(m, Energy)]
    de = [m.id for m in Zmeasurements if isinstance(m, Density)]
    pc = [m.id for m in Zmeasurements if isinstance(m, PairCorrelation)]
    wi = [m.id for m in Zmeasurements if isinstance(m, Winding)]
    sk = [m.id for m in Zmeasurements if isinstance(m, StructureFactor)]
    sched = _L.SCHED_SWEEP if s.schedule == "sweep" else _L.SCHED_FAITHFUL
    return s.engine.run(n, [(every, u.id) for every, u in updates], energies=en, densities=de, sched=sched, paircorrs=pc, windings=wi, structures=sk)


def apply_b(s, f):
    """apply!(s, f::UpdateC) -- src/simulation.jl:19-27"""
    return f(s)


# ---- debug exports of src/Pimc.jl:18 ----
def distance(x1, x2, L_):
    return float(_eng.distance(np.array([x1]), np.array([x2]), L_)[0])


def teleport(x, L_):
    return float(_eng.teleport(np.array([x]), L_)[0])


def levy_b(r, tau, L_, lam, xi):
    """levy!(r', tau, L, lam) with the Gaussians supplied: r is rows x dim (first and last row fixed), xi (rows-2) x dim"""
    r = np.asarray(r, dtype=np.float64)
    out = _eng.levy_bridge(np.ascontiguousarray(r.T)[None], tau, L_, lam, np.asarray(xi, dtype=np.float64)[None])
    r[...] = out[0].T
    return r


def bin(r, nbins, L_):                                   # src/nearest_neighbours.jl:26-33
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    ib = np.clip(np.floor((r + L_) / (2 * L_ / nbins)).astype(np.int64), 0, nbins - 1)
    return int(ib[0] + 1 if len(ib) == 1 else ib[0] + nbins * ib[1] + 1)


def subcycle(world, n):                                  # src/updates/helper.jl:64-85
    cyc, i = [n], n
    while world[i - 1].next not in (0, n) and len(cyc) <= len(world):
        i = world[i - 1].next
        cyc.append(i)
    return len(cyc), cyc


def pcycle(j, pol, Npol, M):                             # src/updates/helper.jl:113-115
    return pol[(math.floor((j - 1) / M)) % Npol]


def update_nnbins_b(s):
    s.engine.update_nnbins()
