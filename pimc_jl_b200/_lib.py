"""ctypes binding of libpimc_b200.so (include/pimc_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
calls raise -- nothing in this package routes through oracle/ or any other CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PIMC_B200_SO") or os.path.join(_HERE, "libpimc_b200.so")  # override only for kernel-variant experiments
HEADER = os.path.join(os.path.dirname(_HERE), "include", "pimc_b200.h")

MAX_ANGLES = 32
POT_ZERO, POT_HARMONIC, POT_SIN2_1D, POT_LATTICE = 0, 1, 2, 3
DV_ZERO, DV_IDENTITY, DV_GRADIENT = 0, 1, 2
UPD_RESHAPE_LINEAR, UPD_RESHAPE_SWAP, UPD_SINGLE_COM, UPD_POLYMER_COM = 0, 1, 2, 3
SCHED_FAITHFUL, SCHED_SWEEP = 0, 1
OPT_SWEEP_IMPL = 1
OPT_FAITHFUL_IMPL = 2
OPT_FUSE_ENERGY = 3
OPT_ISWEEP = 4
COMPAT_PAIR_BYVALUE, COMPAT_SWAP_SIGN, COMPAT_DENSITY_SHIFT, COMPAT_SWAP_STALE_LINK, COMPAT_ALL = 1, 2, 4, 8, 15

f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


class PimcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpimc_b200 error {code}: {msg}")
        self.code = code


class Potential(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dv_kind", C.c_int32), ("k", C.c_double), ("depth", C.c_double),
                ("scale", C.c_double), ("sgn", C.c_double), ("nang", C.c_int32), ("helical", C.c_int32),
                ("ang", C.c_double * MAX_ANGLES)]


class Config(C.Structure):
    _fields_ = [("dim", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("chains", C.c_int32),
                ("chain_offset", C.c_uint32), ("mu", C.c_double), ("lam", C.c_double), ("L", C.c_double),
                ("T", C.c_double), ("interactions", C.c_int32), ("g", C.c_double), ("r_a", C.c_double),
                ("Ncycle", C.c_int32), ("compat", C.c_int32), ("init", C.c_int32), ("seed", C.c_uint64),
                ("pot", Potential), ("tab", f64p), ("tab_n", C.c_int32), ("tab_lo", C.c_double),
                ("tab_hi", C.c_double), ("device", C.c_int32)]


class RunStats(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("proposals", C.c_int64), ("accepted", C.c_int64),
                ("bead_moves", C.c_int64), ("measurements", C.c_int64), ("launches", C.c_int64),
                ("kernel_ms", C.c_double)]


class Measurements(C.Structure):
    _fields_ = [("energy_ids", i32p), ("nenergy", C.c_int32), ("density_ids", i32p), ("ndensity", C.c_int32),
                ("paircorr_ids", i32p), ("npaircorr", C.c_int32), ("winding_ids", i32p), ("nwinding", C.c_int32),
                ("structure_ids", i32p), ("nstructure", C.c_int32)]


_vp, _d, _i32, _i64, _u32, _u64 = C.c_void_p, C.c_double, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
_PP = C.POINTER(Potential)

# every symbol include/pimc_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "pimc_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "pimc_destroy": (None, [_vp]),
    "pimc_last_error": (C.c_char_p, [_vp]),
    "pimc_version": (C.c_int, []),
    "pimc_set_stream": (C.c_int, [_vp, _vp]),
    "pimc_set_option": (C.c_int, [_vp, _i32, _i64]),
    "pimc_launch_count": (C.c_int64, []),
    "pimc_measure_fp64_peak": (C.c_int, [f64p]),
    "pimc_get_paths": (C.c_int, [_vp, _i32, _i32, f64p, f64p, i64p, i64p]),
    "pimc_set_paths": (C.c_int, [_vp, _i32, _i32, f64p, i64p]),
    "pimc_get_scalars": (C.c_int, [_vp, f64p, i64p]),
    "pimc_set_iter": (C.c_int, [_vp, _u64]),
    "pimc_state_size": (C.c_int, [_vp, i64p]),
    "pimc_get_state": (C.c_int, [_vp, _vp, _i64]),
    "pimc_set_state": (C.c_int, [_vp, _vp, _i64]),
    "pimc_distance": (C.c_int, [_i64, f64p, f64p, _d, f64p]),
    "pimc_teleport": (C.c_int, [_i64, f64p, _d, f64p]),
    "pimc_lnK": (C.c_int, [_i64, f64p, f64p, _i32, _d, _d, _d, f64p]),
    "pimc_lnV": (C.c_int, [_i64, f64p, f64p, _i32, _d, _PP, f64p]),
    "pimc_potential_eval": (C.c_int, [_i64, f64p, _i32, _PP, f64p, f64p]),
    "pimc_levy_bridge": (C.c_int, [f64p, _i32, _i32, _d, _d, _d, f64p, _i64]),
    "pimc_gauss_pairs": (C.c_int, [_u64, _u32, _u64, _u32, _u32, _u32, _u32, _i64, f64p]),
    "pimc_energy_now": (C.c_int, [_vp, f64p, f64p, f64p]),
    "pimc_action": (C.c_int, [_vp, f64p, f64p]),
    "pimc_find_nn": (C.c_int, [_vp, _i32, f64p, _i64, _i64, i64p]),
    "pimc_find_nns": (C.c_int, [_vp, _i32, f64p, _i64, _i64, i64p, _i64, i64p]),
    "pimc_update_nnbins": (C.c_int, [_vp]),
    "pimc_reshape_linear_explicit": (C.c_int, [_vp, _i32, _i64, _i64, _i64, f64p, _d, _i32, f64p, f64p, f64p, i32p]),
    "pimc_reshape_swap_explicit": (C.c_int, [_vp, _i32, _i64, _i64, _i64, _i64, f64p, f64p, _d, _i32, f64p, f64p, i32p]),
    "pimc_com_explicit": (C.c_int, [_vp, _i32, _i64, _i32, f64p, _d, _i32, f64p, f64p, i32p]),
    "pimc_swap_weights": (C.c_int, [_vp, _i32, _i64, _i64, _i64, f64p]),
    "pimc_update_create": (C.c_int, [_vp, _i32, _d, i32p]),
    "pimc_update_configure": (C.c_int, [_vp, _i32, _d, _d, _d, _d, _i64, _i64]),
    "pimc_update_get": (C.c_int, [_vp, _i32, _i32, f64p, i64p, i64p, f64p, i64p, i64p]),
    "pimc_energy_create": (C.c_int, [_vp, _i64, i32p]),
    "pimc_energy_read": (C.c_int, [_vp, _i32, _i32, f64p, f64p, _i64, i64p]),
    "pimc_energy_read_range": (C.c_int, [_vp, _i32, _i32, _i64, _i64, f64p, f64p, i64p]),
    "pimc_energy_stats": (C.c_int, [_vp, _i32, f64p]),
    "pimc_density_create": (C.c_int, [_vp, _i64, i32p]),
    "pimc_density_measure": (C.c_int, [_vp, _i32]),
    "pimc_density_read": (C.c_int, [_vp, _i32, f64p, i64p, f64p]),
    "pimc_paircorr_create": (C.c_int, [_vp, _i64, _d, i32p]),
    "pimc_paircorr_measure": (C.c_int, [_vp, _i32]),
    "pimc_paircorr_read": (C.c_int, [_vp, _i32, f64p, i64p, f64p]),
    "pimc_winding_create": (C.c_int, [_vp, _i64, i32p]),
    "pimc_winding_now": (C.c_int, [_vp, f64p]),
    "pimc_winding_read": (C.c_int, [_vp, _i32, _i32, f64p, _i64, i64p]),
    "pimc_structure_create": (C.c_int, [_vp, _i32, i32p]),
    "pimc_structure_measure": (C.c_int, [_vp, _i32]),
    "pimc_structure_read": (C.c_int, [_vp, _i32, f64p, i64p, i32p]),
    "pimc_compressibility": (C.c_int, [_vp, _i32, f64p, f64p]),
    "pimc_run_ex": (C.c_int, [_vp, _i64, i32p, i64p, _i32, C.POINTER(Measurements), _i32, C.POINTER(RunStats)]),
    "pimc_comm_get_unique_id": (C.c_int, [_vp]),
    "pimc_comm_init": (C.c_int, [_vp, _i32, _i32, _vp]),
    "pimc_comm_init_all": (C.c_int, [C.POINTER(_vp), _i32]),
    "pimc_comm_info": (C.c_int, [_vp, i32p, i32p, i64p, i32p]),
    "pimc_build_prop_table": (C.c_int, [_d, _d, _d, _i32, f64p, f64p, f64p]),
    "pimc_prop_int": (C.c_int, [f64p, _i32, _d, _d, f64p, f64p, _i32, _d, f64p]),
    "pimc_determine_nnrange": (C.c_int, [f64p, _i32, _d, _d, _d, _d, _d, f64p]),
    "pimc_run": (C.c_int, [_vp, _i64, i32p, i64p, _i32, i32p, _i32, i32p, _i32, _i32, C.POINTER(RunStats)]),
}

_lib = None


def load():
    """Load libpimc_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). pimc_jl_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = load().pimc_last_error(handle)
        raise PimcError(rc, msg.decode() if msg else "")
    return rc


def make_potential(kind="zero", dv="zero", k=1.0, depth=0.0, scale=1.0, sgn=1.0, angles=(), helical=False):
    p = Potential()
    p.kind = {"zero": 0, "harmonic": 1, "sin2_1d": 2, "lattice": 3}[kind]
    p.dv_kind = {"zero": 0, "identity": 1, "gradient": 2}[dv]
    p.k, p.depth, p.scale, p.sgn = float(k), float(depth), float(scale), float(sgn)
    if len(angles) > MAX_ANGLES:
        raise ValueError(f"at most {MAX_ANGLES} beam angles")
    p.nang, p.helical = len(angles), int(helical)
    for i, a in enumerate(angles):
        p.ang[i] = float(a)
    return p
