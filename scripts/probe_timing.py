"""Per-phase cycle counts of the staging sweep (library built with -DEXP_TIMING, selected through PIMC_B200_SO)."""
import sys
sys.path.insert(0, '.')
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L
C, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pot, impl = sys.argv[4], int(sys.argv[5])
e = pj.Engine(pj.make_potential(pot, "identity"), dim=2, M=M, N=N, chains=C, L_=16.0 if pot == "zero" else 6.0, T=1.0, lam=1.0 if pot == "zero" else 0.5, Ncycle=2, seed=1)
e.set_option(L.OPT_SWEEP_IMPL, impl)
com, rl = e.update_create(L.UPD_SINGLE_COM, 1.0), e.update_create(L.UPD_RESHAPE_LINEAR, 20)
e.run(256, [(1, com), (1, rl)], sched=L.SCHED_SWEEP)
st = e.run(192, [(1, rl)], sched=L.SCHED_SWEEP)
print("reshape-only", impl, st["bead_moves"] / st["kernel_ms"] * 1e3, flush=True)
