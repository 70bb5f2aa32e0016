import sys
sys.path.insert(0, '.')
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = pj.Engine(pj.make_potential("zero", "identity"), dim=2, M=128, N=64, chains=C, L_=16.0, T=1.0, lam=1.0, Ncycle=2, seed=1)
com, rl = e.update_create(L.UPD_SINGLE_COM, 1.0), e.update_create(L.UPD_RESHAPE_LINEAR, 20)
e.run(150, [(1, rl)], sched=L.SCHED_SWEEP)
for ups, name in (([(1, com)], "com"), ([(1, rl)], "reshape"), ([(1, com), (1, rl)], "mix")):
    best = 0
    for rep in range(3):
        st = e.run(40, ups, sched=L.SCHED_SWEEP)
        best = max(best, st["bead_moves"] / st["kernel_ms"] * 1e3)
    print(name, "bead-moves/s %.3e" % best, "hbm frac %.3f" % (best * 48 / 6557.4e9))
