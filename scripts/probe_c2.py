import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
M = int(sys.argv[3]) if len(sys.argv) > 3 else 128
t0 = time.time()
e = pj.Engine(pj.make_potential("zero", "identity"), dim=2, M=M, N=N, chains=C, L_=16.0, T=1.0, lam=1.0, Ncycle=2, seed=1)
print("create", time.time() - t0)
com, rl = e.update_create(L.UPD_SINGLE_COM, 1.0), e.update_create(L.UPD_RESHAPE_LINEAR, 20)
en = e.energy_create(20000)
for ups, name in (([(1, com)], "com"), ([(1, rl)], "reshape"), ([(1, com), (1, rl)], "mix")):
    st = e.run(150, ups, sched=L.SCHED_SWEEP)
    st = e.run(40, ups, energies=[en], sched=L.SCHED_SWEEP)
    print(name, "ms", st["kernel_ms"], "bead-moves/s %.3e" % (st["bead_moves"] / st["kernel_ms"] * 1e3), "hbm frac %.3f" % (st["bead_moves"] * 48 / st["kernel_ms"] * 1e3 / 6557.4e9),
          e.update_get(com), e.update_get(rl))
E, Ev, n = e.energy_read(en)
print("E mean", E.mean(), "expect", N * 2 / 2.0, "n", n)
