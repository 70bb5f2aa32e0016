import sys
sys.path.insert(0, '.')
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = pj.Engine(pj.make_potential("zero", "identity"), dim=2, M=128, N=64, chains=C, L_=16.0, T=1.0, lam=1.0, Ncycle=2, seed=1)
com, rl = e.update_create(L.UPD_SINGLE_COM, 1.0), e.update_create(L.UPD_RESHAPE_LINEAR, 20)
en = e.energy_create(2000)
e.run(int(sys.argv[2]) if len(sys.argv) > 2 else 260, [(1, com), (1, rl)], energies=[en], sched=L.SCHED_SWEEP)
print(e.update_get(rl))
