"""A/B of the sweep implementations on the C2 shape (or N M given): moves-only throughput of COM-only, reshape-only and the mix.
usage: probe_ab.py [chains N M pot impl...]"""
import sys
sys.path.insert(0, '.')
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
M = int(sys.argv[3]) if len(sys.argv) > 3 else 128
pot = sys.argv[4] if len(sys.argv) > 4 else "zero"
impls = [int(x) for x in sys.argv[5:]] or [2, 3]
Lbox = 16.0 if pot == "zero" else 6.0
for impl in impls:
    e = pj.Engine(pj.make_potential(pot, "identity"), dim=2, M=M, N=N, chains=C, L_=Lbox, T=1.0, lam=1.0 if pot == "zero" else 0.5, Ncycle=2, seed=1)
    e.set_option(L.OPT_SWEEP_IMPL, impl)
    com, rl = e.update_create(L.UPD_SINGLE_COM, 1.0), e.update_create(L.UPD_RESHAPE_LINEAR, 20)
    e.run(300, [(1, com), (1, rl)], sched=L.SCHED_SWEEP)     # adaptive variables settle
    for ups, name in (([(1, com)], "com"), ([(1, rl)], "reshape"), ([(1, com), (1, rl)], "mix")):
        e.run(20, ups, sched=L.SCHED_SWEEP)
        st = e.run(100, ups, sched=L.SCHED_SWEEP)
        bm = st["bead_moves"] / st["kernel_ms"] * 1e3
        print(f"impl {impl} {pot} N={N} M={M} {name:8s} {st['kernel_ms'] / 100 * 1e3:8.1f} us/launch  {bm:.3e} bead-moves/s  hbm frac {bm * 48 / 6557.4e9:.3f}  "
              f"m={e.update_get(rl)['var']:.0f} step={e.update_get(com)['var']:.2f}", flush=True)
    e.close()
