"""probe: interacting sweep (optimistic kernels vs the sequential persistent kernel) at growing sizes, with timings per stage"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L

wlname, chains, iters, isw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
upds = sys.argv[5].split(",") if len(sys.argv) > 5 else None
wl = dict(bench.WORKLOADS[wlname])
t0 = time.time()
e = pj.Engine(pj.make_potential(**wl["pot"]), dim=wl["dim"], M=wl["M"], N=wl["N"], chains=chains, L_=wl["L"], T=wl["T"], lam=wl["lam"],
              Ncycle=wl["Ncycle"], seed=1, device=0, **bench.interaction_args(wl))
print(f"create {time.time() - t0:.2f}s  a={e.a} nbins={e.nbins}", flush=True)
e.set_option(L.OPT_ISWEEP, isw)
kind = {"com": L.UPD_SINGLE_COM, "reshape": L.UPD_RESHAPE_LINEAR, "swap": L.UPD_RESHAPE_SWAP, "pcom": L.UPD_POLYMER_COM}
ups = [(every, e.update_create(kind[k], v0)) for k, every, v0 in wl["updates"] if upds is None or k in upds]
for rep in range(4):
    t0 = time.time()
    st = e.run(iters, ups, sched=L.SCHED_SWEEP)
    print(f"run {rep}: {time.time() - t0:.3f}s kernel {st['kernel_ms']:.2f} ms, {st['bead_moves']} bead moves -> {st['bead_moves'] / st['kernel_ms'] * 1e3:.3e}/s, launches {st['launches']}", flush=True)
