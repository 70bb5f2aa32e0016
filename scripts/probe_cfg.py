"""ncu probe of one bench.py workload: thermalise, then a short tail of launches in a fixed order (the ones ncu captures).
usage: probe_cfg.py <workload> <impl> <therm> [isweep]
tail (after <therm> thermalisation iterations of the full update list, moves only):
  per update family alone: 2 iterations (k_sweep with one family: the staging / centre-of-mass halves on their own)
  full list with the workload's estimator: 2 * Ncycle iterations (k_sweep mix + k_measure)
Prints the number of library launches before the tail so the caller can pass it to `ncu -s`."""
import sys
sys.path.insert(0, '.')
import bench
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L

name, impl, therm = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = bench.WORKLOADS[name]
e = pj.Engine(pj.make_potential(**wl["pot"]), dim=wl["dim"], M=wl["M"], N=wl["N"], chains=wl["chains"], L_=wl["L"], T=wl["T"], lam=wl["lam"],
              Ncycle=wl["Ncycle"], seed=1, **bench.interaction_args(wl))
if impl:
    e.set_option(L.OPT_SWEEP_IMPL, impl)
if len(sys.argv) > 4:
    e.set_option(L.OPT_ISWEEP, int(sys.argv[4]))
kind = {"com": L.UPD_SINGLE_COM, "reshape": L.UPD_RESHAPE_LINEAR, "swap": L.UPD_RESHAPE_SWAP, "pcom": L.UPD_POLYMER_COM}
ups = [(every, e.update_create(kind[k], v0)) for k, every, v0 in wl["updates"]]
lib = L.load()
l0 = lib.pimc_launch_count()
e.run(therm, ups, sched=L.SCHED_SWEEP)
print("launches before the tail:", lib.pimc_launch_count() - l0, flush=True)
if not wl.get("interactions"):
    for (k, every, v0), (_, uid) in zip(wl["updates"], ups):
        if k != "swap":
            st = e.run(2, [(1, uid)], sched=L.SCHED_SWEEP)
            print(k, "alone:", st["bead_moves"], "bead-moves in 2 iterations,", st["launches"], "launches", flush=True)
obj = e.density_create(wl["nbins"]) if wl["measure"] == "density" else e.energy_create(64)
kw = dict(densities=[obj]) if wl["measure"] == "density" else dict(energies=[obj])
st = e.run(2 * wl["Ncycle"], ups, sched=L.SCHED_SWEEP, **kw)
print("mix + estimator:", st["bead_moves"], "bead-moves,", st["launches"], "launches,", st["measurements"], "measurement events")
