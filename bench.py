#!/usr/bin/env python
"""bench.py -- bead-moves/sec of the PIMC hot path on N B200s (BASELINE.json metric), roofline and CPU baseline.

A "step" is one `run!(s, iters, updates; Zmeasurements=[Energy])` over all chains resident on the GPU
(sweep schedule), i.e. one pass of the hot path (staging-bridge moves + centre-of-mass moves + Delta-U +
Metropolis + Energy estimator) over one batch of synthetic worldlines.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                     # CPU arm: the oracle port of the reference on host cores

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): examples/energy_2d_free_bose_gas.jl scaled to
V=0, lambda=1, T=1, N=64, M=128, L=16, Ncycle=2, 4096 chains per GPU (weak scaling), updates
SingleCenterOfMass(1.0):1 + ReshapeLinear(20):1, Energy measured every Ncycle iterations.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

B_ALG = 48.0   # algorithmic bytes per bead-move: read+write of position (2x8 B) and cached link action (8 B)
# algorithmic flops per bead-move: 47 (V = 0), 51 (harmonic), ~149 (12-beam lattice) -- per workload below
L25 = [2.214297435588181, 0.9272952180016122, -0.6435011087932844, 0.6435011087932844, -2.498091544796509, 3.141592653589793,
       2.498091544796509, 0, 1.5707963267948966, -2.2142974355881813, -1.5707963267948968, -0.9272952180016123]
_LAT = dict(kind="lattice", dv="zero", depth=6.0, scale=1.0, sgn=-1.0, angles=L25)   # examples/density_SRL_lattice.jl:16, potentialtools.jl:28-29
# SURVEY.md 8d configurations.  pot: keyword arguments of make_potential; measure: "energy" | "density"; sched: "sweep" | "faithful"
WORKLOADS = {
    "c1": dict(name="C1 2D trap N=1 M=5 (as shipped)", pot=dict(kind="harmonic", dv="identity"), dim=2, N=1, M=5, L=100.0, T=1.0, lam=0.5, Ncycle=10,
               chains=4096, updates=[("com", 1, 1.0), ("reshape", 1, 2)], measure="energy", sched="sweep", F_alg=51.0),
    "c2": dict(name="C2 2D free Bose gas N=64 M=128", pot=dict(kind="zero", dv="identity"), dim=2, N=64, M=128, L=16.0, T=1.0, lam=1.0, Ncycle=2,
               chains=4096, updates=[("com", 1, 1.0), ("reshape", 1, 20)], measure="energy", sched="sweep", F_alg=47.0),
    "c2s": dict(name="C2 + ReshapeSwapLinear(20):1", pot=dict(kind="zero", dv="identity"), dim=2, N=64, M=128, L=16.0, T=1.0, lam=1.0, Ncycle=2,
                chains=4096, updates=[("com", 1, 1.0), ("reshape", 1, 20), ("swap", 1, 20)], measure="energy", sched="sweep", F_alg=47.0),
    "c2f": dict(name="C2, reference schedule (one particle per iteration)", pot=dict(kind="zero", dv="identity"), dim=2, N=64, M=128, L=16.0, T=1.0,
                lam=1.0, Ncycle=2, chains=4096, updates=[("com", 1, 1.0), ("reshape", 1, 20)], measure="energy", sched="faithful", F_alg=47.0),
    "c3": dict(name="C3 2D trap density N=256 M=100 L=16, non-interacting as shipped", pot=dict(kind="harmonic", dv="identity"), dim=2, N=256, M=100,
               L=16.0, T=0.5, lam=0.5, Ncycle=5, chains=1024, updates=[("pcom", 1, 1.0), ("reshape", 1, 20), ("swap", 1, 20)], measure="density",
               nbins=500, sched="sweep", F_alg=51.0),
    "c3i": dict(name="C3 with hard core a=0.05, r_a=1.0, lnU table, cell list (32x32 cells per slice)", pot=dict(kind="harmonic", dv="identity"),
                dim=2, N=256, M=100, L=16.0, T=0.5, lam=0.5, Ncycle=5, chains=1024, updates=[("pcom", 1, 1.0), ("reshape", 1, 20), ("swap", 1, 20)],
                measure="density", nbins=500, sched="faithful", interactions=True, a=0.05, r_a=1.0, F_alg=51.0, iters=100, therm=50),
    "c4": dict(name="C4 lattice density N=128 M=256 L=8 V0=6 l25, non-interacting", pot=_LAT, dim=2, N=128, M=256, L=8.0, T=0.2,
               lam=1.0 / 9.869604401089358, Ncycle=3, chains=1024, updates=[("com", 1, 1.0), ("reshape", 1, 5), ("swap", 20, 20)], measure="density",
               nbins=500, sched="sweep", F_alg=149.0),
    "c4i": dict(name="C4 with interactions, g=2 forwarded to System (hard core a=exp(-pi), lnU table, r_a from the propagator)", pot=_LAT, dim=2, N=128, M=256, L=8.0,
                T=0.2, lam=1.0 / 9.869604401089358, Ncycle=3, chains=1024, updates=[("com", 1, 1.0), ("reshape", 1, 5), ("swap", 20, 20)],
                measure="density", nbins=500, sched="faithful", interactions=True, g=2.0, r_a=0.0, F_alg=149.0),
    # examples/density_SRL_lattice.jl:17-19 AS SHIPPED: g enters build_prop_int only and is not forwarded to System, so a = exp(-2 pi / 0.0) = 0
    # (src/system.jl:151): pair action through the lnU table (swap move), no hard core
    "c4i0": dict(name="C4 with interactions exactly as the script ships (g=2 in the pair-propagator table only: a = 0, lnU table, r_a from the propagator)",
                 pot=_LAT, dim=2, N=128, M=256, L=8.0, T=0.2, lam=1.0 / 9.869604401089358, Ncycle=3, chains=1024,
                 updates=[("com", 1, 1.0), ("reshape", 1, 5), ("swap", 20, 20)], measure="density", nbins=500, sched="faithful", interactions=True,
                 g=2.0, g_system=0.0, r_a=0.0, F_alg=149.0),
    "c5": dict(name="C5 2D trapped gas N=1024 M=64", pot=dict(kind="harmonic", dv="identity"), dim=2, N=1024, M=64, L=100.0, T=1.0, lam=0.5, Ncycle=10,
               chains=512, updates=[("com", 1, 1.0), ("reshape", 1, 2)], measure="energy", sched="sweep", F_alg=51.0),
}
UPD_NAMES = {"com": "SingleCenterOfMass", "pcom": "PolymerCenterOfMass", "reshape": "ReshapeLinear", "swap": "ReshapeSwapLinear"}


def interaction_args(wl):
    """g, r_a and the pair-propagator term table of an interacting workload (host-side build, pimc_jl_b200/propint.py)"""
    if not wl.get("interactions"):
        return {}
    import math
    from pimc_jl_b200 import propint
    g = wl.get("g") or -2 * math.pi / math.log(wl["a"])          # a = exp(-2 pi / g), src/system.jl:151
    tau = (1.0 / wl["T"]) / wl["M"]
    p = propint.build_prop_int(math.ceil(math.sqrt(2) * wl["L"]), g, tau)   # examples/density_SRL_lattice.jl:17
    r_a = wl["r_a"] or propint.determine_nnrange(p, tau, 1e-20, wl["L"])
    return dict(interactions=True, g=wl.get("g_system", g), r_a=r_a, tab=p["tab"], tab_lo=p["lo"], tab_hi=p["hi"])   # g_system: what System(...) itself receives


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.rows)}


def oracle_arm(wl, threads, iters, seed=1, therm=60):
    """The CPU restatement (oracle port of the reference) on `threads` host threads, one chain per thread
    (mirrors the example scripts' pmap over independent runs). Returns (bead_moves, seconds)."""
    import oracle_binding as ob
    ob.build()
    kind = {"com": ob.UPD_SINGLE_COM, "reshape": ob.UPD_RESHAPE_LINEAR, "swap": ob.UPD_RESHAPE_SWAP, "pcom": ob.UPD_POLYMER_COM}
    systems = []
    ia = interaction_args(wl)
    sched = ob.SCHED_SWEEP if wl["sched"] == "sweep" else ob.SCHED_FAITHFUL
    for t in range(threads):
        s = ob.System(ob.make_potential(**wl["pot"]), dim=wl["dim"], M=wl["M"], N=wl["N"], L=wl["L"], T=wl["T"], lam=wl["lam"],
                      Ncycle=wl["Ncycle"], seed=seed, chain=t, **ia)
        ups = [(every, ob.Update(s, kind[k], v0)) for k, every, v0 in wl["updates"]]
        en = ob.Energy(max(16, (therm + iters) // wl["Ncycle"] + 16)) if wl["measure"] == "energy" else ob.Density(s, wl["nbins"])
        systems.append((s, ups, en))

    def work(i, n, measure):
        s, ups, en = systems[i]
        kw = {} if not measure else (dict(energies=[en]) if wl["measure"] == "energy" else dict(densities=[en]))
        s.run(n, ups, sched=sched, **kw)

    def par(n, measure):
        th = [threading.Thread(target=work, args=(i, n, measure)) for i in range(threads)]
        [t.start() for t in th]
        [t.join() for t in th]
    par(therm, False)  # let the adaptive slice count / step settle like the GPU arm
    bm0 = sum(u.get()["bead_moves"] for s, ups, en in systems for _, u in ups)
    t0 = time.perf_counter()
    par(iters, True)
    dt = time.perf_counter() - t0
    bm1 = sum(u.get()["bead_moves"] for s, ups, en in systems for _, u in ups)
    return bm1 - bm0, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--iters", type=int, default=0, help="run! iterations per step (default 400; 40000 for the reference schedule)")
    ap.add_argument("--therm", type=int, default=-1, help="untimed thermalisation iterations before the warm-up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sched", default="", choices=["", "faithful", "sweep"], help="override the workload's schedule (side measurements)")
    ap.add_argument("--faithful-impl", type=int, default=0, help="reference-schedule proposals: 0 warp-cooperative (default), 1 one thread (A/B)")
    ap.add_argument("--cpu-iters", type=int, default=0)
    ap.add_argument("--isweep", type=int, default=-1, help="interacting sweep schedule: 0 sequential kernel (library default), 1 adaptive, 2 optimistic kernels (side measurements)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.sched:
        wl["sched"] = args.sched
    if args.chains:
        wl["chains"] = args.chains
    faithful = wl["sched"] == "faithful"
    if not args.iters:
        args.iters = wl.get("iters") or ((1000 if wl.get("interactions") else 40000) if faithful else 400)
    if args.therm < 0:
        args.therm = wl.get("therm") or ((500 if wl.get("interactions") else 20000) if faithful else 300)
    F_ALG = wl["F_alg"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = os.cpu_count() or 1
    config = {"workload": wl["name"], "chains_per_gpu": wl["chains"], "N": wl["N"], "M": wl["M"], "dim": wl["dim"],
              "updates": " + ".join(f"{UPD_NAMES[k]}({v0}):{every}" for k, every, v0 in wl["updates"]),
              "schedule": "sweep (every worldline of a chain proposes per iteration)" if not faithful else "reference (one proposal per chain and iteration)",
              "iters_per_step": args.iters, "measure": f"{wl['measure'].capitalize()} every {wl['Ncycle']} iterations",
              "interactions": bool(wl.get("interactions")),
              "e2e_pipeline": "double-buffered: two Systems of the full batch on two streams / host threads take the steps alternately, the copies of one overlap the moves of the other",
              "l2": f"state ({wl['chains']} chains x {wl['N'] * wl['M'] * 24 // 1024} KiB) larger than L2"}

    if args.impl == "reference":
        # the reference's own CPU implementation of the path: Julia is absent from this image, so the oracle port is timed
        if rank != 0:
            return
        per_step = args.cpu_iters or (max(8, 2000 // max(1, (args.steps + args.warmup))) if not faithful else (wl.get("iters") or (1000 if wl.get("interactions") else 20000)))
        threads = ncpu
        tot_bm, tot_t = 0, 0.0
        for step in range(args.warmup + args.steps):
            bm, dt = oracle_arm(wl, threads, per_step, seed=1 + step)
            if step >= args.warmup:
                tot_bm += bm
                tot_t += dt
        val = tot_bm / tot_t
        sample = f"{threads} chains (one per host thread) x {per_step} run! iterations per step, oracle port (C, -O2), not Julia"
        line = {"metric": "bead-moves/sec", "value": val, "unit": "bead-moves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": val, "unit": "bead-moves/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "bead-moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import pimc_jl_b200 as pj
    trace_on = os.environ.get("PIMC_BENCH_TRACE") is not None

    def trace(msg):
        if trace_on:
            print(f"[bench rank {rank} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)
    if trace_on:
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("PIMC_BENCH_TRACE") or 90), exit=False, file=sys.stderr)
    from pimc_jl_b200 import _lib as L, engine as eng
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = L.load()
    Cc = wl["chains"]
    e = pj.Engine(pj.make_potential(**wl["pot"]), dim=wl["dim"], M=wl["M"], N=wl["N"], chains=Cc, chain_offset=rank * Cc, L_=wl["L"],
                  T=wl["T"], lam=wl["lam"], Ncycle=wl["Ncycle"], seed=1, device=local_rank, **interaction_args(wl))
    SCHED = L.SCHED_FAITHFUL if faithful else L.SCHED_SWEEP

    def attach_comm(engine_):
        """multi-GPU inside the library: ncclCommInitRank on this engine's handle; the 128-byte id travels through torch.distributed (plumbing)"""
        ids = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        engine_.comm_init(world, rank, ids[0])
    if dist is not None:
        attach_comm(e)
    trace('engine + communicator ready')
    if args.faithful_impl:
        e.set_option(L.OPT_FAITHFUL_IMPL, args.faithful_impl)
    if args.isweep >= 0:
        e.set_option(L.OPT_ISWEEP, args.isweep)
    use_density = wl["measure"] == "density"
    kind = {"com": L.UPD_SINGLE_COM, "reshape": L.UPD_RESHAPE_LINEAR, "swap": L.UPD_RESHAPE_SWAP, "pcom": L.UPD_POLYMER_COM}
    ups = [(every, e.update_create(kind[k], v0)) for k, every, v0 in wl["updates"]]
    nmeas_total = (args.therm + (args.warmup + args.steps) * args.iters * 2) // wl["Ncycle"] + 64
    en = e.density_create(wl["nbins"]) if use_density else e.energy_create(nmeas_total)
    mkw = dict(densities=[en]) if use_density else dict(energies=[en])
    stream = torch.cuda.current_stream()
    e.set_stream(stream.cuda_stream)
    e.run(args.therm, ups, sched=SCHED)  # thermalisation: adaptive slice count / step settle
    trace('thermalised')

    def block_read(n_before, engine_=None, obj=None):
        """the step's estimator block.  With N > 1 GPUs the library has all-reduced it itself (ncclAllReduce on its side stream, queued at the
        end of pimc_run; density counters at read-out): chain-mean E, Ev over ALL ranks' chains / density counters and ndata summed over the ranks"""
        engine_, obj = engine_ or e, en if obj is None else obj
        if use_density:
            dens, nd, _ = engine_.density_read(obj, wl["nbins"])
            return torch.from_numpy(dens), nd
        E, Ev, n = engine_.energy_read_range(obj, n_before, 1 << 20)
        return torch.from_numpy(np.stack([E, Ev])), n
    block_allreduce = block_read

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: `value` ----
    n_seen = 0 if use_density else e.energy_read(en, -1, cap=0)[2]
    for _ in range(args.warmup):
        e.run(args.iters, ups, sched=SCHED, **mkw)
        _, n_seen = block_allreduce(n_seen)
    trace('warm-up done')
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    l0 = lib.pimc_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bead_moves, kern_ms = 0, 0.0
    ev0.record(stream)
    for _ in range(args.steps):
        st = e.run(args.iters, ups, sched=SCHED, **mkw)
        bead_moves += st["bead_moves"]
        kern_ms += st["kernel_ms"]
        blk, n_seen = block_allreduce(n_seen)
    ev1.record(stream)
    barrier()
    launches = lib.pimc_launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if sampler:
        sampler.stop_flag = True
        sampler.join()
    t = torch.tensor([ms, kern_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(bead_moves)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms, kern_ms, total_bm = float(t[0]), float(t[1]), float(tot[0])
    value = total_bm / (ms * 1e-3)

    trace('device-resident arm done')
    # ---- end-to-end arm: host buffers in, host buffers out, copies inside the timed region ----
    # The public API as a user with pinned host buffers drives it, double-buffered: TWO Systems of the full batch (two handles, two streams, two
    # host threads -- the library is synchronous per handle) take the steps alternately, so the H2D / D2H copies of one batch overlap the moves
    # of the other.  Every step still uploads its batch, runs it, reads the estimator block and downloads the batch.  Runs and collective
    # read-outs take turns in step order, which keeps the NCCL operations of the two communicators in the same order on every rank.
    per = wl["N"] * wl["dim"] * wl["M"]
    halves = []
    for k in range(2 if args.steps > 1 else 1):
        ek = pj.Engine(pj.make_potential(**wl["pot"]), dim=wl["dim"], M=wl["M"], N=wl["N"], chains=Cc, chain_offset=rank * Cc, L_=wl["L"],
                       T=wl["T"], lam=wl["lam"], Ncycle=wl["Ncycle"], seed=1 + 1000 * (k + 1), device=local_rank, **interaction_args(wl))
        if args.faithful_impl:
            ek.set_option(L.OPT_FAITHFUL_IMPL, args.faithful_impl)
        if args.isweep >= 0:
            ek.set_option(L.OPT_ISWEEP, args.isweep)
        sk = torch.cuda.Stream()
        ek.set_stream(sk.cuda_stream)
        if dist is not None:
            attach_comm(ek)
        uk = [(every, ek.update_create(kind[kk], v0)) for kk, every, v0 in wl["updates"]]
        ok = ek.density_create(wl["nbins"]) if use_density else ek.energy_create(nmeas_total)
        ek.run(args.therm, uk, sched=SCHED)
        hk = torch.empty((Cc, wl["N"], wl["dim"], wl["M"]), dtype=torch.float64).pin_memory().numpy()
        ek.get_r_into(hk)
        halves.append(dict(e=ek, ups=uk, obj=ok, host=hk, stream=sk, mkw=dict(densities=[ok]) if use_density else dict(energies=[ok]), seen=0, bm=0, blk=None))
    trace('e2e buffers ready')
    turn = [0]
    cv = threading.Condition()
    errors = []

    def half_worker(k):
        try:
            torch.cuda.set_device(local_rank)
            hf = halves[k]
            for s_ in range(k, args.steps, len(halves)):               # this buffer's steps
                hf["e"].set_paths(hf["host"])                          # H2D: this step's worldlines (pinned host memory)
                with cv:
                    cv.wait_for(lambda: turn[0] >= s_)
                st_ = hf["e"].run(args.iters, hf["ups"], sched=SCHED, **hf["mkw"])
                hf["bm"] += st_["bead_moves"]
                hf["blk"], hf["seen"] = block_read(hf["seen"], hf["e"], hf["obj"])   # the step's result: estimator block, D2H (global over the GPUs)
                with cv:
                    turn[0] += 1
                    cv.notify_all()
                hf["e"].get_r_into(hf["host"])                         # D2H: updated worldlines
        except Exception as ex:   # noqa: BLE001
            errors.append(ex)
            with cv:
                turn[0] = 1 << 30
                cv.notify_all()
    barrier()
    t0 = time.perf_counter()
    workers = [threading.Thread(target=half_worker, args=(k,)) for k in range(len(halves))]
    [w.start() for w in workers]
    [w.join() for w in workers]
    barrier()
    t_e2e = time.perf_counter() - t0
    trace('e2e arm done')
    if errors:
        raise errors[0]
    e2e_bm = sum(hf["bm"] for hf in halves)
    blk_host = halves[0]["blk"]
    n_seen_e2e = halves[0]["seen"]
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    be = torch.tensor([float(e2e_bm)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(be)
    e2e_val = float(be[0]) / float(te[0])
    h2d = per * Cc * 8
    d2h = per * Cc * 8 + (8 * wl["nbins"] ** wl["dim"] if use_density else 2 * 8 * (args.iters // wl["Ncycle"]))
    comm_info = e.comm_info()

    hbm, how = peaks()
    # dominant kernel: k_sweep (one launch per iteration: staging-bridge + centre-of-mass sweeps of every chain).  Its own launch time
    # is measured live on a moves-only leg (no estimator launches in between): achieved = algorithmic bytes per launch / avg launch time.
    l1 = lib.pimc_launch_count()
    st_mv = e.run(args.iters, ups, sched=SCHED)
    n_launch = max(1, lib.pimc_launch_count() - l1)
    sweep_ms = st_mv["kernel_ms"] / n_launch
    sweep_bytes = st_mv["bead_moves"] * B_ALG / n_launch
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    # per-family legs (after every timed region; they perturb nothing that is reported above): one update family alone, moves only, and the
    # estimator launch as the difference between a measured and an unmeasured run of the same moves
    by_family = {}
    if not faithful and not wl.get("interactions"):   # (one move family alone on a hard-core system spins in the reference's retry loops: not a throughput figure)
        fam_iters = max(8, args.iters // 4)
        for (kk, every, v0), (_, uid) in zip(wl["updates"], ups):
            lf = lib.pimc_launch_count()
            stf = e.run(fam_iters, [(1, uid)], sched=SCHED)
            nl = max(1, lib.pimc_launch_count() - lf)
            if stf["bead_moves"] > 0 and stf["kernel_ms"] > 0:
                gbs = stf["bead_moves"] * B_ALG / (stf["kernel_ms"] * 1e-3) / 1e9
                by_family[UPD_NAMES[kk]] = {"bead_moves_per_s": stf["bead_moves"] / (stf["kernel_ms"] * 1e-3), "achieved": gbs, "frac": gbs / hbm,
                                            "launch_ms": stf["kernel_ms"] / nl, "alg_bytes_per_launch": stf["bead_moves"] * B_ALG / nl}
        n_it = wl["Ncycle"] * max(4, fam_iters // wl["Ncycle"])
        st_a = e.run(n_it, ups, sched=SCHED)
        st_b = e.run(n_it, ups, sched=SCHED, **mkw)
        nm = max(1, st_b["measurements"])
        ms_meas = max(0.0, st_b["kernel_ms"] - st_a["kernel_ms"] * (st_b["bead_moves"] / max(1, st_a["bead_moves"]))) / nm
        bytes_meas = Cc * wl["N"] * wl["M"] * 8.0 * wl["dim"]          # the estimators stream the positions once (16 B per bead at d = 2)
        if ms_meas > 0:
            by_family["measure"] = {"launch_ms_marginal": ms_meas, "alg_bytes_per_launch": bytes_meas, "achieved": bytes_meas / (ms_meas * 1e-3) / 1e9,
                                    "frac": bytes_meas / (ms_meas * 1e-3) / 1e9 / hbm,
                                    "note": "marginal cost of one measurement event inside the step (fused Energy sums ride on the centre-of-mass sweep)"}
    trace('family legs done')
    # (every rank runs the legs above: a measured run queues the library's all-reduce of its Energy block, which all ranks must join)
    # tear the communicators down in the same order on every rank (ncclCommDestroy waits for its peers), before rank 0 goes on alone
    barrier()
    for hf in halves:
        hf["e"].close()
    e.close()
    if dist is not None:
        dist.destroy_process_group()
        dist = None
    trace('communicators closed')
    if rank != 0:
        return
    fp64 = C.c_double(0.0)
    lib.pimc_measure_fp64_peak(C.byref(fp64))
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch of k_sweep from the committed ncu --set full capture
    if os.path.exists(tfile):
        with open(tfile) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload and tj.get("chains") == Cc:
            traffic = tj.get("k_sweep_dram_bytes_per_launch")
    whole = total_bm / world / (kern_ms * 1e-3)            # moves + estimator kernels, per GPU
    big = Cc * wl["N"] * wl["M"] >= 2 ** 20
    kname = "k_sweep" if st_mv["launches"] > 1 else ("k_run_cells" if wl.get("interactions") else ("k_chain" if (big and not faithful) else "k_run"))
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                "peak_source": f"MEASURED_PEAKS.json ({how})", "alg_bytes_per_bead_move": B_ALG,
                "alg_bytes_per_launch": sweep_bytes, "launch_ms": sweep_ms, "by_family": by_family,
                "traffic_source": "committed ncu --set full capture (profiles/traffic.json), not measured in this run" if traffic is not None else None,
                "step_including_estimator": {"achieved": whole * B_ALG / 1e9, "frac": whole * B_ALG / 1e9 / hbm},
                "fp64": {"achieved_tflops": achieved * 1e9 / B_ALG * F_ALG / 1e12, "peak_tflops_measured_dfma": fp64.value,
                         "frac": (achieved * 1e9 / B_ALG * F_ALG / 1e12 / fp64.value) if fp64.value else None, "alg_flops_per_bead_move": F_ALG}}
    if use_density:   # sum(dens)/ndata vs N (test/testmeasurements.jl:24-29), chain-summed
        check = {"density_sum_over_ndata": float(blk_host.sum()) / n_seen_e2e if n_seen_e2e else None, "expected": float(wl["N"])}   # both global
    else:
        Em = float(blk_host[0].mean()) if blk_host.numel() else None
        check = {"E_mean_last_block": Em, "E_expected_boltzmannon": wl["dim"] * wl["N"] / 2.0 * wl["T"] if wl["pot"]["kind"] == "zero" else None}
    line = {"metric": "bead-moves/sec", "value": value, "unit": "bead-moves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": sampler.summary() if sampler else None,
            "e2e": {"value": e2e_val, "unit": "bead-moves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline,
            "collective": {"where": "libpimc_b200 (ncclAllReduce on the handle's side stream, queued at the end of pimc_run; density counters at read-out)",
                           "comm_nranks_seen": comm_info["nranks"], "chains_total": comm_info["chains_total"], "nccl_version": comm_info["nccl_version"]}, "kernel_ms_per_step": kern_ms / args.steps,
            "check": check}
    if not args.no_cpu_baseline:
        it = args.cpu_iters or (150 if not faithful else (wl.get("iters") or (1000 if wl.get("interactions") else 20000)))
        bm, dt = oracle_arm(wl, ncpu, it)
        line["cpu_baseline"] = {"value": bm / dt, "unit": "bead-moves/s", "cores": ncpu, "kind": "port",
                                "sample": f"{ncpu} chains (one per host thread) x {it} run! iterations after 60 thermalisation iterations; "
                                          f"oracle port of the reference (C, -O2), not Julia"}
        # the reference's real execution model (src/ has no threading): one chain on one host thread
        bm1, dt1 = oracle_arm(wl, 1, it)
        line["cpu_baseline"]["single_thread"] = {"value": bm1 / dt1, "unit": "bead-moves/s", "cores": 1, "sample": f"1 chain x {it} run! iterations, same port"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
