"""probe: g(r) and winding estimators at C3 shape (N = 256, M = 100), timed per launch; also the target of the ncu captures of k_paircorr / k_winding"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
e = pj.Engine(pj.make_potential("harmonic", "identity"), dim=2, M=100, N=256, chains=C, L_=16.0, T=0.5, lam=0.5, Ncycle=5, seed=1)
ups = [(1, e.update_create(L.UPD_POLYMER_COM, 1.0)), (1, e.update_create(L.UPD_RESHAPE_LINEAR, 20)), (1, e.update_create(L.UPD_RESHAPE_SWAP, 20))]
pc, wi = e.paircorr_create(400, 8.0), e.winding_create(4096)
e.run(20, ups, sched=L.SCHED_SWEEP)
a = e.run(50, ups, sched=L.SCHED_SWEEP)
b = e.run(50, ups, sched=L.SCHED_SWEEP, paircorrs=[pc], windings=[wi])
h, nd, _ = e.paircorr_read(pc, 400)
print(f"C={C}: moves only {a['kernel_ms']:.2f} ms, with g(r)+winding every 5 iterations {b['kernel_ms']:.2f} ms -> {(b['kernel_ms'] - a['kernel_ms']) / 10:.3f} ms per measurement event "
      f"({C * 100 * 256 * 255 // 2 / ((b['kernel_ms'] - a['kernel_ms']) / 10 * 1e-3):.3e} pair distances/s); pairs counted {int(h.sum())}, ndata {nd}, "
      f"<W^2> {e.winding_read(wi, -1)[0].mean():.4f}")
# static structure factor (k_structure), kmax = 4: timed per functor call
import torch
sk = e.structure_create(4)
e.structure_measure(sk)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    e.structure_measure(sk)
t1 = time.perf_counter()
S, nds = e.structure_read(sk, 4)
print(f"S(k), kmax 4: {(t1 - t0) / 5 * 1e3:.3f} ms per event ({C * 100 * 256 / ((t1 - t0) / 5):.3e} beads/s); S(pi/L (1,0)) = {S[1, 4] / (nds * 256):.4f}, kappa_T = {e.compressibility(sk)[0]:.4f}")
