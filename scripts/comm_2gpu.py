"""2-rank check of the library's communicator (run under torchrun on 2 GPUs: gpurun --gpus 2):
chains sharded over the ranks, estimator blocks all-reduced by libpimc_b200 itself; the global read-outs on every rank must equal a
single-GPU run holding all the chains (Energy means to 1e-13 -- the sums associate differently --, density counters integer-equal)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pimc_jl_b200 as pj
from pimc_jl_b200 import _lib as L, engine

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
Ctot = 10                                           # uneven shards: 5 + 5 at world 2, 4 + 3 + 3 at world 3
base, rem = divmod(Ctot, world)
cnt = base + (1 if rank < rem else 0)
off = rank * base + min(rank, rem)
kw = dict(dim=2, M=32, N=8, T=1.0, lam=0.5, Ncycle=2, seed=11, L_=4.0)


def run(e):
    ups = [(1, e.update_create(L.UPD_SINGLE_COM, 1.0)), (1, e.update_create(L.UPD_RESHAPE_LINEAR, 6)), (2, e.update_create(L.UPD_RESHAPE_SWAP, 6))]
    en, de = e.energy_create(128), e.density_create(24)
    blocks = []
    for b in range(4):
        e.run(20, ups, energies=[en], densities=[de], sched=L.SCHED_SWEEP)
        blocks.append(e.energy_read_range(en, 10 * b, 10)[0])
    E, Ev, n = e.energy_read(en, -1)
    d, nd, _ = e.density_read(de, 24)
    return np.concatenate(blocks), E, Ev, n, d, nd


e = pj.Engine(pj.make_potential("harmonic", "identity"), chains=cnt, chain_offset=off, device=lr, **kw)
ids = [engine.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
e.comm_init(world, rank, ids[0])
info = e.comm_info()
assert info["nranks"] == world and info["rank"] == rank and info["chains_total"] == Ctot, info
got = run(e)
ref = run(pj.Engine(pj.make_potential("harmonic", "identity"), chains=Ctot, chain_offset=0, device=lr, **kw))   # all chains on this GPU, no communicator
assert got[3] == ref[3] == 40
for a, b in ((got[0], ref[0]), (got[1], ref[1]), (got[2], ref[2])):
    assert np.allclose(a, b, rtol=1e-13, atol=0), np.abs(a - b).max()
assert np.array_equal(got[4], ref[4]) and got[5] == ref[5]
print(f"rank {rank}/{world}: library communicator ok (nccl {info['nccl_version']}), global Energy blocks and density equal the single-GPU run", flush=True)
dist.barrier()

# ---- the same through the host-side mirror of the reference API (pimc.System under torchrun attaches the library communicator itself) ----
import pimc_jl_b200.pimc as P
kwm = dict(dV="identity", dim=2, M=32, N=8, L=4.0, T=1.0, lam=0.5, length_measurement_cycle=2, seed=11, schedule="sweep", device=lr)


def run_mirror(s):
    ups = [(1, P.SingleCenterOfMass(s, 1.0)), (1, P.ReshapeLinear(s, 6))]
    en, de = P.Energy(s, 128), P.Density(s, nbins=24)
    P.run_b(s, 40, ups, Zmeasurements=[en, de])
    return en.energy[s.N], de.dens, de.ndata


s = P.System(P.harmonic(), chains=Ctot, **kwm)               # sharded over the ranks
assert s.library_comm and s.engine.C == cnt
Em, dm, ndm = run_mirror(s)
dist.barrier()
ref_s = pj.Engine(pj.make_potential("harmonic", "identity"), chains=Ctot, chain_offset=0, device=lr, dim=2, M=32, N=8, T=1.0, lam=0.5, Ncycle=2, seed=11, L_=4.0)
ur = [(1, ref_s.update_create(L.UPD_SINGLE_COM, 1.0)), (1, ref_s.update_create(L.UPD_RESHAPE_LINEAR, 6))]
er, dr = ref_s.energy_create(128), ref_s.density_create(24)
ref_s.run(40, ur, energies=[er], densities=[dr], sched=L.SCHED_SWEEP)
assert np.allclose(Em, ref_s.energy_read(er, -1)[0], rtol=1e-13, atol=0) and np.array_equal(dm, ref_s.density_read(dr, 24)[0]) and ndm == ref_s.density_read(dr, 24)[1]
print(f"rank {rank}/{world}: pimc.System mirror with the library communicator ok", flush=True)
dist.barrier()
s.engine.close()
e.close()
dist.destroy_process_group()
