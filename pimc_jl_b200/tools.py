"""Host-side counterparts of the reference's example tools (examples/tools/savetools.jl, logtools.jl) and state
checkpoint / restore.  Pure numpy on arrays fetched through the C ABI; nothing here is on the hot path.

    paths_table / save_paths      savetools.jl:4-34    "tau, p1 x, p1 y, ..." with M + 1 rows (the ring closed on the next particle)
    density_table / save_density  savetools.jl:36-71   dens / (bin * ndata) with a leading `pos` column
    info_system / info_updates    logtools.jl:6-44
    checkpoint / restore          paths, permutation, iteration counter (= Philox counter) and adaptive variables of one System
"""
import numpy as np


def paths_table(r, nxt, beta):
    """r: (N, dim, M) worldlines of one chain, nxt: (N,) 1-based `next`.  Returns (column names, (M + 1, 1 + dim * N) array)."""
    N, dim, M = r.shape
    names = ["tau"]
    cols = [np.array([j * beta / M for j in range(M + 1)])]
    for n in range(N):
        closing = int(nxt[n]) - 1               # pcycle(M + 1, pol, Npol, M): the particle that owns slice M + 1
        for d in range(dim):
            names.append(f"p{n + 1} {'xy'[d]}" if dim == 2 else f"p{n + 1} x")
            cols.append(np.concatenate([r[n, d, :], r[closing, d, :1]]))
    return names, np.stack(cols, axis=1)


def density_table(dens, nbins, L, binw, ndata):
    """dens / (bin * ndata) with the `pos` column of density_df (range(-L, L, length = nbins))."""
    den = np.asarray(dens, dtype=np.float64) / (binw * ndata)
    pos = np.linspace(-L, L, nbins)
    if den.ndim == 1:
        return ["pos", "Density"], np.stack([pos, den], axis=1)
    return ["pos"] + [f"x{i + 1}" for i in range(den.shape[1])], np.concatenate([pos[:, None], den], axis=1)


def _write_csv(path, names, data):
    with open(path, "w") as f:
        f.write(",".join(names) + "\n")
        for row in data:
            f.write(",".join(repr(float(v)) for v in row) + "\n")


def save_paths(s, path="paths.csv", chain=0):
    r, _, _, nxt = s.engine.paths(chain, 1, want=("r", "next"))
    names, data = paths_table(r[0], nxt[0], s.beta)
    _write_csv(path, names, data)
    return path


def save_density(s, d, g, name, V, path=None):
    dens, ndata = d._read()
    names, data = density_table(dens, d.nbins, s.L, d.bin, ndata)
    path = path or f"density{name}M{s.M}beta{s.beta}N{s.N}g{g}V{V}.csv"
    _write_csv(path, names, data)
    return path


def info_system(s):
    return "\n".join([f"PIMC: {s.N} particles in a {s.dim}D potential",
                      f"Number of time slices M = {s.M}", f"Chemical potential    mu = {s.mu}", f"scattering length     a = {s.a}",
                      f"Inverse temperature   beta = {s.beta}", f"Time slice size       tau = {s.tau}",
                      "Canonical ensemble with No Worm Algorithm"])


def info_updates(updates):
    out = []
    for every, u in updates:
        name = type(u).__name__
        g = u._get()
        if "Center" in name:
            out.append(f"{name}({every}):\n\tAcceptance {g['acc_window']:.3f}\n\tStep\t   {g['var']:.3f}")
        else:
            out.append(f"{name}({every}):\n\tAcceptance {g['acc_window']:.3f}\n\tSlices\t   {int(round(g['var']))}")
    return "\n".join(out)


def checkpoint(s, path):
    """Complete state of all chains of this rank (pimc_get_state): worldlines, permutation, cached link actions, cell lists, RNG iteration
    counter, measurement cadence, every update object's adaptive variable / counters / acceptance window, estimator accumulators.
    The reference's savetools (examples/tools/savetools.jl:4-34) write the paths only."""
    e = s.engine
    path = str(path) if str(path).endswith(".npz") else str(path) + ".npz"
    np.savez_compressed(path, state=e.get_state(), shape=np.array([e.C, e.N, e.dim, e.M]))
    return path


def restore(s, path):
    """Puts a checkpoint back into a System built with the same arguments and the same update / Energy / Density objects created in the same
    order; the run then continues bit for bit (tests/test_gpu_api.py::test_checkpoint_restore_continues_bit_for_bit)."""
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    e = s.engine
    if tuple(z["shape"]) != (e.C, e.N, e.dim, e.M):
        raise ValueError(f"checkpoint shape {tuple(z['shape'])} does not match this System {(e.C, e.N, e.dim, e.M)}")
    e.set_state(z["state"])
