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


def checkpoint(s, path, updates=()):
    """All chains of this rank: worldlines, permutation, iteration counter, measurement counters, adaptive variables."""
    e = s.engine
    r, _, _, nxt = e.paths(want=("r", "next"))
    sc = e.scalars()
    var = np.array([[e.update_get(u.id, c)["var"] for c in range(e.C)] for _, u in updates]) if updates else np.zeros((0, e.C))
    np.savez_compressed(path, r=r, next=nxt, iter=sc["iter"], N_MC=sc["N_MC"], Nctr=sc["Nctr"], var=var,
                        shape=np.array([e.C, e.N, e.dim, e.M]))
    return path


def restore(s, path):
    """Puts the worldlines / permutation back and re-arms the Philox iteration counter; the link cache and the cell lists are rebuilt by
    the library (pimc_set_paths).  Returns the saved adaptive variables (one row per update) for the caller to re-create its updates with."""
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    e = s.engine
    if tuple(z["shape"]) != (e.C, e.N, e.dim, e.M):
        raise ValueError(f"checkpoint shape {tuple(z['shape'])} does not match this System {(e.C, e.N, e.dim, e.M)}")
    e.set_paths(z["r"], z["next"])
    e.set_iter(int(z["iter"]))
    return z["var"]
