"""examples/tools/savetools.jl counterparts (pure numpy part): table layout of save_paths / save_density."""
import numpy as np
from pimc_jl_b200 import tools


def test_paths_table_layout():
    N, dim, M, beta = 3, 2, 4, 2.0
    r = np.arange(N * dim * M, dtype=float).reshape(N, dim, M)
    nxt = np.array([2, 1, 3])                      # particles 1 and 2 exchange, 3 is closed on itself
    names, data = tools.paths_table(r, nxt, beta)
    assert names == ["tau", "p1 x", "p1 y", "p2 x", "p2 y", "p3 x", "p3 y"]          # savetools.jl:6-10
    assert data.shape == (M + 1, 1 + dim * N)
    assert np.allclose(data[:, 0], [j * beta / M for j in range(M + 1)])             # savetools.jl:12
    assert np.array_equal(data[:M, 1], r[0, 0]) and data[M, 1] == r[1, 0, 0]         # ring of particle 1 closes on particle 2
    assert data[M, 3] == r[0, 0, 0] and data[M, 5] == r[2, 0, 0] and data[M, 6] == r[2, 1, 0]
    names1, data1 = tools.paths_table(r[:, :1], nxt, beta)
    assert names1 == ["tau", "p1 x", "p2 x", "p3 x"] and data1.shape == (M + 1, 4)


def test_density_table_normalisation(tmp_path):
    nb, L = 5, 2.0
    dens = np.arange(nb * nb, dtype=float).reshape(nb, nb)
    names, data = tools.density_table(dens, nb, L, 2 * L / nb, 10)
    assert names[0] == "pos" and data.shape == (nb, nb + 1)
    assert np.allclose(data[:, 0], np.linspace(-L, L, nb)) and np.allclose(data[:, 1:], dens / (0.8 * 10))   # savetools.jl:37-45
    n1, d1 = tools.density_table(dens[0], nb, L, 0.8, 10)
    assert n1 == ["pos", "Density"] and d1.shape == (nb, 2)
    p = tmp_path / "d.csv"
    tools._write_csv(p, names, data)
    back = np.loadtxt(p, delimiter=",", skiprows=1)
    assert np.array_equal(back, data)
