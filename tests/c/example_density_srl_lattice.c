/* The call sequence of examples/density_SRL_lattice.jl (main(), lines 16-44, and the pmap body, lines 48-54) through the C ABI of
 * include/pimc_b200.h -- the calls julia/Pimc/src/Pimc.jl issues when that script runs against the shim -- scaled down in n / times
 * (the script: n = 100_000, times = 10) so that the test finishes in seconds.  Julia is not available in this image; this driver is how
 * the boundary is exercised without it.  Built and run by tests/test_gpu_api.py::test_c_driver_example_density_srl_lattice, which holds the
 * density histogram it writes against the oracle's.
 *
 *   usage: example_density_srl_lattice <g> <V0> <n> <times> <out.bin> [schedule]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pimc_b200.h"

#define CHECK(h, call) do { int rc_ = (call); if (rc_ != PIMC_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, pimc_last_error(h)); return 1; } } while (0)

/* examples/tools/potentialtools.jl:28-29, :l25 */
static const double L25[12] = { 2.214297435588181, 0.9272952180016122, -0.6435011087932844, 0.6435011087932844, -2.498091544796509, 3.141592653589793,
                                2.498091544796509, 0, 1.5707963267948966, -2.2142974355881813, -1.5707963267948968, -0.9272952180016123 };

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s g V0 n times out.bin [sched]\n", argv[0]); return 2; }
    const double g = atof(argv[1]), V0 = atof(argv[2]);
    const long n = atol(argv[3]), times = atol(argv[4]);
    const int sched = argc > 6 ? atoi(argv[6]) : PIMC_SCHED_FAITHFUL;
    /* pmap body: scale = 1.0; name = :l25; potential = generate_V(scale, V0, name; attractive = true); L = 8.0; M = 200; N = 20; T = 0.2 */
    const double scale = 1.0, L = 8.0, T = 0.2;
    const int M = 200, N = 20;

    /* main():17  propint = build_prop_int(round(sqrt(L^2+L^2), RoundUp), g, 1/(T*M)) */
    const int delta = 600;
    double *tab = malloc(sizeof(double) * delta * delta), lo, hi;
    CHECK(NULL, pimc_build_prop_table(ceil(sqrt(L * L + L * L)), g, 1.0 / (T * M), delta, tab, &lo, &hi));

    /* main():18-19  s = System(potential; lambda = 1/pi^2, M, N, L, T, propint, interactions = true, length_measurement_cycle = 3)
     *   -- g and r_a are NOT forwarded: g = 0.0 => a = exp(-2 pi / 0.0) = 0 (system.jl:151); r_a = 0.0 => determine_nnrange (system.jl:29-31) */
    pimc_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.dim = 2; cfg.M = M; cfg.N = N; cfg.chains = 1; cfg.chain_offset = 0; cfg.mu = 0.0; cfg.lambda = 1.0 / (M_PI * M_PI); cfg.L = L; cfg.T = T;
    cfg.interactions = 1; cfg.g = 0.0; cfg.r_a = 0.0; cfg.Ncycle = 3; cfg.compat = PIMC_COMPAT_ALL; cfg.init = 1; cfg.seed = 0x5EEDB200ull;
    cfg.pot.kind = PIMC_POT_LATTICE; cfg.pot.dv_kind = PIMC_DV_ZERO; cfg.pot.k = 1.0; cfg.pot.depth = V0; cfg.pot.scale = scale; cfg.pot.sgn = -1.0;
    cfg.pot.nang = 12; memcpy(cfg.pot.ang, L25, sizeof L25);
    cfg.tab = tab; cfg.tab_n = delta; cfg.tab_lo = lo; cfg.tab_hi = hi; cfg.device = -1;
    pimc_handle *s = NULL;
    CHECK(NULL, pimc_create(&cfg, &s));

    /* main():21-25  updates = [(1, SingleCenterOfMass(s, 1.0)), (1, ReshapeLinear(s, 5)), (1, ReshapeSwapLinear(s, 20))] */
    int32_t upd[3]; const int64_t every[3] = { 1, 1, 1 };
    CHECK(s, pimc_update_create(s, PIMC_UPD_SINGLE_COM, 1.0, &upd[0]));
    CHECK(s, pimc_update_configure(s, upd[0], 1e-1, L / 2, 0.4, 0.6, 10, 10000));
    CHECK(s, pimc_update_create(s, PIMC_UPD_RESHAPE_LINEAR, 5, &upd[1]));
    CHECK(s, pimc_update_configure(s, upd[1], 2, M - 2, 0.6, 0.8, 10, 10000));
    CHECK(s, pimc_update_create(s, PIMC_UPD_RESHAPE_SWAP, 20, &upd[2]));
    CHECK(s, pimc_update_configure(s, upd[2], 2, M - 2, 0.6, 0.8, 10, 10000));
    /* main():26  mea = ZMeasurement[Density(s)]  (nbins = 500) */
    int32_t dens; CHECK(s, pimc_density_create(s, 500, &dens));

    /* main():28  info(s) reads these */
    double sc[5]; int64_t isc[5];
    CHECK(s, pimc_get_scalars(s, sc, isc));
    printf("beta %.17g tau %.17g vol %.17g a %.17g r_a %.17g nbins %lld\n", sc[0], sc[1], sc[2], sc[3], sc[4], (long long)isc[0]);

    /* main():32  run!(s, n*times, updates)  -- thermalisation, no measurements */
    pimc_run_stats st;
    CHECK(s, pimc_run(s, n * times, upd, every, 3, NULL, 0, NULL, 0, sched, &st));
    /* main():35-38  while s.N_MC[s.N] < n*times; run!(s, n, updates, Zmeasurements = mea); @info s.N_MC[s.N]; end */
    for (;;) {
        CHECK(s, pimc_get_scalars(s, sc, isc));
        if (isc[1] >= n * times) break;
        CHECK(s, pimc_run(s, n, upd, every, 3, NULL, 0, &dens, 1, sched, &st));
    }
    /* pmap body:53  save_density(s, d, g, name, V0) reads d.dens, d.bin and d.ndata (examples/tools/savetools.jl:36-71) */
    double *d = malloc(sizeof(double) * 500 * 500), bin; int64_t ndata;
    CHECK(s, pimc_density_read(s, dens, d, &ndata, &bin));
    double sum = 0; for (int i = 0; i < 500 * 500; ++i) sum += d[i];
    printf("N_MC %lld ndata %lld bin %.17g sum %.17g\n", (long long)isc[1], (long long)ndata, bin, sum);
    FILE *f = fopen(argv[5], "wb"); if (!f) return 3;
    fwrite(d, sizeof(double), 500 * 500, f); fclose(f);
    pimc_destroy(s); free(tab); free(d);
    return 0;
}
