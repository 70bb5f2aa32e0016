/*
 * pimc_b200.h -- C ABI of libpimc_b200.so: the B200-native (sm_100a) engine behind the
 * Julia-facing API of oameye/PIMC.jl.  Plain pointers and sizes; no torch / CUDA types.
 *
 * The reference has no FFI boundary (pure Julia); these entry points are
 * what a `ccall` shim of the reference's `System` / `run!` / update functors / measurement
 * functors binds (INTEGRATION.md shows the Julia side).  Each entry cites the reference
 * interface it replaces as file:line relative to the reference tree.
 *
 * Conventions: every call returns int (0 ok, <0 error class; text via pimc_last_error);
 * no C++ exception crosses the boundary; one host thread per handle; calls are synchronous
 * at return; pointer arguments are caller-owned HOST buffers valid only during the call;
 * particle / slice / bin indices crossing the boundary are 1-based like the reference's.
 * A handle holds `chains` independent replicas of one reference `System`
 * (global chain id = chain_offset + local index, so results do not depend on the sharding).
 */
#ifndef PIMC_B200_H
#define PIMC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIMC_OK               0
#define PIMC_ERR_INVALID     -1
#define PIMC_ERR_CUDA        -2
#define PIMC_ERR_UNSUPPORTED -3
#define PIMC_ERR_NOMEM       -4
#define PIMC_ERR_STATE       -5

/* Potential descriptor: the reference passes Julia closures `V`, `dV` (src/system.jl:114-115,129-131);
 * a kernel cannot call them, so the shim lowers them to one of these families
 * (examples/ *.jl, test/testsystem.jl:13, examples/tools/potentialtools.jl:1-39). */
enum { PIMC_POT_ZERO = 0, PIMC_POT_HARMONIC = 1, PIMC_POT_SIN2_1D = 2, PIMC_POT_LATTICE = 3 };
enum { PIMC_DV_ZERO = 0, PIMC_DV_IDENTITY = 1, PIMC_DV_GRADIENT = 2 };
#define PIMC_MAX_ANGLES 32
typedef struct {
    int32_t kind, dv_kind;
    double k;                 /* harmonic: (0.5*k)*(x^2+y^2)                                      */
    double depth, scale, sgn; /* sin2_1d: depth*sin(2pi*x*scale)^2 ; lattice: sgn*depth*I(r)      */
    int32_t nang, helical;
    double ang[PIMC_MAX_ANGLES];
} pimc_potential;

/* compat flags: set = reproduce the reference as shipped */
#define PIMC_COMPAT_PAIR_BYVALUE  1
#define PIMC_COMPAT_SWAP_SIGN     2
#define PIMC_COMPAT_DENSITY_SHIFT 4
#define PIMC_COMPAT_SWAP_STALE_LINK 8 /* B14: an accepted swap leaves the cached link at slice j_m un-exchanged (reshape.jl:269-275) */
#define PIMC_COMPAT_ALL           15

/* keyword arguments of System(...) (src/system.jl:129-145) + batching / sharding / seed */
typedef struct {
    int32_t dim, M, N;
    int32_t chains;           /* independent Markov chains held by this handle (this GPU)          */
    uint32_t chain_offset;    /* global id of local chain 0                                        */
    double mu, lambda, L, T;
    int32_t interactions;
    double g, r_a;
    int32_t Ncycle;           /* length_measurement_cycle                                          */
    int32_t compat;
    int32_t init;             /* 1: init_world (src/system.jl:36-78) on the device ; 0: zeros      */
    uint64_t seed;
    pimc_potential pot;
    const double *tab;        /* prop_rel_interpolate_terms table (src/propagator.jl:35-70), n x n col-major, or NULL */
    int32_t tab_n;
    double tab_lo, tab_hi;
    int32_t device;           /* CUDA device ordinal, -1 = current                                 */
} pimc_config;

typedef struct pimc_handle pimc_handle;

enum { PIMC_UPD_RESHAPE_LINEAR = 0, PIMC_UPD_RESHAPE_SWAP = 1, PIMC_UPD_SINGLE_COM = 2, PIMC_UPD_POLYMER_COM = 3 };
/* FAITHFUL: one update per chain per iteration, exactly run! (src/simulation.jl:29-42).
 * SWEEP: per iteration the picked update is proposed once for EVERY worldline of the chain inside one
 * imaginary-time window (independent when a == 0 and the moves carry no pair action); DESIGN.md. */
enum { PIMC_SCHED_FAITHFUL = 0, PIMC_SCHED_SWEEP = 1 };

typedef struct {
    int64_t iterations;       /* run! iterations executed per chain                      */
    int64_t proposals;        /* update proposals, all chains                            */
    int64_t accepted;         /* accepted proposals, all chains                          */
    int64_t bead_moves;       /* proposed new bead positions, all chains  */
    int64_t measurements;     /* measurement events per chain (N_MC increment)           */
    int64_t launches;         /* kernels of this library launched by the call            */
    double kernel_ms;         /* device time of the run kernel(s), CUDA events on the run stream */
} pimc_run_stats;

/* ---- lifetime ---- */
int  pimc_create(const pimc_config *cfg, pimc_handle **out);           /* System(...)  src/system.jl:129-167 */
void pimc_destroy(pimc_handle *h);                                      /* Julia finalizer                    */
const char *pimc_last_error(const pimc_handle *h);                      /* NULL handle: last create error     */
int  pimc_version(void);
int  pimc_set_stream(pimc_handle *h, void *cuda_stream);                /* run kernels on this cudaStream_t   */
/* engine options (no reference counterpart). PIMC_OPT_SWEEP_IMPL (sweep schedule, independent worldlines): 0 auto, 1 persistent kernel of the
 * reference-schedule bodies (k_run), 2 per-iteration sweep kernels (k_sweep + k_swap_iter + k_measure), 3 chain-major persistent sweep kernel
 * (k_chain: every CTA takes a chain through all iterations of the call).  Auto: below 2^20 beads 1; above, 3 when the chains fill at most two
 * rounds of CTA slots (measured faster there) and 2 for larger batches (measured faster there).  Identical trajectories. */
#define PIMC_OPT_SWEEP_IMPL 1
/* PIMC_OPT_FAITHFUL_IMPL: proposals of the reference schedule: 0 warp-cooperative (default), 1 one thread per proposal (A/B, same bits) */
#define PIMC_OPT_FAITHFUL_IMPL 2
/* PIMC_OPT_FUSE_ENERGY: 1 evaluates the Energy functor inside the sweep launch for chains whose centre-of-mass sweep streams every worldline
 * anyway; 0 (default) uses the estimator launch.  Same values to 1e-12.  Measured on C2 (profiles/r02_summary.md): the fused sums push the
 * HBM-bound centre-of-mass sweep over its issue budget and cost more than the TMA-fed estimator pass they save. */
#define PIMC_OPT_FUSE_ENERGY 3
/* PIMC_OPT_ISWEEP: sweep schedule of interacting worldlines (hard core; pair action not counted in ReshapeLinear / centre-of-mass moves =
 * the reference as shipped): 0 (default) the sequential sweep inside the persistent kernel; 2 the optimistic-parallel per-iteration kernels
 * (pimc_isweep.cuh); 1 adaptive between the two (falls back to the sequential kernel for a few runs whenever more than 30 % of the proposals
 * had to be replayed serially).  All three produce identical trajectories.  Measured on B200 (profiles/r02_summary.md) the optimistic path
 * does not yet beat the sequential kernel on C3i / C4i, hence the default. */
#define PIMC_OPT_ISWEEP 4
int  pimc_set_option(pimc_handle *h, int32_t option, int64_t value);
int64_t pimc_launch_count(void);                                        /* kernels launched by this library so far (bench evidence) */
/* measurement utility (no reference counterpart): sustained non-tensor fp64 FMA rate of the current device, in TFLOP/s */
int  pimc_measure_fp64_peak(double *tflops);

/* ---- state: s.world[n].{r,V,bins,next} (src/system.jl:1-6), host layout = reference layout:
 *      r[chain][n][dim][M] (per particle the M x dim column-major matrix), V[chain][n][M], bins, next[chain][n] ---- */
int pimc_get_paths(pimc_handle *h, int32_t chain0, int32_t nchains, double *r, double *V, int64_t *bins, int64_t *next);
int pimc_set_paths(pimc_handle *h, int32_t chain0, int32_t nchains, const double *r, const int64_t *next);
int pimc_get_scalars(pimc_handle *h, double *out5 /* beta,tau,vol,a,r_a */, int64_t *iout5 /* nbins,N_MC,Nctr,ctr,iter */);
int pimc_set_iter(pimc_handle *h, uint64_t iter);
/* checkpoint / resume (the reference's examples/tools/savetools.jl:4-34 saves the paths only): the COMPLETE state of every chain
 * as one opaque host blob -- positions, permutation, cached link actions, cell lists (order and multiplicities), the iteration counter of the
 * addressed RNG, N_MC / Nctr, every update object's variable / counters / acceptance window, every estimator's accumulators.
 * pimc_set_state needs a handle of the same System with the same update / Energy / Density objects created in the same order;
 * the run then continues bit for bit. */
int pimc_state_size(pimc_handle *h, int64_t *bytes);
int pimc_get_state(pimc_handle *h, void *buf, int64_t cap);
int pimc_set_state(pimc_handle *h, const void *buf, int64_t bytes);

/* ---- propagator primitives, evaluated on the device (parity hooks; src/propagator.jl) ---- */
int pimc_distance(int64_t n, const double *x1, const double *x2, double L, double *out);          /* :6-9   */
int pimc_teleport(int64_t n, const double *x, double L, double *out);                              /* :30-32 */
int pimc_lnK(int64_t n, const double *r1, const double *r2, int32_t dim, double tau, double lambda, double L, double *out); /* :16-19 */
int pimc_lnV(int64_t n, const double *r1, const double *r2, int32_t dim, double tau, const pimc_potential *p, double *out); /* :26-28 */
int pimc_potential_eval(int64_t n, const double *r, int32_t dim, const pimc_potential *p, double *V, double *dV);
/* levy! (src/updates/helper.jl:118-139): nb bridges, each rows x dim column-major, Gaussians xi (rows-2) x dim row-major */
int pimc_levy_bridge(double *r, int32_t rows, int32_t dim, double tau, double L, double lambda, const double *xi, int64_t nb);
/* the RNG spec of include/pimc_rng.h evaluated on the device: n consecutive beads of one address */
int pimc_gauss_pairs(uint64_t seed, uint32_t chain, uint64_t iter, uint32_t slot, uint32_t kind, uint32_t retry,
                     uint32_t bead0, int64_t n, double *g /* 2n */);

/* ---- estimators / action of the CURRENT configuration, every chain ---- */
/* Energy functor (src/measurement.jl:92-122): E[chains], Ev[chains], parts[chains][3] = link, pot, vkin */
int pimc_energy_now(pimc_handle *h, double *E, double *Ev, double *parts);
/* sum of the link cache and of recomputed lnV links (src/system.jl:72-74, src/updates/reshape.jl:68-77) */
int pimc_action(pimc_handle *h, double *links_cached, double *links_recomputed);

/* ---- neighbour search: GPU cell list replacing src/nearest_neighbours.jl ---- */
int pimc_find_nn(pimc_handle *h, int32_t chain, const double *r, int64_t slice, int64_t exception, int64_t *nn);   /* :156-179 */
int pimc_find_nns(pimc_handle *h, int32_t chain, const double *r, int64_t slice, int64_t exception, int64_t *out, int64_t cap, int64_t *count); /* :131-154 */
int pimc_update_nnbins(pimc_handle *h);                                                                            /* :182-196 */

/* ---- single moves with caller-supplied randomness (deterministic Delta-U parity hooks) ---- */
/* ReshapeLinear functor body, src/updates/reshape.jl:56-87 */
int pimc_reshape_linear_explicit(pimc_handle *h, int32_t chain, int64_t n, int64_t j0, int64_t m, const double *xi, double u,
                                 int32_t commit, double *w_initial, double *w_updated, double *rprime, int32_t *acc);
/* ReshapeSwapLinear functor body, src/updates/reshape.jl:138-279 */
int pimc_reshape_swap_explicit(pimc_handle *h, int32_t chain, int64_t n1, int64_t n2, int64_t j0, int64_t m, const double *xi1,
                               const double *xi2, double u, int32_t commit, double *w_initial, double *w_updated, int32_t *acc);
/* Single/PolymerCenterOfMass functor body, src/updates/com.jl:47-100,168-220 ; d = displacement */
int pimc_com_explicit(pimc_handle *h, int32_t chain, int64_t n, int32_t polymer, const double *d, double u, int32_t commit,
                      double *w_initial, double *w_updated, int32_t *acc);
/* sampleparticles weight table exp.(t + y), src/updates/helper.jl:230-260 */
int pimc_swap_weights(pimc_handle *h, int32_t chain, int64_t n1, int64_t j0, int64_t m, double *w);

/* ---- update objects (structs of src/updates/com.jl:7-28,112-133, src/updates/reshape.jl:7-29,99-121) ---- */
int pimc_update_create(pimc_handle *h, int32_t kind, double var0, int32_t *id);
int pimc_update_configure(pimc_handle *h, int32_t id, double vmin, double vmax, double minacc, double maxacc, int64_t adj, int64_t range);
/* chain >= 0: that chain ; chain = -1: totals over chains (var, acc_window = chain means) */
int pimc_update_get(pimc_handle *h, int32_t id, int32_t chain, double *var, int64_t *tries, int64_t *tries_var,
                    double *acc_window, int64_t *accepted, int64_t *bead_moves);

/* ---- measurement objects (src/measurement.jl:31-38,78-86) ---- */
int pimc_energy_create(pimc_handle *h, int64_t cap, int32_t *id);
/* chain >= 0: that chain's series ; chain = -1: mean over this handle's chains per measurement index */
int pimc_energy_read(pimc_handle *h, int32_t id, int32_t chain, double *E, double *Ev, int64_t cap, int64_t *n);
/* same, entries [start, start+count) only (a measurement block); *n = total number of measurements taken */
int pimc_energy_read_range(pimc_handle *h, int32_t id, int32_t chain, int64_t start, int64_t count, double *E, double *Ev, int64_t *n);
/* per-chain accumulators: out[chains][5] = n, sum E, sum E^2, sum Ev, sum Ev^2 */
int pimc_energy_stats(pimc_handle *h, int32_t id, double *out);
int pimc_density_create(pimc_handle *h, int64_t nbins, int32_t *id);
int pimc_density_measure(pimc_handle *h, int32_t id);                   /* Density functor now, src/measurement.jl:45-55 */
/* dens: nbins^dim counts (column-major like the Julia array), summed over this handle's chains; ndata likewise */
int pimc_density_read(pimc_handle *h, int32_t id, double *dens, int64_t *ndata, double *bin);

/* ---- estimators the reference lists as TODO (src/measurement.jl:125-127 `#TODO radial distribution`, `#TODO Superfluid Fraction`), written in the
 *      style of its functors (the CPU checker under tests/ restates the same definitions).
 *  g(r): per measurement, for every slice and every pair i < j, ib = floor(|distance.(r_i, r_j, L)| / (rmax / nbins)); hist[ib] += 1 if
 *        ib < nbins; ndata += M.   g(r_b) = hist[b] * vol / (ndata * N (N - 1) / 2 * shell_b) at read-out.
 *  winding: W_k = (1 / 2L) * sum over links of teleport(r_next[k] - r[k], L), an integer for closed paths; superfluid fraction
 *        rho_s / rho = <W^2> (2L)^2 / (2 dim lambda beta N). ---- */
int pimc_paircorr_create(pimc_handle *h, int64_t nbins, double rmax, int32_t *id);
int pimc_paircorr_measure(pimc_handle *h, int32_t id);                                       /* functor call now, every chain */
int pimc_paircorr_read(pimc_handle *h, int32_t id, double *hist, int64_t *ndata, double *bin); /* summed over chains (and ranks) */
int pimc_winding_create(pimc_handle *h, int64_t cap, int32_t *id);
int pimc_winding_now(pimc_handle *h, double *W /* [chains][dim] */);
/* chain >= 0: out[n][dim] that chain's series; chain = -1: out[n] = mean over this handle's chains of W^2 per measurement */
int pimc_winding_read(pimc_handle *h, int32_t id, int32_t chain, double *out, int64_t cap, int64_t *n);
/* `#TODO Compressibilty` (src/measurement.jl:127), canonical ensemble: static structure factor on the wave vectors of the periodic box,
 *  k = (pi / L)(a, b), a = 0..kmax, |b| <= kmax (half plane a > 0, or a = 0 and b > 0; 1-D: b = 0, a >= 1), kmax <= 6.  Per measurement and slice m:
 *  rho_k(m) = sum_n exp(i k . r_n[m, :]); sums[a * (2 kmax + 1) + b + kmax] += |rho_k(m)|^2; ndata += M.  S(k) = sums / (ndata * N).
 *  Compressibility from the long-wavelength limit S(k -> 0) = rho k_B T kappa_T on the smallest shell |k| = pi / L:
 *  kappa_T = beta * S(k_min) / rho, rho = N / (2L)^dim (a finite-size estimate; a number-fluctuation estimator would need the grand-canonical
 *  worm sector the reference never shipped). */
int pimc_structure_create(pimc_handle *h, int32_t kmax, int32_t *id);
int pimc_structure_measure(pimc_handle *h, int32_t id);                                        /* functor call now, every chain */
int pimc_structure_read(pimc_handle *h, int32_t id, double *sums, int64_t *ndata, int32_t *kmax); /* summed over chains (and ranks) */
int pimc_compressibility(pimc_handle *h, int32_t id, double *kappa, double *s_kmin);
/* run! with the full Zmeasurements list (src/simulation.jl:29-42): pimc_run plus the estimators above, same cadence */
typedef struct {
    const int32_t *energy_ids; int32_t nenergy;
    const int32_t *density_ids; int32_t ndensity;
    const int32_t *paircorr_ids; int32_t npaircorr;
    const int32_t *winding_ids; int32_t nwinding;
    const int32_t *structure_ids; int32_t nstructure;
} pimc_measurements;
int pimc_run_ex(pimc_handle *h, int64_t n, const int32_t *update_ids, const int64_t *every, int32_t nupd,
                const pimc_measurements *meas, int32_t sched, pimc_run_stats *stats);

/* ---- multi-GPU: chains shard over ranks (one handle per GPU, chain_offset = first global chain id of the shard), no
 *      data-path collective; the reference's analogue is `pmap` over independent runs (examples/density_SRL_lattice.jl:1-2,46).  With a
 *      communicator attached the library all-reduces the estimator accumulators itself (NCCL over NVLink, on a side stream: the block
 *      reduced at the end of pimc_run overlaps the next block's moves) and the read-outs become collective and global:
 *      pimc_energy_read* with chain = -1 -> mean over the chains of ALL ranks; pimc_density_read / pimc_paircorr_read / pimc_structure_read
 *      (and pimc_compressibility, which reads the structure factor) -> counters / sums and ndata summed over the ranks.
 *      Every rank must issue the same read-outs in the same order.  One rank per process (torchrun, MPI, Julia Distributed workers): rank 0
 *      calls pimc_comm_get_unique_id and ships the 128 bytes to the others, every rank calls pimc_comm_init.  One process, several GPUs:
 *      pimc_comm_init_all on handles created on different devices, then one host thread per handle. ---- */
#define PIMC_COMM_ID_BYTES 128
int pimc_comm_get_unique_id(void *id128);
int pimc_comm_init(pimc_handle *h, int32_t nranks, int32_t rank, const void *id128);
int pimc_comm_init_all(pimc_handle **handles, int32_t n);
int pimc_comm_info(pimc_handle *h, int32_t *nranks, int32_t *rank, int64_t *chains_total, int32_t *nccl_version);

/* ---- pair propagator of interacting Systems: host-side construction, no GPU needed (csrc/pimc_propint.cu) ---- */
/* prop_rel_interpolate_terms (src/propagator.jl:34-70): the sampled term table on range(1e-20, L, delta)^2 (delta = 600 in the reference),
 * column-major delta x delta = the `tab` of pimc_config; *lo = 1e-20, *hi = L */
int pimc_build_prop_table(double L, double g0, double tau, int32_t delta, double *tab, double *lo, double *hi);
/* prop_int(r1_rel, r2_rel, tau) = 1 + terms(|r1|, |r2|) / prop_rel0(r1, r2, tau), the closure build_prop_int returns (src/propagator.jl:73-89) */
int pimc_prop_int(const double *tab, int32_t n, double lo, double hi, const double *r1_rel, const double *r2_rel, int32_t dim, double tau, double *out);
/* determine_nnrange(propint, tau, a, b) (src/system.jl:10-15); pimc_create calls it with (1e-20, L) when interactions != 0 and r_a == 0 (system.jl:29-31) */
int pimc_determine_nnrange(const double *tab, int32_t n, double lo, double hi, double tau, double a, double b, double *r_a);

/* ---- run! (src/simulation.jl:29-42) on every chain ---- */
int pimc_run(pimc_handle *h, int64_t n, const int32_t *update_ids, const int64_t *every, int32_t nupd,
             const int32_t *energy_ids, int32_t nen, const int32_t *density_ids, int32_t nde, int32_t sched,
             pimc_run_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
