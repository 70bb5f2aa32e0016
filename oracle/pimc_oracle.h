/*
 * pimc_oracle.h -- CPU ORACLE for the pimc-b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the algorithm of oameye/PIMC.jl (reference at /root/reference,
 * pure Julia).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (pimc_jl_b200/) never does.
 *
 * PARITY UNPINNED: Julia is not installed in this image (and cannot be), the reference
 * ships no golden vectors / known-answer tests for this path, and its RNG is
 * unseeded.  The oracle is pinned instead by (i) an independent numpy/mpmath re-derivation
 * of levy!/teleport/Energy/Density in tests/, (ii) the analytic constants the reference's own
 * example and tests hold (exact = 2.1639534137386534, density normalisation, box invariant,
 * lattice-intensity peaks) and (iii) closed-form finite-M harmonic energies.
 *
 * Every function cites the reference file:line it follows.  Indices in this API are
 * 1-based wherever the reference's are (particles, slices, bins).
 */
#ifndef PIMC_ORACLE_H
#define PIMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* potential descriptor (closures of the reference cannot cross a C ABI) */
enum { ORA_POT_ZERO = 0, ORA_POT_HARMONIC = 1, ORA_POT_SIN2_1D = 2, ORA_POT_LATTICE = 3 };
enum { ORA_DV_ZERO = 0, ORA_DV_IDENTITY = 1, ORA_DV_GRADIENT = 2 };
#define ORA_MAX_ANGLES 32
typedef struct {
    int32_t kind, dv_kind;
    double k;                 /* harmonic: V = (0.5*k)*(x^2+y^2)                      */
    double depth, scale, sgn; /* sin2: depth*sin(2pi*x*scale)^2 ; lattice: sgn*depth*I */
    int32_t nang, helical;
    double ang[ORA_MAX_ANGLES];
} ora_potential;

/* compat flags: set = reproduce the reference as shipped */
#define ORA_COMPAT_PAIR_BYVALUE  1 /* B3: interaction_action! mutates a by-value Float64 -> no effect */
#define ORA_COMPAT_SWAP_SIGN     2 /* B4: new-configuration lnU added to w_initial                   */
#define ORA_COMPAT_DENSITY_SHIFT 4 /* B5: Density drops floor-bin 0, shifted by one bin               */
#define ORA_COMPAT_SWAP_STALE_LINK 8 /* B14: accepted swap leaves the cached link at slice j_m un-exchanged */
#define ORA_COMPAT_ALL           15

typedef struct {
    int32_t dim, M, N;
    double mu, lambda, L, T;
    int32_t interactions;
    double g, r_a;
    int32_t Ncycle;
    int32_t compat;
    int32_t init;             /* 1: init_world from the stream ; 0: all beads at 0 (set_paths later) */
    uint64_t seed;
    uint32_t chain;
    ora_potential pot;
    const double *tab;        /* integral-term table of prop_rel_interpolate_terms, tab_n x tab_n, column-major */
    int32_t tab_n;
    double tab_lo, tab_hi;
} ora_config;

typedef struct ora_system ora_system;
typedef struct ora_update ora_update;
typedef struct ora_energy ora_energy;
typedef struct ora_density ora_density;

enum { ORA_UPD_RESHAPE_LINEAR = 0, ORA_UPD_RESHAPE_SWAP = 1, ORA_UPD_SINGLE_COM = 2, ORA_UPD_POLYMER_COM = 3 };
/* ORA_SCHED_SWEEP_SEQ: the sweep executed strictly sequentially (proposal n sees the committed results of the proposals before it), also for
 * systems with a hard core / pair action -- the definition a future GPU sweep for interacting worldlines is held to.
 * ORA_SCHED_SWEEP is the same loop restricted to independent worldlines, where the order cannot matter. */
enum { ORA_SCHED_FAITHFUL = 0, ORA_SCHED_SWEEP = 1, ORA_SCHED_SWEEP_SEQ = 2 };

/* ---- pure functions ---- */
double ora_distance(double x1, double x2, double L);
double ora_teleport(double x, double L);
double ora_lnK(const double *r1, const double *r2, int dim, double tau, double lambda, double L);
double ora_prop_0(const double *r1, const double *r2, int dim, double tau, double lambda, double L);
double ora_potential_eval(const ora_potential *p, const double *r, int dim);
void   ora_potential_grad(const ora_potential *p, const double *r, int dim, double *dv);
double ora_lnV(const double *r1, const double *r2, int dim, double tau, const ora_potential *p);
void   ora_levy(double *r, int rows, int dim, double tau, double L, double lambda, const double *xi);
int    ora_metropolis(double delta, double u);
int64_t ora_bin(const double *r, int dim, int64_t nbins, double L);
void   ora_bin_neighbors(int64_t b, int64_t nbins, int dim, int64_t *out /* 9 or 3 */);
int64_t ora_pcycle(int64_t j, const int64_t *pol, int64_t Npol, int64_t M);
double ora_adjust_step(double size, double minstep, double maxstep, double minacc, double maxacc, double acc);
int64_t ora_adjust_slices(int64_t m, int64_t minslices, int64_t maxslices, double minacc, double maxacc, double acc);
double ora_prop_rel0(const double *r1, const double *r2, int dim, double tau);
void   ora_gauss_pair(uint64_t seed, uint32_t chain, uint64_t iter, uint32_t slot, uint32_t kind,
                      uint32_t retry, uint32_t bead, double *g0, double *g1);

void   ora_uniform_pair(uint64_t seed, uint32_t chain, uint64_t iter, uint32_t slot, uint32_t kind,
                        uint32_t retry, uint32_t bead, double *u0, double *u1);

/* ---- system ---- */
ora_system *ora_create(const ora_config *cfg);
void ora_destroy(ora_system *s);
const char *ora_last_error(void);
void ora_get_paths(const ora_system *s, double *r, double *V, int64_t *bins, int64_t *next);
void ora_set_paths(ora_system *s, const double *r, const int64_t *next); /* recomputes link cache + nn grid */
void ora_get_scalars(const ora_system *s, double *out /* beta,tau,vol,a,r_a */, int64_t *iout /* nbins,N_MC,Nctr,ctr,iter */);
void ora_set_iter(ora_system *s, uint64_t iter);
void ora_set_ctr(ora_system *s, int64_t ctr);
void ora_update_nnbins(ora_system *s);
int64_t ora_subcycle(const ora_system *s, int64_t n, int64_t *cycle);
int64_t ora_cycle_findprev(const ora_system *s, int64_t n);
int64_t ora_find_nn(const ora_system *s, const double *r, int64_t j, const int64_t *exc, int nexc);
int64_t ora_find_nns_pos(const ora_system *s, const double *r, int64_t j, const int64_t *exc, int nexc, int64_t *out);
int64_t ora_find_nns_idx(const ora_system *s, int64_t i, int64_t j, const int64_t *exc, int nexc, int64_t *out);
int64_t ora_nn_cell(const ora_system *s, int64_t j, int64_t b, int64_t *out);
double ora_lnU(const ora_system *s, const double *r1, const double *r2);
double ora_action_links(const ora_system *s);   /* sum of the link cache          */
double ora_action_links_recomputed(const ora_system *s);
double ora_action_pairs(const ora_system *s);   /* intended pair action: sum over slices and pairs of lnU */

/* ---- updates ---- */
ora_update *ora_update_create(const ora_system *s, int kind, double var0);
void ora_update_configure(ora_update *u, double vmin, double vmax, double minacc, double maxacc, int64_t adj, int64_t range);
void ora_update_destroy(ora_update *u);
void ora_update_get(const ora_update *u, double *var, int64_t *tries, int64_t *tries_var, double *acc_window,
                    int64_t *accepted, int64_t *bead_moves);
/* one functor call with the addressed stream; forced_n / forced_j0 = 0 -> drawn. returns acc. */
int ora_update_call(ora_system *s, ora_update *u, uint32_t slot, int64_t forced_n, int64_t forced_j0);
/* explicit-input variants (no stream): deterministic parity hooks */
int ora_reshape_linear_explicit(ora_system *s, int64_t n, int64_t j0, int64_t m, const double *xi, double u,
                                int commit, double *w_initial, double *w_updated, double *rprime);
int ora_reshape_swap_explicit(ora_system *s, int64_t n1, int64_t n2, int64_t j0, int64_t m, const double *xi1,
                              const double *xi2, double u, int commit, double *w_initial, double *w_updated);
int ora_com_explicit(ora_system *s, int64_t n, int polymer, const double *d, double u, int commit,
                     double *w_initial, double *w_updated);
void ora_swap_weights(const ora_system *s, int64_t n1, int64_t j0, int64_t m, double *w /* N */);

/* ---- measurements ---- */
ora_energy *ora_energy_create(int64_t cap);
void ora_energy_destroy(ora_energy *e);
int64_t ora_energy_read(const ora_energy *e, double *E, double *Ev, int64_t cap);
void ora_energy_now(const ora_system *s, double *E, double *Ev, double *parts /* link,pot,vkin */);
ora_density *ora_density_create(const ora_system *s, int64_t nbins);
void ora_density_destroy(ora_density *d);
void ora_density_measure(ora_density *d, const ora_system *s);
int64_t ora_density_read(const ora_density *d, double *dens, double *bin);

/* estimators the reference lists as TODO (measurement.jl:125-127): radial distribution and winding number, definitions in pimc_oracle.c */
typedef struct ora_paircorr ora_paircorr;
ora_paircorr *ora_paircorr_create(const ora_system *s, int64_t nbins, double rmax);
void ora_paircorr_destroy(ora_paircorr *g);
void ora_paircorr_measure(ora_paircorr *g, const ora_system *s);
int64_t ora_paircorr_read(const ora_paircorr *g, double *hist, double *bin);
void ora_winding_now(const ora_system *s, double *W /* dim */);
/* static structure factor sums of the current configuration (definition in pimc_oracle.c): out[(kmax + 1) * (2 kmax + 1)] */
void ora_structure_now(const ora_system *s, int kmax, double *out);

/* ---- driver ---- */
int ora_run(ora_system *s, int64_t n, ora_update **upd, const int64_t *every, int nupd,
            ora_energy **en, int nen, ora_density **de, int nde, int sched);

#ifdef __cplusplus
}
#endif
#endif
