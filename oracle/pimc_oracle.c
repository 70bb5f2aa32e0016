/*
 * pimc_oracle.c -- CPU ORACLE (test infrastructure, see pimc_oracle.h).  PARITY UNPINNED
 * against a running reference (no Julia in this image); pinned by tests/ second opinions.
 *
 * Restates, function by function, oameye/PIMC.jl:
 *   src/propagator.jl:6-32,73-89       distance, lnK, prop_0, lnV, teleport, prop_rel0, prop_int lookup
 *   src/system.jl:17-91,129-167        init_int, init_world, init_nn, System constructor
 *   src/nearest_neighbours.jl:8-246    bin grid, stencil, find_nn(s), maintenance
 *   src/updates/helper.jl:3-395        metropolis, Counter/Step/NumbOfSlices, cycles, levy!, hardspherelevy!,
 *                                      sampleparticles, prev/next, interaction_action!, move_polymer!
 *   src/updates/com.jl:31-104,136-224  PolymerCenterOfMass (worms = 0 reading), SingleCenterOfMass
 *   src/updates/reshape.jl:31-91,123-283 ReshapeLinear, ReshapeSwapLinear
 *   src/measurement.jl:1-17,45-55,92-122 measurement cadence, Density, Energy
 *   src/simulation.jl:1-42             acceptance, queue!, apply!, run!
 * Julia semantics mirrored: 1-based indices, mod1, floor toward -inf, sign(0)=0, no FMA contraction
 * (compile with -ffp-contract=off), left-to-right evaluation of a*b*c and a+b+c.
 * Random draws come from the addressed Philox stream of include/pimc_rng.h (Julia's RNG cannot be reproduced).
 */
#include "pimc_oracle.h"
#include "../include/pimc_rng.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

static char g_err[256] = "";
__attribute__((constructor)) static void ora_init_tables(void) { (void)pimc_logtab_get(); }
const char *ora_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); } while (0)

static inline int64_t mod1(int64_t x, int64_t y) { int64_t m = ((x - 1) % y + y) % y; return m + 1; }
static inline int64_t imod(int64_t x, int64_t y) { return ((x % y) + y) % y; }
static inline double jl_sign(double x) { return (x > 0) - (x < 0); }

typedef struct { int64_t *v; int32_t len, cap; } ivec;
static void iv_push(ivec *a, int64_t x)
{
    if (a->len == a->cap) { a->cap = a->cap ? 2 * a->cap : 4; a->v = (int64_t *)realloc(a->v, sizeof(int64_t) * a->cap); }
    a->v[a->len++] = x;
}
static void iv_filter_ne(ivec *a, int64_t x) /* filter!(y -> y != x, a) */
{
    int32_t o = 0;
    for (int32_t i = 0; i < a->len; ++i) if (a->v[i] != x) a->v[o++] = a->v[i];
    a->len = o;
}

struct ora_system {
    int dim, M, N, Ninit;
    double mu, lambda, L, vol, beta, tau;
    double *r;      /* [N][dim][M] : per particle the reference's M x dim column-major matrix */
    double *V;      /* [N][M] link cache                                                       */
    int64_t *bins;  /* [N][M] 1-based bin ids                                                  */
    int64_t *next;  /* [N]    1-based                                                          */
    ivec *nn;       /* [M][ncell]                                                              */
    int64_t *nbs;   /* [ncell][nst]                                                            */
    int64_t nbins, ncell; int nst;
    ora_potential pot;
    double a, r_a; int64_t ctr;
    int interactions;
    double *tab; int tab_n; double tab_lo, tab_hi;
    int64_t N_MC, Nctr, Ncycle;
    int compat; uint64_t seed; uint32_t chain; uint64_t iter;
};
#define R_(s, n, j, k) ((s)->r[(((int64_t)(n) - 1) * (s)->dim + ((k) - 1)) * (s)->M + ((j) - 1)])
#define V_(s, n, j)    ((s)->V[((int64_t)(n) - 1) * (s)->M + ((j) - 1)])
#define B_(s, n, j)    ((s)->bins[((int64_t)(n) - 1) * (s)->M + ((j) - 1)])
#define NN_(s, j, b)   ((s)->nn[((int64_t)(j) - 1) * (s)->ncell + ((b) - 1)])

/* ================= src/propagator.jl ================= */
/* propagator.jl:6-9 */
double ora_distance(double x1, double x2, double L)
{
    double dx = fabs(x1 - x2);
    double alt = (2 * L) - dx;
    return alt < dx ? alt : dx; /* Julia min */
}
/* propagator.jl:30-32 */
double ora_teleport(double x, double L)
{
    return ((x + L) - floor(x / (2 * L) + 0.5) * (2 * L)) - L;
}
/* propagator.jl:16-19 */
double ora_lnK(const double *r1, const double *r2, int dim, double tau, double lambda, double L)
{
    double d2 = 0.0;
    for (int k = 0; k < dim; ++k) { double dr = ora_distance(r1[k], r2[k], L); d2 = (k == 0) ? dr * dr : d2 + dr * dr; }
    return -d2 / (4 * lambda * tau);
}
/* propagator.jl:20-24 */
double ora_prop_0(const double *r1, const double *r2, int dim, double tau, double lambda, double L)
{
    return exp(ora_lnK(r1, r2, dim, tau, lambda, L));
}
/* potential closures of the example scripts, test/testsystem.jl:13 and examples/tools/potentialtools.jl:1-16,25-39 */
static double lattice_intensity(const ora_potential *p, const double *c)
{
    double s = 0.0, cc = 0.0;
    for (int i = 0; i < p->nang; ++i) {
        double ang = p->ang[i];
        double rr = c[0] * sin(ang) + c[1] * cos(ang);
        double ph = 2 * M_PI * rr * p->scale;
        if (p->helical) ph = ph + ang;
        s += sin(ph);
        cc += cos(ph);
    }
    s /= p->nang;
    cc /= p->nang;
    return s * s + cc * cc;
}
double ora_potential_eval(const ora_potential *p, const double *r, int dim)
{
    switch (p->kind) {
    case ORA_POT_ZERO: return 0.0;
    case ORA_POT_HARMONIC: { /* (r) -> 0.5*(r[1]^2+r[2]^2) */
        double s = r[0] * r[0];
        for (int k = 1; k < dim; ++k) s = s + r[k] * r[k];
        return (0.5 * p->k) * s; }
    case ORA_POT_SIN2_1D: { /* depth * sin(2pi * x[1] * scale)^2 , test/testsystem.jl:13 */
        double sn = sin(2 * M_PI * r[0] * p->scale);
        return p->depth * (sn * sn); }
    case ORA_POT_LATTICE: { /* sgn*depth * normalized_intensity(coord, ang, scale) */
        double c[2] = { r[0], dim > 1 ? r[1] : 0.0 };
        return (p->sgn * p->depth) * lattice_intensity(p, c); }
    }
    return 0.0;
}
void ora_potential_grad(const ora_potential *p, const double *r, int dim, double *dv)
{
    for (int k = 0; k < dim; ++k) dv[k] = 0.0;
    if (p->dv_kind == ORA_DV_ZERO) return;                  /* dV = zero  (system.jl:131) */
    if (p->dv_kind == ORA_DV_IDENTITY) { for (int k = 0; k < dim; ++k) dv[k] = r[k]; return; } /* dV = identity */
    switch (p->kind) { /* analytic gradient ("intended" virial estimator) */
    case ORA_POT_HARMONIC: for (int k = 0; k < dim; ++k) dv[k] = p->k * r[k]; break;
    case ORA_POT_SIN2_1D: { double ph = 2 * M_PI * r[0] * p->scale; dv[0] = p->depth * 2 * sin(ph) * cos(ph) * (2 * M_PI * p->scale); break; }
    case ORA_POT_LATTICE: {
        double s = 0, c = 0, sx = 0, sy = 0, cx = 0, cy = 0;
        double y = dim > 1 ? r[1] : 0.0;
        for (int i = 0; i < p->nang; ++i) {
            double ang = p->ang[i], sa = sin(ang), ca = cos(ang);
            double ph = 2 * M_PI * (r[0] * sa + y * ca) * p->scale + (p->helical ? ang : 0.0);
            double f = 2 * M_PI * p->scale;
            s += sin(ph); c += cos(ph);
            sx += cos(ph) * f * sa; sy += cos(ph) * f * ca;
            cx += -sin(ph) * f * sa; cy += -sin(ph) * f * ca;
        }
        double n = p->nang; s /= n; c /= n; sx /= n; sy /= n; cx /= n; cy /= n;
        dv[0] = p->sgn * p->depth * 2 * (s * sx + c * cx);
        if (dim > 1) dv[1] = p->sgn * p->depth * 2 * (s * sy + c * cy);
        break; }
    default: break;
    }
}
/* propagator.jl:26-28 */
double ora_lnV(const double *r1, const double *r2, int dim, double tau, const ora_potential *p)
{
    return -0.5 * tau * (ora_potential_eval(p, r1, dim) + ora_potential_eval(p, r2, dim));
}
/* propagator.jl:73-76 */
double ora_prop_rel0(const double *r1, const double *r2, int dim, double tau)
{
    double d2 = 0.0;
    for (int k = 0; k < dim; ++k) { double d = r1[k] - r2[k]; d2 = (k == 0) ? d * d : d2 + d * d; }
    return exp(-d2 / (4 * tau)) / (4 * M_PI * tau);
}
static double vnorm(const double *v, int dim)
{
    double s = v[0] * v[0];
    for (int k = 1; k < dim; ++k) s = s + v[k] * v[k];
    return sqrt(s);
}
/* scaled BSpline(Linear()) lookup of propagator.jl:64-67 on a uniform tab_n x tab_n grid */
static double tab_lookup(const ora_system *s, double x, double y)
{
    int n = s->tab_n;
    double h = (s->tab_hi - s->tab_lo) / (n - 1);
    double tx = (x - s->tab_lo) / h, ty = (y - s->tab_lo) / h;
    double fx0 = floor(tx), fy0 = floor(ty);
    if (fx0 < 0) fx0 = 0; if (fx0 > n - 2) fx0 = n - 2;
    if (fy0 < 0) fy0 = 0; if (fy0 > n - 2) fy0 = n - 2;
    int ix = (int)fx0, iy = (int)fy0;
    double fx = tx - fx0, fy = ty - fy0;
    const double *A = s->tab;
    double a00 = A[ix + (int64_t)n * iy], a10 = A[ix + 1 + (int64_t)n * iy];
    double a01 = A[ix + (int64_t)n * (iy + 1)], a11 = A[ix + 1 + (int64_t)n * (iy + 1)];
    double c0 = (1 - fx) * a00 + fx * a10;
    double c1 = (1 - fx) * a01 + fx * a11;
    return (1 - fy) * c0 + fy * c1;
}
/* system.jl:17-34 (lnU closure) with propagator.jl:82-86 (prop_int) */
double ora_lnU(const ora_system *s, const double *r1, const double *r2)
{
    if (!s->interactions || !s->tab) return 0.0;
    double p = 1 + tab_lookup(s, vnorm(r1, s->dim), vnorm(r2, s->dim)) / ora_prop_rel0(r1, r2, s->dim, s->tau);
    return p < 0.0 ? -s->mu : log(p);
}

/* ================= src/nearest_neighbours.jl ================= */
/* nearest_neighbours.jl:8-33 ; out-of-range bins (r == +L after rounding) are clamped (reference: BoundsError) */
int64_t ora_bin(const double *r, int dim, int64_t nbins, double L)
{
    double w = 2 * L / nbins;
    int64_t ib[2] = { 0, 0 };
    for (int k = 0; k < dim; ++k) {
        int64_t i = (int64_t)floor((r[k] + L) / w);
        if (i < 0) i = 0;
        if (i > nbins - 1) i = nbins - 1;
        ib[k] = i;
    }
    return dim == 2 ? ib[0] + nbins * ib[1] + 1 : ib[0] + 1;
}
/* nearest_neighbours.jl:55-65 */
void ora_bin_neighbors(int64_t b, int64_t nbins, int dim, int64_t *out)
{
    if (dim == 2) {
        static const int dxy[9][2] = { {0,0}, {-1,1}, {0,1}, {1,1}, {-1,0}, {1,0}, {-1,-1}, {0,-1}, {1,-1} };
        int64_t x = (b - 1) % nbins, y = (b - 1) / nbins;
        for (int i = 0; i < 9; ++i) out[i] = imod(x + dxy[i][0], nbins) + nbins * imod(y + dxy[i][1], nbins) + 1;
    } else {
        static const int dx[3] = { 0, -1, 1 };
        int64_t x = (b - 1) % nbins;
        for (int i = 0; i < 3; ++i) out[i] = imod(x + dx[i], nbins) + 1;
    }
}
/* nearest_neighbours.jl:182-196 (also psort!, :40-52, and init_nn's fill, system.jl:80-91) */
void ora_update_nnbins(ora_system *s)
{
    for (int64_t m = 1; m <= s->M; ++m) {
        for (int64_t b = 1; b <= s->ncell; ++b) NN_(s, m, b).len = 0;
        for (int64_t n = 1; n <= s->N; ++n) {
            double c[2]; for (int k = 1; k <= s->dim; ++k) c[k - 1] = R_(s, n, m, k);
            int64_t b = ora_bin(c, s->dim, s->nbins, s->L);
            iv_push(&NN_(s, m, b), n);
            B_(s, n, m) = b;
        }
    }
}
/* nearest_neighbours.jl:198-209 */
static void update_nn_bead(ora_system *s, int64_t n, int64_t j, const double *rp)
{
    iv_filter_ne(&NN_(s, j, B_(s, n, j)), n);
    int64_t b = ora_bin(rp, s->dim, s->nbins, s->L);
    iv_push(&NN_(s, j, b), n);
    B_(s, n, j) = b;
}
/* nearest_neighbours.jl:211-217 */
static void add_nn(ora_system *s, const int64_t *pol, int64_t Npol)
{
    for (int64_t i = 0; i < Npol; ++i) for (int64_t j = 1; j <= s->M; ++j) iv_push(&NN_(s, j, B_(s, pol[i], j)), pol[i]);
}
/* nearest_neighbours.jl:229-235 */
static void rm_nn(ora_system *s, const int64_t *pol, int64_t Npol)
{
    for (int64_t i = 0; i < Npol; ++i) for (int64_t j = 1; j <= s->M; ++j) iv_filter_ne(&NN_(s, j, B_(s, pol[i], j)), pol[i]);
}
static int in_list(int64_t x, const int64_t *l, int n) { for (int i = 0; i < n; ++i) if (l[i] == x) return 1; return 0; }
/* nearest_neighbours.jl:72-104 : vcat of the stencil cells minus exceptions (duplicates kept as the reference keeps them) */
static int64_t within_neighbourhood(const ora_system *s, int64_t b, int64_t j, const int64_t *exc, int nexc, int64_t *out)
{
    int64_t cnt = 0;
    for (int i = 0; i < s->nst; ++i) {
        const ivec *c = &NN_(s, j, s->nbs[(b - 1) * s->nst + i]);
        for (int32_t q = 0; q < c->len; ++q) if (!in_list(c->v[q], exc, nexc)) out[cnt++] = c->v[q];
    }
    return cnt;
}
int64_t ora_nn_cell(const ora_system *s, int64_t j, int64_t b, int64_t *out)
{
    const ivec *c = &NN_(s, j, b);
    for (int32_t q = 0; q < c->len; ++q) out[q] = c->v[q];
    return c->len;
}
/* Distances.PeriodicEuclidean(2L) on coordinates shifted by +L (nearest_neighbours.jl:122-127,147-152,172-177) */
static double periodic_euclid(const ora_system *s, const double *a, const double *b)
{
    double p = 2 * s->L, acc = 0.0;
    for (int k = 0; k < s->dim; ++k) {
        double s1 = fabs((a[k] + s->L) - (b[k] + s->L));
        double s2 = s1 - p * floor(s1 / p);
        double s3 = s2 < p - s2 ? s2 : p - s2;
        acc = (k == 0) ? s3 * s3 : acc + s3 * s3;
    }
    return sqrt(acc);
}
static int64_t max_cand(const ora_system *s) { return (int64_t)s->nst * s->N * 2 + 16; }
/* nearest_neighbours.jl:156-179 */
int64_t ora_find_nn(const ora_system *s, const double *r, int64_t j, const int64_t *exc, int nexc)
{
    int64_t *cand = (int64_t *)malloc(sizeof(int64_t) * max_cand(s));
    int64_t nc = within_neighbourhood(s, ora_bin(r, s->dim, s->nbins, s->L), j, exc, nexc, cand);
    int64_t best = -1; double bd = 0.0;
    for (int64_t i = 0; i < nc; ++i) {
        double c[2]; for (int k = 1; k <= s->dim; ++k) c[k - 1] = R_(s, cand[i], j, k);
        double d = periodic_euclid(s, c, r);
        if (best < 0 || d < bd) { best = cand[i]; bd = d; }
    }
    free(cand);
    return best;
}
static int64_t find_nns_core(const ora_system *s, const double *target, int64_t b, int64_t j, const int64_t *exc, int nexc, int64_t *out)
{
    int64_t *cand = (int64_t *)malloc(sizeof(int64_t) * max_cand(s));
    int64_t nc = within_neighbourhood(s, b, j, exc, nexc, cand), cnt = 0;
    double rad = (2 * s->L) / s->nbins;
    for (int64_t i = 0; i < nc; ++i) {
        double c[2]; for (int k = 1; k <= s->dim; ++k) c[k - 1] = R_(s, cand[i], j, k);
        if (periodic_euclid(s, c, target) <= rad) out[cnt++] = cand[i];
    }
    free(cand);
    return cnt;
}
/* nearest_neighbours.jl:131-154 */
int64_t ora_find_nns_pos(const ora_system *s, const double *r, int64_t j, const int64_t *exc, int nexc, int64_t *out)
{
    return find_nns_core(s, r, ora_bin(r, s->dim, s->nbins, s->L), j, exc, nexc, out);
}
/* nearest_neighbours.jl:106-129 (uses the STORED bin of particle i) */
int64_t ora_find_nns_idx(const ora_system *s, int64_t i, int64_t j, const int64_t *exc, int nexc, int64_t *out)
{
    double c[2]; for (int k = 1; k <= s->dim; ++k) c[k - 1] = R_(s, i, j, k);
    return find_nns_core(s, c, B_(s, i, j), j, exc, nexc, out);
}

/* ================= src/updates/helper.jl ================= */
/* helper.jl:3-5 ; u is the uniform that rand() would return, consumed only if delta < 1 */
int ora_metropolis(double delta, double u) { return (delta >= 1.0) || (delta > u); }

typedef struct { uint8_t *ring; int64_t range, head, len, sum, tries, adj; } counter_t;
static void counter_init(counter_t *c, int64_t range, int64_t adj)
{
    c->ring = (uint8_t *)calloc((size_t)range + 1, 1); c->range = range; c->head = 0; c->len = 0; c->sum = 0; c->tries = 0; c->adj = adj;
}
/* simulation.jl:3-10 */
static void queue_push(counter_t *c, int acc)
{
    c->tries += 1;
    int64_t cap = c->range + 1;
    c->ring[(c->head + c->len) % cap] = (uint8_t)(acc ? 1 : 0);
    c->len += 1; c->sum += acc ? 1 : 0;
    if (c->len > c->range) { c->sum -= c->ring[c->head]; c->head = (c->head + 1) % cap; c->len -= 1; }
}
/* simulation.jl:1 ; empty queue -> 0/0 = NaN */
static double acceptance(const counter_t *c) { return (double)c->sum / (double)c->len; }
/* helper.jl:22-32 */
double ora_adjust_step(double size, double minstep, double maxstep, double minacc, double maxacc, double acc)
{
    if (acc < minacc) size *= 0.9; else if (acc > maxacc) size *= 1.1;
    size = minstep > size ? minstep : size;
    size = maxstep < size ? maxstep : size;
    return size;
}
/* helper.jl:42-52 */
int64_t ora_adjust_slices(int64_t m, int64_t minslices, int64_t maxslices, double minacc, double maxacc, double acc)
{
    if (acc < minacc) m -= 1; else if (acc > maxacc) m += 1;
    m = minslices > m ? minslices : m;
    m = maxslices < m ? maxslices : m;
    return m;
}
/* helper.jl:55-62 */
int64_t ora_cycle_findprev(const ora_system *s, int64_t n)
{
    for (int64_t i = 1; i <= s->N; ++i) if (s->next[i - 1] != 0 && n == s->next[i - 1]) return i;
    return 0;
}
/* helper.jl:64-85 */
int64_t ora_subcycle(const ora_system *s, int64_t n, int64_t *cycle)
{
    int64_t Ncycle = 1; cycle[0] = n;
    if (s->next[n - 1] != 0 && s->next[n - 1] != n) {
        int64_t i = n, ctr = 0;
        for (;;) {
            ctr += 1; i = s->next[i - 1];
            if (i == 0 || i == n) break;
            cycle[Ncycle++] = i;
            if (ctr > s->N) { FAIL("subcycle: broken permutation"); break; }
        }
    }
    return Ncycle;
}
/* helper.jl:113-115 */
int64_t ora_pcycle(int64_t j, const int64_t *pol, int64_t Npol, int64_t M)
{
    int64_t q = (j - 1) >= 0 ? (j - 1) / M : -((-(j - 1) + M - 1) / M);
    return pol[mod1(1 + q, Npol) - 1];
}
/* helper.jl:287-300 */
static void next_bead(const ora_system *s, int64_t n0, int64_t j0, int64_t *n, int64_t *j)
{
    if (n0 == 0) { *n = 0; *j = 0; return; }
    *n = (j0 == s->M) ? s->next[n0 - 1] : n0;
    *j = mod1(j0 + 1, s->M);
    if (*n == 0) { *n = 0; *j = 0; }
}
/* helper.jl:118-139 ; r is rows x dim COLUMN-major, xi is (rows-2) x dim ROW-major */
void ora_levy(double *r, int rows, int dim, double tau, double L, double lambda, const double *xi)
{
    for (int k = 0; k < dim; ++k)
        if (fabs(r[k * rows] - r[k * rows + rows - 1]) > L) r[k * rows + rows - 1] += jl_sign(r[k * rows]) * (2 * L);
    int m = rows - 2;
    for (int j = 1; j <= m; ++j) {
        double alpha = (double)(m + 1 - j) / (double)(m + 2 - j);
        double sig = sqrt(2 * lambda * alpha * tau);
        for (int k = 0; k < dim; ++k)
            r[k * rows + j] = alpha * r[k * rows + j - 1] + (1 - alpha) * r[k * rows + rows - 1] + xi[(j - 1) * dim + k] * sig;
    }
    for (int j = 0; j < rows; ++j) for (int k = 0; k < dim; ++k) r[k * rows + j] = ora_teleport(r[k * rows + j], L);
}

/* Gaussian source: explicit array (row-major (bead, dim), retries unsupported) or addressed stream */
typedef struct { const double *xi; pimc_stream st; uint32_t slot, kind; int dim; } gsrc;
static void gs_get(const gsrc *g, int bead, int retry, double *out)
{
    if (g->xi) { for (int k = 0; k < g->dim; ++k) out[k] = g->xi[(bead - 1) * g->dim + k]; return; }
    double g0, g1;
    pimc_gauss_pair(pimc_draw(g->st, g->slot, g->kind, (uint32_t)retry, (uint32_t)bead), &g0, &g1);
    out[0] = g0; if (g->dim > 1) out[1] = g1;
}
void ora_gauss_pair(uint64_t seed, uint32_t chain, uint64_t iter, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead, double *g0, double *g1)
{
    pimc_gauss_pair(pimc_draw(pimc_stream_make(seed, chain, iter), slot, kind, retry, bead), g0, g1);
}

/* the two uniforms in [0, 1) of one addressed draw (words 0-1 and 2-3): lets a test restate the retry loops that consume them */
void ora_uniform_pair(uint64_t seed, uint32_t chain, uint64_t iter, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead, double *u0, double *u1)
{
    pimc_u4 w = pimc_draw(pimc_stream_make(seed, chain, iter), slot, kind, retry, bead);
    *u0 = pimc_u01_co(w.w[0], w.w[1]); *u1 = pimc_u01_co(w.w[2], w.w[3]);
}
/* helper.jl:141-181 ; rp is rows x dim column-major; exc = fpcycle(mod1(j0+j, M)) = first particle of the closure's cycle */
static int hardspherelevy(double *rp, int rows, const ora_system *s, int64_t j0, int64_t exc, const gsrc *g)
{
    int dim = s->dim;
    for (int k = 0; k < dim; ++k)
        if (fabs(rp[k * rows] - rp[k * rows + rows - 1]) > s->L) rp[k * rows + rows - 1] += jl_sign(rp[k * rows]) * (2 * s->L);
    int m = rows - 2;
    for (int j = 1; j <= m; ++j) {
        double alpha = (double)(m + 1 - j) / (double)(m + 2 - j);
        double sig = sqrt(2 * s->lambda * alpha * s->tau);
        int pass = 1; int64_t ctr = 0;
        while (pass) {
            pass = 0; ctr += 1;
            if (ctr > s->ctr) { pass = 1; break; }
            double xi[2]; gs_get(g, j, (int)(ctr - 1), xi);
            for (int k = 0; k < dim; ++k)
                rp[k * rows + j] = alpha * rp[k * rows + j - 1] + (1 - alpha) * rp[k * rows + rows - 1] + xi[k] * sig;
            if (s->a > 0.0) { /* with a == 0 the test `norm(...) < a` can never hold: query skipped */
                double tp[2], dd[2]; int64_t sl = mod1(j0 + j, s->M);
                for (int k = 0; k < dim; ++k) tp[k] = ora_teleport(rp[k * rows + j], s->L);
                int64_t nn = ora_find_nn(s, tp, sl, &exc, 1);
                if (nn != -1) {
                    for (int k = 0; k < dim; ++k) dd[k] = ora_distance(tp[k], R_(s, nn, sl, k + 1), s->L);
                    if (vnorm(dd, dim) < s->a) pass = 1;
                }
            }
        }
        if (pass) return 0;
    }
    for (int j = 0; j < rows; ++j) for (int k = 0; k < dim; ++k) rp[k * rows + j] = ora_teleport(rp[k * rows + j], s->L);
    return 1;
}
static double lnK_beads(const ora_system *s, int64_t n1, int64_t j1, int64_t n2, int64_t j2, double tau)
{
    double a[2], b[2];
    for (int k = 1; k <= s->dim; ++k) { a[k - 1] = R_(s, n1, j1, k); b[k - 1] = R_(s, n2, j2, k); }
    /* system.jl:163 passes (lambda, tau) into the (tau, lambda) slots; only the product is used */
    return ora_lnK(a, b, s->dim, s->lambda, tau, s->L);
}
/* helper.jl:224-260 : the unnormalised table exp.(t + y) */
void ora_swap_weights(const ora_system *s, int64_t n1, int64_t j0, int64_t m, double *w)
{
    int64_t *pol = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1));
    int64_t jm = mod1(j0 + m, s->M);
    int64_t Np1 = ora_subcycle(s, n1, pol);
    int64_t n1next = ora_pcycle(j0 + m, pol, Np1, s->M);
    for (int64_t i = 1; i <= s->N; ++i) {
        int64_t Np = ora_subcycle(s, i, pol);
        int64_t inext = ora_pcycle(j0 + m, pol, Np, s->M);
        double t = lnK_beads(s, n1, j0, inext, jm, m * s->tau);
        double y = lnK_beads(s, i, j0, n1next, jm, m * s->tau);
        w[i - 1] = pimc_exp(t + y);
    }
    free(pol);
}
/* helper.jl:262-266 with StatsBase.sample(::AbstractWeights): t = rand()*sum(w); walk the cumulative sum */
static int64_t sample_weighted(const double *w, int64_t n, double u)
{
    double wsum = 0.0; for (int64_t i = 0; i < n; ++i) wsum = (i == 0) ? w[0] : wsum + w[i];
    double t = u * wsum; int64_t i = 1; double cw = w[0];
    while (cw < t && i < n) { i += 1; cw += w[i - 1]; }
    return i;
}

/* helper.jl:306-324 (old configuration, whole-bead variant) */
static double interaction_old(const ora_system *s, int64_t p, int64_t jw)
{
    double w = 0.0; int64_t *nl = (int64_t *)malloc(sizeof(int64_t) * max_cand(s));
    int64_t cnt = ora_find_nns_idx(s, p, jw, &p, 1, nl);
    for (int64_t q = 0; q < cnt; ++q) {
        int64_t nn = nl[q], nn_n, nn_j, p_n, p_j;
        next_bead(s, nn, jw, &nn_n, &nn_j); next_bead(s, p, jw, &p_n, &p_j);
        if (nn_n == 0 || p_n == 0) continue;
        double r1[2], r2[2];
        for (int k = 1; k <= s->dim; ++k) {
            r1[k - 1] = ora_distance(R_(s, nn, jw, k), R_(s, p, jw, k), s->L);
            r2[k - 1] = ora_distance(R_(s, nn_n, nn_j, k), R_(s, p_n, p_j, k), s->L);
        }
        w += ora_lnU(s, r1, r2);
    }
    free(nl);
    return w;
}
/* helper.jl:326-347 / :349-366 (new configuration: bead ra at slice jw, its successor rb), exceptions list;
 * skip_next_in_exc mirrors `nn_nextbead[1] in exceptions` of the second method only. */
static double interaction_new(const ora_system *s, const double *ra, const double *rb, int64_t jw, const int64_t *exc, int nexc, int skip_next_in_exc)
{
    double w = 0.0; int64_t *nl = (int64_t *)malloc(sizeof(int64_t) * max_cand(s));
    int64_t cnt = ora_find_nns_pos(s, ra, jw, exc, nexc, nl);
    for (int64_t q = 0; q < cnt; ++q) {
        int64_t nn = nl[q], nn_n, nn_j;
        next_bead(s, nn, jw, &nn_n, &nn_j);
        if (nn_n == 0 || (skip_next_in_exc && in_list(nn_n, exc, nexc))) continue;
        double r1[2], r2[2];
        for (int k = 1; k <= s->dim; ++k) {
            r1[k - 1] = ora_distance(R_(s, nn, jw, k), ra[k - 1], s->L);
            r2[k - 1] = ora_distance(R_(s, nn_n, nn_j, k), rb[k - 1], s->L);
        }
        w += ora_lnU(s, r1, r2);
    }
    free(nl);
    return w;
}

/* ================= updates ================= */
struct ora_update {
    int kind;
    counter_t counter, counter_var;
    double size, minstep, maxstep;       /* Step         helper.jl:14-20 */
    int64_t m, minslices, maxslices;     /* NumbOfSlices helper.jl:34-40 */
    double minacc, maxacc;
    int64_t accepted, bead_moves;
};
/* constructors: reshape.jl:12-28,104-120 ; com.jl:12-27,117-132 */
ora_update *ora_update_create(const ora_system *s, int kind, double var0)
{
    ora_update *u = (ora_update *)calloc(1, sizeof *u);
    u->kind = kind;
    counter_init(&u->counter, 100000, 10);
    counter_init(&u->counter_var, 10000, 10);
    if (kind == ORA_UPD_RESHAPE_LINEAR || kind == ORA_UPD_RESHAPE_SWAP) {
        u->minslices = 2; u->maxslices = s->M - 2; u->minacc = 0.6; u->maxacc = 0.8;
        int64_t sl = (int64_t)var0; u->m = sl < u->maxslices ? sl : u->maxslices;
    } else {
        u->minstep = 1e-1; u->maxstep = s->L / 2; u->minacc = 0.4; u->maxacc = 0.6; u->size = var0;
    }
    return u;
}
void ora_update_configure(ora_update *u, double vmin, double vmax, double minacc, double maxacc, int64_t adj, int64_t range)
{
    if (u->kind == ORA_UPD_RESHAPE_LINEAR || u->kind == ORA_UPD_RESHAPE_SWAP) {
        u->minslices = (int64_t)vmin; u->maxslices = (int64_t)vmax; if (u->m > u->maxslices) u->m = u->maxslices;
    } else { u->minstep = vmin; u->maxstep = vmax; }
    u->minacc = minacc; u->maxacc = maxacc;
    free(u->counter_var.ring); counter_init(&u->counter_var, range, adj);
}
void ora_update_destroy(ora_update *u) { if (!u) return; free(u->counter.ring); free(u->counter_var.ring); free(u); }
void ora_update_get(const ora_update *u, double *var, int64_t *tries, int64_t *tries_var, double *acc_window, int64_t *accepted, int64_t *bead_moves)
{
    *var = (u->kind == ORA_UPD_RESHAPE_LINEAR || u->kind == ORA_UPD_RESHAPE_SWAP) ? (double)u->m : u->size;
    *tries = u->counter.tries; *tries_var = u->counter_var.tries; *acc_window = acceptance(&u->counter_var);
    *accepted = u->accepted; *bead_moves = u->bead_moves;
}

/* ---- ReshapeLinear  reshape.jl:31-91 ---- returns 1 acc, 0 rejected, -1 bridge failed */
static int reshape_linear_core(ora_system *s, int64_t n, int64_t j0, int64_t m, const gsrc *g, double u, int commit,
                               double *w_i_out, double *w_u_out, double *rp_out)
{
    int dim = s->dim; int64_t M = s->M, jm = j0 + m; int rows = (int)m + 1;
    int64_t *cycle = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1));
    int64_t Ncyc = ora_subcycle(s, n, cycle);
#define FP(j) ora_pcycle((j), cycle, Ncyc, M)
    double *rp = (double *)calloc((size_t)rows * dim, sizeof(double));
    double *Vp = (double *)calloc((size_t)m, sizeof(double));
    for (int k = 0; k < dim; ++k) { rp[k * rows] = R_(s, n, j0, k + 1); rp[k * rows + m] = R_(s, FP(jm), mod1(jm, M), k + 1); }
    int ret = -1; double w_initial = 0.0, w_updated = 0.0;
    if (hardspherelevy(rp, rows, s, j0, cycle[0], g)) {
        for (int64_t j = j0; j <= jm - 1; ++j) {
            int64_t jp = j - j0 + 1;
            w_initial += V_(s, FP(j), mod1(j, M));
            double a[2], b[2];
            for (int k = 0; k < dim; ++k) { a[k] = rp[k * rows + jp - 1]; b[k] = rp[k * rows + jp]; }
            if (!(s->compat & ORA_COMPAT_PAIR_BYVALUE) && s->interactions) {
                w_initial += interaction_old(s, FP(j), mod1(j, M));
                int64_t exc = FP(j);
                w_updated += interaction_new(s, a, b, mod1(j0 + jp - 1, M), &exc, 1, 1);
            }
            Vp[jp - 1] = ora_lnV(a, b, dim, s->tau, &s->pot);
        }
        double sv = 0.0; for (int64_t i = 0; i < m; ++i) sv = (i == 0) ? Vp[0] : sv + Vp[i];
        w_updated += sv;
        ret = ora_metropolis(pimc_exp(w_updated - w_initial), u);
        if (ret && commit) {
            for (int64_t j = 1; j <= m; ++j) {
                int64_t p = FP(j0 + j - 1), sl = mod1(j0 + j - 1, M); double c[2];
                for (int k = 0; k < dim; ++k) { c[k] = rp[k * rows + j - 1]; R_(s, p, sl, k + 1) = c[k]; }
                V_(s, p, sl) = Vp[j - 1];
                update_nn_bead(s, p, sl, c);
            }
        }
    }
    if (w_i_out) *w_i_out = w_initial;
    if (w_u_out) *w_u_out = w_updated;
    if (rp_out) memcpy(rp_out, rp, sizeof(double) * rows * dim);
#undef FP
    free(rp); free(Vp); free(cycle);
    return ret;
}
int ora_reshape_linear_explicit(ora_system *s, int64_t n, int64_t j0, int64_t m, const double *xi, double u, int commit,
                                double *w_initial, double *w_updated, double *rprime)
{
    gsrc g; memset(&g, 0, sizeof g); g.xi = xi; g.dim = s->dim;
    return reshape_linear_core(s, n, j0, m, &g, u, commit, w_initial, w_updated, rprime);
}

/* ---- ReshapeSwapLinear  reshape.jl:123-283 ---- returns 1 acc, 0 rejected, -1 bridge failed, -2 n1==n2 */
static int reshape_swap_core(ora_system *s, int64_t n1, int64_t n2, int64_t j0, int64_t m, const gsrc *g1, const gsrc *g2,
                             double u, int commit, double *w_i_out, double *w_u_out)
{
    if (n1 == n2) return -2;
    int dim = s->dim; int64_t M = s->M, jm = j0 + m; int rows = (int)m + 1;
    int64_t *pol1 = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1)), *pol2 = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1));
    int64_t Np1 = ora_subcycle(s, n1, pol1), Np2 = ora_subcycle(s, n2, pol2);
#define FP1(j) ora_pcycle((j), pol1, Np1, M)
#define FP2(j) ora_pcycle((j), pol2, Np2, M)
    double *r1 = (double *)calloc((size_t)rows * dim, sizeof(double)), *r2 = (double *)calloc((size_t)rows * dim, sizeof(double));
    double *V1 = (double *)calloc((size_t)m, sizeof(double)), *V2 = (double *)calloc((size_t)m, sizeof(double));
    for (int k = 0; k < dim; ++k) {
        r1[k * rows] = R_(s, n1, j0, k + 1); r2[k * rows] = R_(s, n2, j0, k + 1);
        r1[k * rows + m] = R_(s, FP2(jm), mod1(jm, M), k + 1); r2[k * rows + m] = R_(s, FP1(jm), mod1(jm, M), k + 1);
    }
    int b1 = hardspherelevy(r1, rows, s, j0, pol2[0], g1);
    int b2 = hardspherelevy(r2, rows, s, j0, pol1[0], g2);
    int ret = -1; double w_initial = 0.0, w_updated = 0.0;
    if (b1 && b2) {
        int64_t *nl = (int64_t *)malloc(sizeof(int64_t) * max_cand(s));
        for (int64_t j = j0; j <= jm - 1; ++j) {
            int64_t sl = mod1(j, M), p1 = FP1(j), p2 = FP2(j);
            w_initial += V_(s, p1, sl) + V_(s, p2, sl);
            if (s->interactions) { w_initial += interaction_old(s, p1, sl); w_initial += interaction_old(s, p2, sl); }
        }
        for (int64_t j = j0; j <= jm - 1; ++j) {
            int64_t jp = j - j0 + 1, sl = mod1(j, M); double a1[2], b1v[2], a2[2], b2v[2];
            for (int k = 0; k < dim; ++k) { a1[k] = r1[k * rows + jp - 1]; b1v[k] = r1[k * rows + jp]; a2[k] = r2[k * rows + jp - 1]; b2v[k] = r2[k * rows + jp]; }
            V1[jp - 1] = ora_lnV(a1, b1v, dim, s->tau, &s->pot);
            V2[jp - 1] = ora_lnV(a2, b2v, dim, s->tau, &s->pot);
            if (s->interactions) {
                int64_t exc[2] = { FP1(j), FP2(j) };
                double add = interaction_new(s, a1, b1v, sl, exc, 2, 0);
                add += interaction_new(s, a2, b2v, sl, exc, 2, 0);
                if (s->compat & ORA_COMPAT_SWAP_SIGN) w_initial += add; else w_updated += add; /* reshape.jl:224,239 */
            }
        }
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = 0; i < m; ++i) { s1 = (i == 0) ? V1[0] : s1 + V1[i]; s2 = (i == 0) ? V2[0] : s2 + V2[i]; }
        w_updated += s1 + s2;
        free(nl);
        ret = ora_metropolis(pimc_exp(w_updated - w_initial), u);
        if (ret && commit) {
            rm_nn(s, pol1, Np1); rm_nn(s, pol2, Np2);
            int64_t t = s->next[n1 - 1]; s->next[n1 - 1] = s->next[n2 - 1]; s->next[n2 - 1] = t;
            Np1 = ora_subcycle(s, n1, pol1); Np2 = ora_subcycle(s, n2, pol2); /* closures see the new cycles */
            for (int64_t j = 2; j <= m + 1; ++j) {
                int64_t sl = mod1(j0 + j - 1, M), p1 = FP1(j0 + j - 1), p2 = FP2(j0 + j - 1); double c1[2], c2[2];
                for (int k = 0; k < dim; ++k) { c1[k] = r1[k * rows + j - 1]; c2[k] = r2[k * rows + j - 1]; }
                for (int k = 0; k < dim; ++k) R_(s, p1, sl, k + 1) = c1[k];
                B_(s, p1, sl) = ora_bin(c1, dim, s->nbins, s->L);
                for (int k = 0; k < dim; ++k) R_(s, p2, sl, k + 1) = c2[k];
                B_(s, p2, sl) = ora_bin(c2, dim, s->nbins, s->L);
            }
            for (int64_t j = 1; j <= m; ++j) {
                int64_t sl = mod1(j0 + j - 1, M);
                V_(s, FP1(j0 + j - 1), sl) = V1[j - 1];
                V_(s, FP2(j0 + j - 1), sl) = V2[j - 1];
            }
            if (jm < M) {
                for (int64_t j = jm + 1; j <= M; ++j) {
                    for (int k = 1; k <= dim; ++k) { double x = R_(s, n1, j, k); R_(s, n1, j, k) = R_(s, n2, j, k); R_(s, n2, j, k) = x; }
                    int64_t bb = B_(s, n1, j); B_(s, n1, j) = B_(s, n2, j); B_(s, n2, j) = bb;
                    double vv = V_(s, n1, j); V_(s, n1, j) = V_(s, n2, j); V_(s, n2, j) = vv;
                }
            }
            if (!(s->compat & ORA_COMPAT_SWAP_STALE_LINK) && jm <= M) { /* intended: the link leaving slice j_m changes owner too */
                double vv = V_(s, n1, jm); V_(s, n1, jm) = V_(s, n2, jm); V_(s, n2, jm) = vv;
            }
            add_nn(s, pol1, Np1); add_nn(s, pol2, Np2); /* merged cycles are added twice, as in the reference */
        }
    }
    if (w_i_out) *w_i_out = w_initial;
    if (w_u_out) *w_u_out = w_updated;
#undef FP1
#undef FP2
    free(r1); free(r2); free(V1); free(V2); free(pol1); free(pol2);
    return ret;
}
int ora_reshape_swap_explicit(ora_system *s, int64_t n1, int64_t n2, int64_t j0, int64_t m, const double *xi1, const double *xi2,
                              double u, int commit, double *w_initial, double *w_updated)
{
    gsrc g1, g2; memset(&g1, 0, sizeof g1); memset(&g2, 0, sizeof g2);
    g1.xi = xi1; g1.dim = s->dim; g2.xi = xi2; g2.dim = s->dim;
    return reshape_swap_core(s, n1, n2, j0, m, &g1, &g2, u, commit, w_initial, w_updated);
}

/* ---- centre-of-mass moves  com.jl:31-104 (worms = 0) and com.jl:136-224 ; move_polymer! helper.jl:368-395 ---- */
typedef struct { const double *d; pimc_stream st; uint32_t slot; } dsrc;
static int com_core(ora_system *s, const int64_t *pol, int64_t Npol, int polymer, double maxd, const dsrc *ds, double u, int commit,
                    double *w_i_out, double *w_u_out)
{
    int dim = s->dim; int64_t M = s->M;
    double w_initial = 0.0, w_updated = 0.0;
    for (int64_t i = 0; i < Npol; ++i) {
        double sv = 0.0; for (int64_t j = 1; j <= M; ++j) sv = (j == 1) ? V_(s, pol[i], 1) : sv + V_(s, pol[i], j);
        w_initial += sv;
        if (!(s->compat & ORA_COMPAT_PAIR_BYVALUE) && s->interactions)
            for (int64_t j = 1; j <= M; ++j) w_initial += interaction_old(s, pol[i], j);
    }
    double *rp = (double *)calloc((size_t)M * dim * Npol, sizeof(double)); /* [i][k][j] */
    double *Vp = (double *)calloc((size_t)M * Npol, sizeof(double));
#define RP(j, k, i) rp[(((i) * dim) + (k)) * M + ((j) - 1)]
    int ok = 0, pass = 1; int64_t ctr = 0;
    while (pass) {
        pass = 0; ctr += 1;
        if (ctr > s->ctr) break;
        double d[2];
        if (ds->d) { if (ctr > 1) break; for (int k = 0; k < dim; ++k) d[k] = ds->d[k]; }
        else {
            pimc_u4 w = pimc_draw(ds->st, ds->slot, PIMC_K_COM, (uint32_t)(ctr - 1), 0);
            d[0] = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5);
            d[1] = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        }
        for (int64_t i = 0; i < Npol; ++i) {
            int64_t n = pol[i];
            for (int64_t j = 1; j <= M; ++j) {
                double c[2];
                for (int k = 0; k < dim; ++k) { c[k] = ora_teleport(R_(s, n, j, k + 1) + d[k], s->L); RP(j, k, i) = c[k]; }
                if (s->a > 0.0) {
                    int64_t nn = ora_find_nn(s, c, j, &n, 1);
                    if (nn != -1) {
                        double dd[2]; for (int k = 0; k < dim; ++k) dd[k] = ora_distance(c[k], R_(s, nn, j, k + 1), s->L);
                        if (vnorm(dd, dim) < s->a) { pass = 1; break; }
                    }
                }
            }
        }
        if (!pass) ok = 1;
    }
    int ret = -1;
    if (ok) {
        for (int64_t i = 0; i < Npol; ++i) {
            int64_t n = pol[i];
            for (int64_t j = 1; j <= M; ++j) {
                int64_t nnext, mnext; next_bead(s, n, j, &nnext, &mnext);
                if (nnext == 0) continue;
                int64_t ii = (nnext == n) ? i : (polymer ? mod1(i + 2, Npol) - 1 : i + 1);
                if (ii >= Npol) ii = Npol - 1; /* unreachable for single particles (next == self) */
                double a[2], b[2]; for (int k = 0; k < dim; ++k) { a[k] = RP(j, k, i); b[k] = RP(mnext, k, ii); }
                Vp[i * M + j - 1] = ora_lnV(a, b, dim, s->tau, &s->pot);
                if (!(s->compat & ORA_COMPAT_PAIR_BYVALUE) && s->interactions) w_updated += interaction_new(s, a, b, j, &n, 1, 0);
            }
        }
        double sv = 0.0; for (int64_t q = 0; q < M * Npol; ++q) sv = (q == 0) ? Vp[0] : sv + Vp[q];
        w_updated += sv;
        ret = ora_metropolis(pimc_exp(w_updated - w_initial), u);
        if (ret && commit) {
            for (int64_t i = 0; i < Npol; ++i) {
                int64_t n = pol[i];
                for (int64_t j = 1; j <= M; ++j) {
                    double c[2];
                    V_(s, n, j) = Vp[i * M + j - 1];
                    for (int k = 0; k < dim; ++k) { c[k] = RP(j, k, i); R_(s, n, j, k + 1) = c[k]; }
                    update_nn_bead(s, n, j, c);
                }
            }
        }
    }
#undef RP
    if (w_i_out) *w_i_out = w_initial;
    if (w_u_out) *w_u_out = w_updated;
    free(rp); free(Vp);
    return ret;
}
int ora_com_explicit(ora_system *s, int64_t n, int polymer, const double *d, double u, int commit, double *w_initial, double *w_updated)
{
    int64_t *pol = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1));
    int64_t Npol = ora_subcycle(s, n, pol);
    dsrc ds; memset(&ds, 0, sizeof ds); ds.d = d;
    int r = com_core(s, pol, Npol, polymer, 0.0, &ds, u, commit, w_initial, w_updated);
    free(pol);
    return r;
}

/* One functor call + the bookkeeping of apply! (simulation.jl:19-27). forced_n/forced_j0: sweep schedule. */
static int functor_call(ora_system *s, ora_update *u, uint32_t slot, int64_t forced_n, int64_t forced_j0, int64_t m_snapshot, double size_snapshot)
{
    pimc_stream st = pimc_stream_make(s->seed, s->chain, s->iter);
    pimc_u4 dt = pimc_draw(st, slot, PIMC_K_TASK, 0, 0);
    pimc_u4 dm = pimc_draw(st, slot, PIMC_K_TASK, 0, 1);
    double umet = pimc_u01_co(dm.w[0], dm.w[1]);
    int acc = 0;
    switch (u->kind) {
    case ORA_UPD_RESHAPE_LINEAR: {
        if (s->N == 0) return 0;
        int64_t n = forced_n ? forced_n : 1 + pimc_index(dt.w[0], (uint32_t)s->N);
        int64_t j0 = forced_j0 ? forced_j0 : 1 + pimc_index(dt.w[1], (uint32_t)s->M);
        int64_t mm = 2 + pimc_index(dt.w[2], (uint32_t)(m_snapshot - 1));
        int64_t m = u->maxslices < mm ? u->maxslices : mm;
        gsrc g; memset(&g, 0, sizeof g); g.st = st; g.slot = slot; g.kind = PIMC_K_BRIDGE; g.dim = s->dim;
        int r = reshape_linear_core(s, n, j0, m, &g, umet, 1, NULL, NULL, NULL);
        acc = r == 1; u->bead_moves += m - 1;
        queue_push(&u->counter_var, acc);
        break; }
    case ORA_UPD_RESHAPE_SWAP: {
        if (s->N <= 1) return 0;
        int64_t j0 = forced_j0 ? forced_j0 : 1 + pimc_index(dt.w[1], (uint32_t)s->M);
        int64_t mm = 2 + pimc_index(dt.w[2], (uint32_t)(m_snapshot - 1));
        int64_t m = u->maxslices < mm ? u->maxslices : mm;
        pimc_u4 dsw = pimc_draw(st, slot, PIMC_K_SWAP, 0, 0);
        int64_t n1 = 1 + pimc_index(dsw.w[0], (uint32_t)s->N);
        double *w = (double *)malloc(sizeof(double) * s->N);
        ora_swap_weights(s, n1, j0, m, w);
        double norm = 0.0; for (int64_t i = 0; i < s->N; ++i) norm = (i == 0) ? w[0] : norm + w[i];
        for (int64_t i = 0; i < s->N; ++i) w[i] = w[i] / norm;
        int64_t n2 = sample_weighted(w, s->N, pimc_u01_co(dsw.w[2], dsw.w[3]));
        free(w);
        if (n1 == n2) return 0; /* reshape.jl:134-136 : no queue! */
        gsrc g1, g2; memset(&g1, 0, sizeof g1); memset(&g2, 0, sizeof g2);
        g1.st = st; g1.slot = slot; g1.kind = PIMC_K_BRIDGE; g1.dim = s->dim; g2 = g1; g2.kind = PIMC_K_BRIDGE2;
        int r = reshape_swap_core(s, n1, n2, j0, m, &g1, &g2, umet, 1, NULL, NULL);
        acc = r == 1; u->bead_moves += 2 * (m - 1);
        queue_push(&u->counter_var, acc);
        break; }
    case ORA_UPD_SINGLE_COM: {
        if (s->N == 0) return 0;
        int64_t n1 = forced_n;
        if (!n1) {
            int64_t *sp = (int64_t *)malloc(sizeof(int64_t) * s->N), cnt = 0;
            for (int64_t i = 1; i <= s->N; ++i) if (s->next[i - 1] == i) sp[cnt++] = i;
            if (cnt == 0) { free(sp); return 0; } /* com.jl:158-160 : no queue! */
            n1 = sp[pimc_index(dt.w[0], (uint32_t)cnt)];
            free(sp);
        }
        dsrc ds; memset(&ds, 0, sizeof ds); ds.st = st; ds.slot = slot;
        int64_t pol[1] = { n1 };
        int r = com_core(s, pol, 1, 0, size_snapshot, &ds, umet, 1, NULL, NULL);
        acc = r == 1; u->bead_moves += s->M;
        queue_push(&u->counter_var, acc);
        break; }
    case ORA_UPD_POLYMER_COM: {
        if (s->N == 0) return 0;
        int64_t np = forced_n ? forced_n : 1 + pimc_index(dt.w[0], (uint32_t)s->N);
        int64_t *pol = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1));
        int64_t Npol = ora_subcycle(s, np, pol);
        dsrc ds; memset(&ds, 0, sizeof ds); ds.st = st; ds.slot = slot;
        int r = com_core(s, pol, Npol, 1, size_snapshot, &ds, umet, 1, NULL, NULL);
        acc = r == 1; u->bead_moves += s->M * Npol;
        queue_push(&u->counter_var, acc);
        free(pol);
        break; }
    }
    return acc;
}
/* simulation.jl:19-27 */
static void apply_bookkeeping(ora_update *u, int acc)
{
    queue_push(&u->counter, acc);
    u->accepted += acc;
    if (u->counter_var.tries % u->counter_var.adj == 0) {
        double a = acceptance(&u->counter_var);
        if (u->kind == ORA_UPD_RESHAPE_LINEAR || u->kind == ORA_UPD_RESHAPE_SWAP)
            u->m = ora_adjust_slices(u->m, u->minslices, u->maxslices, u->minacc, u->maxacc, a);
        else
            u->size = ora_adjust_step(u->size, u->minstep, u->maxstep, u->minacc, u->maxacc, a);
    }
}
int ora_update_call(ora_system *s, ora_update *u, uint32_t slot, int64_t forced_n, int64_t forced_j0)
{
    int acc = functor_call(s, u, slot, forced_n, forced_j0, u->m, u->size);
    apply_bookkeeping(u, acc);
    return acc;
}

/* ================= src/measurement.jl ================= */
struct ora_energy { double *E, *Ev; int64_t cap, n; };
ora_energy *ora_energy_create(int64_t cap)
{
    ora_energy *e = (ora_energy *)calloc(1, sizeof *e);
    e->cap = cap; e->E = (double *)calloc((size_t)cap, sizeof(double)); e->Ev = (double *)calloc((size_t)cap, sizeof(double));
    return e;
}
void ora_energy_destroy(ora_energy *e) { if (!e) return; free(e->E); free(e->Ev); free(e); }
int64_t ora_energy_read(const ora_energy *e, double *E, double *Ev, int64_t cap)
{
    int64_t n = e->n < cap ? e->n : cap;
    memcpy(E, e->E, sizeof(double) * n); memcpy(Ev, e->Ev, sizeof(double) * n);
    return e->n;
}
/* measurement.jl:92-122 */
void ora_energy_now(const ora_system *s, double *E, double *Ev, double *parts)
{
    double link = 0, pot = 0, vkin = 0, vpot = 0; int dim = s->dim;
    for (int64_t i = 1; i <= s->N; ++i)
        for (int64_t j = 1; j <= s->M; ++j) {
            int64_t inext = (j == s->M) ? s->next[i - 1] : i, jnext = mod1(j + 1, s->M);
            double a[2], b[2], dv[2], d2 = 0.0, rdv = 0.0;
            for (int k = 0; k < dim; ++k) { a[k] = R_(s, i, j, k + 1); b[k] = R_(s, inext, jnext, k + 1); }
            for (int k = 0; k < dim; ++k) { double dr = ora_distance(a[k], b[k], s->L); d2 = (k == 0) ? dr * dr : d2 + dr * dr; }
            link += d2;
            double va = ora_potential_eval(&s->pot, a, dim), vb = ora_potential_eval(&s->pot, b, dim);
            pot += va + vb;
            ora_potential_grad(&s->pot, a, dim, dv);
            for (int k = 0; k < dim; ++k) rdv = (k == 0) ? a[0] * dv[0] : rdv + a[k] * dv[k];
            vkin += rdv;
            vpot += va + vb;
        }
    *E = (double)(s->dim * s->N) / (2 * s->tau) - 1 / (4 * s->lambda * (s->tau * s->tau) * s->M) * link + 1.0 / (2 * s->M) * pot;
    *Ev = 1.0 / (2 * s->M) * vkin + 1.0 / (2 * s->M) * vpot;
    if (parts) { parts[0] = link; parts[1] = pot; parts[2] = vkin; }
}
/* ---- estimators the reference lists as TODO (measurement.jl:125-127), defined here in the style of its functors ---- */
/* Radial distribution (`#TODO radial distribution`): histogram of equal-time pair distances.  Per measurement, for every time slice m and every
 * pair i < j: d = |distance.(r_i[m, :], r_j[m, :], L)| (minimum image, propagator.jl:6-9), ib = floor(d / bin), bin = rmax / nbins;
 * hist[ib] += 1 when ib < nbins; ndata += M (the convention of Density, measurement.jl:54).  g(r) is the histogram divided by the ideal-gas
 * count ndata * N(N-1)/2 * shell(r) / vol at read-out. */
struct ora_paircorr { int64_t nbins; double rmax, bin; double *hist; int64_t ndata; };
ora_paircorr *ora_paircorr_create(const ora_system *s, int64_t nbins, double rmax)
{
    (void)s;
    ora_paircorr *g = (ora_paircorr *)calloc(1, sizeof *g);
    g->nbins = nbins; g->rmax = rmax; g->bin = rmax / (double)nbins;
    g->hist = (double *)calloc((size_t)nbins, sizeof(double));
    return g;
}
void ora_paircorr_destroy(ora_paircorr *g) { if (!g) return; free(g->hist); free(g); }
void ora_paircorr_measure(ora_paircorr *g, const ora_system *s)
{
    for (int64_t m = 1; m <= s->M; ++m)
        for (int64_t i = 1; i <= s->N; ++i)
            for (int64_t j = i + 1; j <= s->N; ++j) {
                double d2 = 0.0;
                for (int k = 1; k <= s->dim; ++k) { double dr = ora_distance(R_(s, i, m, k), R_(s, j, m, k), s->L); d2 = (k == 1) ? dr * dr : d2 + dr * dr; }
                const double ib = floor(sqrt(d2) / g->bin);
                if (ib < (double)g->nbins) g->hist[(int64_t)ib] += 1.0;
            }
    g->ndata += s->M;
}
int64_t ora_paircorr_read(const ora_paircorr *g, double *hist, double *bin)
{
    if (hist) memcpy(hist, g->hist, sizeof(double) * (size_t)g->nbins);
    if (bin) *bin = g->bin;
    return g->ndata;
}
/* Winding number (`#TODO Superfluid Fraction`): W_k = (1 / 2L) * sum over all links of the minimum-image displacement
 * teleport(r_next[k] - r[k], L) (propagator.jl:30-32 wraps into [-L, L)); an integer for closed paths.  The superfluid fraction follows at
 * read-out from <W^2> (2L)^2 / (2 dim lambda beta N) (Pollock & Ceperley 1987). */
void ora_winding_now(const ora_system *s, double *W)
{
    for (int k = 1; k <= s->dim; ++k) {
        double acc = 0.0;
        for (int64_t i = 1; i <= s->N; ++i)
            for (int64_t j = 1; j <= s->M; ++j) {
                int64_t inext = (j == s->M) ? s->next[i - 1] : i, jnext = mod1(j + 1, s->M);
                acc += ora_teleport(R_(s, inext, jnext, k) - R_(s, i, j, k), s->L);
            }
        W[k - 1] = acc / (2 * s->L);
    }
}
/* Static structure factor (`#TODO Compressibilty`, measurement.jl:127): for the wave vectors of the periodic box k = (pi / L) (a, b),
 * a = 0..kmax, b = -kmax..kmax (the half plane a > 0, or a = 0 and b > 0; 1-D: b = 0, a = 1..kmax), and every time slice m:
 * rho_k(m) = sum_n exp(i k . r_n[m, :]);  out[a * (2 kmax + 1) + (b + kmax)] = sum_m |rho_k(m)|^2 (entries outside the half plane stay 0).
 * S(k) = <|rho_k|^2> / N at read-out (ndata += M per call, the convention of Density, measurement.jl:54); the isothermal compressibility follows
 * from the long-wavelength limit S(k -> 0) = rho k_B T kappa_T, estimated on the smallest shell |k| = pi / L. */
void ora_structure_now(const ora_system *s, int kmax, double *out)
{
    const int nb = 2 * kmax + 1;
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < (kmax + 1) * nb; ++i) out[i] = 0.0;
    for (int a = 0; a <= kmax; ++a)
        for (int b = -kmax; b <= kmax; ++b) {
            if (s->dim == 1 ? (b != 0 || a == 0) : !(a > 0 || b > 0)) continue;
            double acc = 0.0;
            for (int64_t m = 1; m <= s->M; ++m) {
                double re = 0.0, im = 0.0;
                for (int64_t n = 1; n <= s->N; ++n) {
                    const double ph = PI * ((a * R_(s, n, m, 1) + (s->dim > 1 ? b * R_(s, n, m, 2) : 0.0)) / s->L);
                    re += cos(ph); im += sin(ph);
                }
                acc += re * re + im * im;
            }
            out[a * nb + (b + kmax)] = acc;
        }
}
struct ora_density { int64_t nbins; int dim; double bin; double *dens; int64_t ndata; };
/* measurement.jl:31-38 */
ora_density *ora_density_create(const ora_system *s, int64_t nbins)
{
    ora_density *d = (ora_density *)calloc(1, sizeof *d);
    d->nbins = nbins; d->dim = s->dim; d->bin = (2 * s->L) / nbins;
    d->dens = (double *)calloc((size_t)(s->dim == 2 ? nbins * nbins : nbins), sizeof(double));
    return d;
}
void ora_density_destroy(ora_density *d) { if (!d) return; free(d->dens); free(d); }
/* measurement.jl:45-55 ; dens is column-major dens[ix + nbins*iy] like the Julia array */
void ora_density_measure(ora_density *d, const ora_system *s)
{
    int shift = (s->compat & ORA_COMPAT_DENSITY_SHIFT) != 0;
    for (int64_t n = 1; n <= s->N; ++n)
        for (int64_t m = 1; m <= s->M; ++m) {
            int64_t ib[2] = { 1, 1 }; int ok = 1;
            for (int k = 0; k < s->dim; ++k) {
                ib[k] = (int64_t)floor((R_(s, n, m, k + 1) + s->L) / d->bin);
                if (shift) { if (!(ib[k] > 0 && ib[k] < d->nbins + 1)) ok = 0; }
                else { if (!(ib[k] >= 0 && ib[k] < d->nbins)) ok = 0; ib[k] += 1; }
            }
            if (ok) d->dens[(ib[0] - 1) + (s->dim == 2 ? d->nbins * (ib[1] - 1) : 0)] += 1;
        }
    d->ndata += s->M;
}
int64_t ora_density_read(const ora_density *d, double *dens, double *bin)
{
    int64_t sz = d->dim == 2 ? d->nbins * d->nbins : d->nbins;
    if (dens) memcpy(dens, d->dens, sizeof(double) * sz);
    if (bin) *bin = d->bin;
    return d->ndata;
}

/* ================= src/system.jl ================= */
/* system.jl:36-78 */
static int init_world(ora_system *s)
{
    int dim = s->dim; int64_t M = s->M; pimc_stream st = pimc_stream_make(s->seed, s->chain, 0);
    double *r = (double *)malloc(sizeof(double) * M * dim), *xi = (double *)malloc(sizeof(double) * M * dim);
    for (int64_t n = 1; n <= s->N; ++n) {
        uint32_t slot = (uint32_t)(n - 1); int64_t ctr = 0, levy_calls = 0; int pass = (n == 1) ? 0 : 1, first = 1;
#define START(attempt) do { pimc_u4 w = pimc_draw(st, slot, PIMC_K_INIT0, (uint32_t)(attempt), 0); \
            double uu[2] = { pimc_u01_co(w.w[0], w.w[1]), pimc_u01_co(w.w[2], w.w[3]) }; \
            for (int k = 0; k < dim; ++k) { r[k * M] = 2 * s->L * (uu[k] - 0.5); r[k * M + M - 1] = r[k * M]; } } while (0)
#define LEVY() do { for (int64_t t = 1; t <= M - 2; ++t) { double g0, g1; \
                pimc_gauss_pair(pimc_draw(st, slot, PIMC_K_INIT, (uint32_t)levy_calls, (uint32_t)t), &g0, &g1); \
                xi[(t - 1) * dim] = g0; if (dim > 1) xi[(t - 1) * dim + 1] = g1; } \
            ora_levy(r, (int)M, dim, s->tau, s->L, s->lambda, xi); levy_calls += 1; } while (0)
        while (pass) {
            pass = 0; ctr += 1;
            if (ctr > 10000) { FAIL("creating world of %d particles with hardspheres of length %g failed", s->N, s->a); free(r); free(xi); return -1; }
            START(ctr - 1); LEVY();
            for (int64_t m = 1; m <= M; ++m)
                for (int64_t i = 1; i <= n - 1; ++i) {
                    double dd[2]; for (int k = 0; k < dim; ++k) dd[k] = ora_distance(R_(s, i, m, k + 1), r[k * M + m - 1], s->L);
                    if (vnorm(dd, dim) < s->a) { LEVY(); pass = 1; }
                }
            first = 0;
        }
        if (n == 1) { START(0); LEVY(); }
        (void)first;
        for (int64_t m = 1; m <= M; ++m) for (int k = 0; k < dim; ++k) R_(s, n, m, k + 1) = r[k * M + m - 1];
        for (int64_t m = 1; m <= M; ++m) {
            double a[2], b[2]; int64_t mn = mod1(m + 1, M);
            for (int k = 0; k < dim; ++k) { a[k] = r[k * M + m - 1]; b[k] = r[k * M + mn - 1]; }
            V_(s, n, m) = ora_lnV(a, b, dim, s->tau, &s->pot);
        }
        s->next[n - 1] = n;
#undef START
#undef LEVY
    }
    free(r); free(xi);
    return 0;
}
static void relink_cache(ora_system *s) /* system.jl:72-74 generalised to permuted worlds (link M -> next particle) */
{
    for (int64_t n = 1; n <= s->N; ++n)
        for (int64_t m = 1; m <= s->M; ++m) {
            int64_t nn, mn; next_bead(s, n, m, &nn, &mn);
            double a[2], b[2]; for (int k = 0; k < s->dim; ++k) { a[k] = R_(s, n, m, k + 1); b[k] = R_(s, nn, mn, k + 1); }
            V_(s, n, m) = ora_lnV(a, b, s->dim, s->tau, &s->pot);
        }
}
/* system.jl:129-167 (constructor), :17-34 (init_int), :80-91 (init_nn) */
ora_system *ora_create(const ora_config *c)
{
    if (c->dim < 1 || c->dim > 2 || c->M < 3 || c->N < 0) { FAIL("bad config"); return NULL; }
    ora_system *s = (ora_system *)calloc(1, sizeof *s);
    s->dim = c->dim; s->M = c->M; s->N = c->N; s->Ninit = c->N; s->mu = c->mu; s->lambda = c->lambda; s->L = c->L;
    s->beta = 1.0 / c->T; s->tau = s->beta / c->M; s->vol = pow(2 * c->L, c->dim);
    s->a = c->interactions ? exp(-2 * M_PI / c->g) : 0.0;
    s->pot = c->pot; s->interactions = c->interactions; s->compat = c->compat; s->seed = c->seed; s->chain = c->chain; s->iter = 0;
    s->ctr = 10000; s->Ncycle = c->Ncycle; s->N_MC = 0; s->Nctr = 0;
    size_t nb = (size_t)(c->N > 0 ? c->N : 1) * c->M;
    s->r = (double *)calloc(nb * c->dim, sizeof(double)); s->V = (double *)calloc(nb, sizeof(double));
    s->bins = (int64_t *)calloc(nb, sizeof(int64_t)); s->next = (int64_t *)calloc((size_t)c->N + 1, sizeof(int64_t));
    for (int64_t n = 1; n <= s->N; ++n) s->next[n - 1] = n;
    if (c->tab && c->tab_n > 1) {
        s->tab_n = c->tab_n; s->tab_lo = c->tab_lo; s->tab_hi = c->tab_hi;
        s->tab = (double *)malloc(sizeof(double) * c->tab_n * c->tab_n); memcpy(s->tab, c->tab, sizeof(double) * c->tab_n * c->tab_n);
    }
    if (c->init) { if (init_world(s) != 0) { ora_destroy(s); return NULL; } }
    s->r_a = c->r_a;
    if (s->r_a == 0.0) {
        if (!c->interactions) s->r_a = c->L / 4;
        else { FAIL("interactions with r_a == 0 need determine_nnrange (Optim/Roots): out of scope, pass r_a"); ora_destroy(s); return NULL; }
    }
    s->nbins = (int64_t)floor((2 * c->L) / s->r_a);
    if (s->nbins < 1) s->nbins = 1;
    s->nst = c->dim == 2 ? 9 : 3;
    s->ncell = c->dim == 2 ? s->nbins * s->nbins : s->nbins;
    s->nn = (ivec *)calloc((size_t)s->M * s->ncell, sizeof(ivec));
    s->nbs = (int64_t *)malloc(sizeof(int64_t) * s->ncell * s->nst);
    for (int64_t b = 1; b <= s->ncell; ++b) ora_bin_neighbors(b, s->nbins, s->dim, s->nbs + (b - 1) * s->nst);
    ora_update_nnbins(s);
    return s;
}
void ora_destroy(ora_system *s)
{
    if (!s) return;
    if (s->nn) { for (int64_t i = 0; i < s->M * s->ncell; ++i) free(s->nn[i].v); free(s->nn); }
    free(s->nbs); free(s->r); free(s->V); free(s->bins); free(s->next); free(s->tab); free(s);
}
void ora_get_paths(const ora_system *s, double *r, double *V, int64_t *bins, int64_t *next)
{
    size_t nb = (size_t)s->N * s->M;
    if (r) memcpy(r, s->r, sizeof(double) * nb * s->dim);
    if (V) memcpy(V, s->V, sizeof(double) * nb);
    if (bins) memcpy(bins, s->bins, sizeof(int64_t) * nb);
    if (next) memcpy(next, s->next, sizeof(int64_t) * s->N);
}
void ora_set_paths(ora_system *s, const double *r, const int64_t *next)
{
    size_t nb = (size_t)s->N * s->M;
    if (r) memcpy(s->r, r, sizeof(double) * nb * s->dim);
    if (next) memcpy(s->next, next, sizeof(int64_t) * s->N);
    relink_cache(s);
    ora_update_nnbins(s);
}
void ora_get_scalars(const ora_system *s, double *out, int64_t *iout)
{
    out[0] = s->beta; out[1] = s->tau; out[2] = s->vol; out[3] = s->a; out[4] = s->r_a;
    iout[0] = s->nbins; iout[1] = s->N_MC; iout[2] = s->Nctr; iout[3] = s->ctr; iout[4] = (int64_t)s->iter;
}
void ora_set_iter(ora_system *s, uint64_t iter) { s->iter = iter; }
void ora_set_ctr(ora_system *s, int64_t ctr) { s->ctr = ctr; }
double ora_action_links(const ora_system *s)
{
    double w = 0.0;
    for (int64_t n = 1; n <= s->N; ++n) for (int64_t m = 1; m <= s->M; ++m) w += V_(s, n, m);
    return w;
}
double ora_action_links_recomputed(const ora_system *s)
{
    double w = 0.0;
    for (int64_t n = 1; n <= s->N; ++n)
        for (int64_t m = 1; m <= s->M; ++m) {
            int64_t nn, mn; next_bead(s, n, m, &nn, &mn);
            double a[2], b[2]; for (int k = 0; k < s->dim; ++k) { a[k] = R_(s, n, m, k + 1); b[k] = R_(s, nn, mn, k + 1); }
            w += ora_lnV(a, b, s->dim, s->tau, &s->pot);
        }
    return w;
}
/* intended pair action: every bead's interaction_action! (helper.jl:306-324) summed; each pair counted twice like the moves see it */
double ora_action_pairs(const ora_system *s)
{
    double w = 0.0;
    for (int64_t n = 1; n <= s->N; ++n) for (int64_t m = 1; m <= s->M; ++m) w += interaction_old(s, n, m);
    return w;
}

/* ================= src/simulation.jl ================= */
/* measurement.jl:1-17 */
static void measurement_Z_sector(ora_system *s, ora_energy **en, int nen, ora_density **de, int nde)
{
    if (nen + nde == 0) return;
    s->Nctr += 1;
    if (s->Nctr == s->Ncycle) {
        s->N_MC += 1;
        for (int i = 0; i < nen; ++i) {
            double E, Ev; ora_energy_now(s, &E, &Ev, NULL);
            if (en[i]->n < en[i]->cap) { en[i]->E[en[i]->n] = E; en[i]->Ev[en[i]->n] = Ev; }
            en[i]->n += 1; /* the reference errors on overflow (findfirst -> nothing) */
        }
        for (int i = 0; i < nde; ++i) ora_density_measure(de[i], s);
        s->Nctr = 0;
    }
}
/* simulation.jl:29-42 ; sched = SWEEP is the engine's batched schedule, executed here sequentially */
int ora_run(ora_system *s, int64_t n, ora_update **upd, const int64_t *every, int nupd, ora_energy **en, int nen, ora_density **de, int nde, int sched)
{
    s->ctr = (nen + nde == 0) ? 10000 : 1000;
    if (nupd <= 0) return -1;
    double *w = (double *)malloc(sizeof(double) * nupd);
    for (int i = 0; i < nupd; ++i) w[i] = 1.0 / (double)every[i];
    int64_t *cyc = (int64_t *)malloc(sizeof(int64_t) * (s->N + 1));
    uint8_t *accs = (uint8_t *)malloc((size_t)s->N + 1);
    if (sched == ORA_SCHED_SWEEP && (s->a > 0.0 || (s->interactions && !(s->compat & ORA_COMPAT_PAIR_BYVALUE)))) {
        FAIL("sweep schedule needs independent worldlines (a == 0, no pair action in the moves)"); free(w); free(cyc); free(accs); return -2;
    }
    for (int64_t it = 0; it < n; ++it) {
        pimc_stream st = pimc_stream_make(s->seed, s->chain, s->iter);
        pimc_u4 di = pimc_draw(st, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
        int64_t pick = sample_weighted(w, nupd, pimc_u01_co(di.w[0], di.w[1]));
        ora_update *u = upd[pick - 1];
        if ((sched != ORA_SCHED_SWEEP && sched != ORA_SCHED_SWEEP_SEQ) || u->kind == ORA_UPD_RESHAPE_SWAP) {
            int acc = functor_call(s, u, 0, 0, 0, u->m, u->size);
            apply_bookkeeping(u, acc);
        } else {
            int64_t j0w = 1 + pimc_index(di.w[2], (uint32_t)s->M), msnap = u->m; double ssnap = u->size; int64_t cnt = 0;
            int64_t tries0 = u->counter_var.tries;
            for (int64_t p = 1; p <= s->N; ++p) {
                int run_it = 1;
                if (u->kind == ORA_UPD_SINGLE_COM) run_it = (s->next[p - 1] == p);
                if (u->kind == ORA_UPD_POLYMER_COM) {
                    int64_t nc = ora_subcycle(s, p, cyc);
                    for (int64_t q = 0; q < nc; ++q) if (cyc[q] < p) run_it = 0;
                }
                if (!run_it) continue;
                accs[cnt++] = (uint8_t)functor_call(s, u, (uint32_t)(p - 1), p, u->kind == ORA_UPD_RESHAPE_LINEAR ? j0w : 0, msnap, ssnap);
            }
            /* bookkeeping replayed in slot order; counter_var was queued inside the functor calls.  The step/slice
             * variable is frozen during a sweep and adjusted once after it whenever counter_var.tries crossed a
             * multiple of adj (for N = 1 this is exactly apply!'s `tries % adj == 0`) -- DESIGN.md "sweep schedule" */
            for (int64_t q = 0; q < cnt; ++q) { queue_push(&u->counter, accs[q]); u->accepted += accs[q]; }
            if (cnt > 0 && (u->counter_var.tries / u->counter_var.adj) != (tries0 / u->counter_var.adj)) {
                double a = acceptance(&u->counter_var);
                if (u->kind == ORA_UPD_RESHAPE_LINEAR) u->m = ora_adjust_slices(u->m, u->minslices, u->maxslices, u->minacc, u->maxacc, a);
                else u->size = ora_adjust_step(u->size, u->minstep, u->maxstep, u->minacc, u->maxacc, a);
            }
        }
        measurement_Z_sector(s, en, nen, de, nde);
        s->iter += 1;
    }
    free(w); free(cyc); free(accs);
    return 0;
}
