// pimc_device.cuh -- device-side state descriptors and primitives of the sm_100a PIMC engine.
// All arithmetic is fp64 with FMA contraction disabled (-fmad=false) so that bridges are bit-identical to the
// reference's Julia arithmetic (no auto-FMA).  Reference citations are relative to the reference tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pimc_b200.h"
#include "../../include/pimc_rng.h"

#define PIMC_MAXU 8
#define PIMC_MAXE 4
#define PIMC_MAXD 4

struct PotDev {
    int kind, dv_kind;
    double k, depth, scale, sgn;
    int nang, helical;
    double ang[PIMC_MAX_ANGLES], sn[PIMC_MAX_ANGLES], cs[PIMC_MAX_ANGLES]; // sin/cos of the beam angles, host-evaluated
    // commensurate, inversion-symmetric beam sets (l25 = the 3-4-5 angles, l65, cubic: all the sets of examples/tools/potentialtools.jl): every
    // x-projection sin(a_i) is an integer multiple of one base value, every y-projection cos(a_i) of another, and the beams come in (a, -a, pi - a)
    // families, so the sine sum vanishes identically and the cosine sum is  C = sum_{a,b} wcc[a][b] cos(a tx) cos(b ty):  TWO sincos, two Chebyshev
    // recurrences and one small static double sum instead of one sincos per beam (host-detected in pot_to_dev; fast = 0 keeps the general form)
    int fast, nmx, nmy;
    double bx, by;                      // base projection * scale
    double wcc[(8 + 1) * (8 + 1)];      // [a][b], a <= nmx, b <= nmy
};
#define PIMC_FAST_NMAX 8

// HBM layout (SoA, fp64): r[C][N][dim][M]  -- the reference's per-particle M x dim column-major matrices, concatenated;
// Vl[C][N][M] cached link action; next[C][N] (0-based).  Scratch: prop[C][N][dim][M], propV[C][N][M], wtab[C][N].
struct DevSys {
    int dim, M, N, C;
    uint32_t chain_offset;
    double lambda, tau, L, mu, a, beta;
    int interactions, compat;
    long long ctr;
    unsigned long long seed;
    double *r, *Vl;
    int *next;
    double *prop, *propV, *wtab;
    const double *logtab;              // [2*128] table of pimc_log_tab (include/pimc_rng.h)
    const double *tab_alpha, *tab_sig; // [M+1] staging tables indexed by k = rows left: alpha_k = (k-1)/k, sigma_k = sqrt(((2 lambda) alpha_k) tau)  (helper.jl:131-134)
    // cell list (replaces src/nearest_neighbours.jl): per (chain, slice) singly linked lists
    int need_cells, nbins, ncell;
    double cellw;
    int *bins;      // [C][N][M]  0-based cell of every bead
    int *cell_head; // [C][ncell][M]  cell-major, slice fastest: consecutive beads of a worldline sit in the same or a neighbouring cell, so lanes
    int *cell_next; // [C][N][M]      that stride the slices of one worldline read contiguous heads / successors (HIDX / NIDX below)
    unsigned char *mult; // [C][N][M] multiplicity of the bead in its cell list (2 after a cycle-merging swap: add_nn! twice)
    const double *tab; int tab_n; double tab_lo, tab_hi;
    PotDev pot;
};

struct UpdDev {
    int kind;
    double vmin, vmax, minacc, maxacc;
    long long adj, range;
    int ring_words;
    double *var;             // [C] step size or number of slices
    long long *tries, *accepted, *tries_var, *bead_moves; // [C]
    int *ring_head, *ring_len, *ring_sum;                 // [C]
    unsigned *ring;                                       // [C][ring_words]
};
struct EnDev { double *E, *Ev; long long cap; double *acc; /* [C][5] n, sumE, sumE2, sumEv, sumEv2 */ };
struct DeDev { unsigned long long *dens; long long nbins; double bin; };
struct DevTables { UpdDev upd[PIMC_MAXU]; EnDev en[PIMC_MAXE]; DeDev de[PIMC_MAXD]; };
// estimators the reference lists as TODO (measurement.jl:125-127); passed to their kernels by value
#define PIMC_MAXP 4
#define PIMC_MAXW 4
struct PcDev { unsigned long long *hist; long long nbins; double rmax, bin; };   // g(r): counts per radial bin, all chains of the handle
struct WiDev { double *W; long long cap; };                                      // winding series W[cap][dim][C]
#define PIMC_MAXS 4
#define PIMC_SK_KMAX 6   // wave vectors (pi / L)(a, b) with a <= kmax, |b| <= kmax; one warp of k_structure per value of a
struct SkDev { double *S; int kmax; };                                           // structure-factor sums S[C][kmax + 1][2 kmax + 1] of |rho_k|^2 (per chain: deterministic)

struct RunParams {
    long long n;
    unsigned long long iter0;
    int nupd; int upd_id[PIMC_MAXU]; double w[PIMC_MAXU];
    int nen; int en_id[PIMC_MAXE]; long long en_k0[PIMC_MAXE];   // en_k0: samples the Energy object already holds (its own count)
    int nde; int de_id[PIMC_MAXD];
    int sched;
    long long Nctr0, N_MC0; int Ncycle;
    unsigned long long *stats; // [4] proposals, accepted, bead_moves, (unused)
    int fimpl;                 // FAITHFUL proposals: 0 warp-cooperative (pimc_faithful.cuh), 1 one thread (pimc_moves.cuh)
    unsigned long long *prof;  // [10] per-phase cycle counters (PIMC_PROF=1), else null
    double *fscr;              // HBM scratch of the warp-cooperative proposal when it does not fit shared memory, else null
};

#define RIDX(S, c, n, k, j) ((((size_t)(c) * (S).N + (n)) * (S).dim + (k)) * (S).M + (j))
#define VIDX(S, c, n, j)    (((size_t)(c) * (S).N + (n)) * (S).M + (j))
#define HIDX(S, c, j, b)    (((size_t)(c) * (S).ncell + (b)) * (S).M + (j))   /* head of the list of cell b at slice j */
#define NIDX(S, c, j, n)    (((size_t)(c) * (S).N + (n)) * (S).M + (j))       /* successor of particle n in its list at slice j */

// ---------------- src/propagator.jl ----------------
__device__ __forceinline__ double d_distance(double x1, double x2, double L) // propagator.jl:6-9
{
    double dx = fabs(x1 - x2);
    double alt = (2 * L) - dx;
    return alt < dx ? alt : dx;
}
__device__ __forceinline__ double d_teleport(double x, double L) // propagator.jl:30-32
{
    return ((x + L) - floor(x / (2 * L) + 0.5) * (2 * L)) - L;
}
// the same value without the IEEE division on the fast path: q = x * (1 / 2L) differs from x / 2L by <= 1 ulp, so floor(q + 0.5) can
// differ only when q + 0.5 sits within a few ulp of an integer; that case (and NaN / huge arguments) takes the exact path
__device__ __forceinline__ double d_teleport_q(double x, double L)
{
    const double twoL = 2 * L, inv2L = 1.0 / twoL;
    double s = x * inv2L + 0.5;
    double f = floor(s);
    double d = s - f;
    if (!(fabs(d - 0.5) <= 0.5 - 1e-9) || !(fabs(s) < 1e6)) f = floor(x / twoL + 0.5);
    return ((x + L) - f * twoL) - L;
}
__device__ __forceinline__ double d_sign(double x) { return (double)((x > 0) - (x < 0)); }

__device__ __forceinline__ double d_pot(const PotDev &p, double x, double y, int dim)
{
    switch (p.kind) {
    case PIMC_POT_HARMONIC: { double s = x * x; if (dim > 1) s = s + y * y; return (0.5 * p.k) * s; }
    case PIMC_POT_SIN2_1D: { double sn = sin(6.283185307179586 * x * p.scale); return p.depth * (sn * sn); }
    case PIMC_POT_LATTICE: {
        // normalized_intensity (examples/tools/potentialtools.jl:1-16).  The phase is 2 pi t with t = r * scale: t is reduced to
        // [0, 1) exactly and handed to the 2 pi u kernel of the Gaussian transform (1 ulp) -- half the instructions of a general
        // sincos and no loss from rounding 2 pi t first.  Parity with the reference / oracle (libm on 2 pi t): <= 1e-13 absolute.
        double s = 0.0, c = 0.0;
        if (dim < 2) y = 0.0;
        if (p.fast) {
            const double tx = x * p.bx, ty = y * p.by;
            double s1, c1, s2, c2;
            pimc_sincos2pi(tx - floor(tx), &s1, &c1);
            pimc_sincos2pi(ty - floor(ty), &s2, &c2);
            double CX[PIMC_FAST_NMAX + 1], CY[PIMC_FAST_NMAX + 1];          // static indices only: registers
            CX[0] = 1.0; CX[1] = c1; CY[0] = 1.0; CY[1] = c2;
            const double tcx = 2 * c1, tcy = 2 * c2;
#pragma unroll
            for (int n = 2; n <= PIMC_FAST_NMAX; ++n) { CX[n] = tcx * CX[n - 1] - CX[n - 2]; CY[n] = tcy * CY[n - 1] - CY[n - 2]; }   // cos(n t)
            double cs = 0.0;
#pragma unroll
            for (int a = 0; a <= PIMC_FAST_NMAX; ++a) {
                if (a > p.nmx) break;                                       // launch-uniform
                double row = 0.0;
#pragma unroll
                for (int b = 0; b <= PIMC_FAST_NMAX; ++b) { if (b > p.nmy) break; row += p.wcc[a * (PIMC_FAST_NMAX + 1) + b] * CY[b]; }
                cs += row * CX[a];
            }
            cs /= p.nang;
            return (p.sgn * p.depth) * (cs * cs);
        }
        for (int i = 0; i < p.nang; ++i) {
            const double t = (x * p.sn[i] + y * p.cs[i]) * p.scale;
            double s1, c1;
            pimc_sincos2pi(t - floor(t), &s1, &c1);
            if (p.helical) { const double sa = p.sn[i], ca = p.cs[i], s0 = s1; s1 = s0 * ca + c1 * sa; c1 = c1 * ca - s0 * sa; }   // + angle_i
            s += s1; c += c1;
        }
        s /= p.nang; c /= p.nang;
        return (p.sgn * p.depth) * (s * s + c * c); }
    default: return 0.0;
    }
}
// r . dV(r) of the virial estimator (measurement.jl:105)
__device__ __forceinline__ double d_rdv(const PotDev &p, double x, double y, int dim)
{
    if (p.dv_kind == PIMC_DV_ZERO) return 0.0;
    if (p.dv_kind == PIMC_DV_IDENTITY) { double s = x * x; if (dim > 1) s = s + y * y; return s; }
    double dx = 0.0, dy = 0.0;
    switch (p.kind) {
    case PIMC_POT_HARMONIC: dx = p.k * x; dy = p.k * y; break;
    case PIMC_POT_SIN2_1D: { double ph = 6.283185307179586 * x * p.scale; dx = p.depth * 2 * sin(ph) * cos(ph) * (6.283185307179586 * p.scale); break; }
    case PIMC_POT_LATTICE: {
        double s = 0, c = 0, sx = 0, sy = 0, cx = 0, cy = 0, f = 6.283185307179586 * p.scale;
        if (dim < 2) y = 0.0;
        for (int i = 0; i < p.nang; ++i) {
            double ph = 6.283185307179586 * (x * p.sn[i] + y * p.cs[i]) * p.scale + (p.helical ? p.ang[i] : 0.0);
            double s1, c1; sincos(ph, &s1, &c1);
            s += s1; c += c1; sx += c1 * f * p.sn[i]; sy += c1 * f * p.cs[i]; cx += -s1 * f * p.sn[i]; cy += -s1 * f * p.cs[i];
        }
        double n = p.nang; s /= n; c /= n; sx /= n; sy /= n; cx /= n; cy /= n;
        dx = p.sgn * p.depth * 2 * (s * sx + c * cx); dy = p.sgn * p.depth * 2 * (s * sy + c * cy); break; }
    default: break;
    }
    double s = x * dx; if (dim > 1) s = s + y * dy; return s;
}
__device__ __forceinline__ void d_grad(const PotDev &p, double x, double y, int dim, double *dx, double *dy)
{
    *dx = 0.0; *dy = 0.0;
    if (p.dv_kind == PIMC_DV_ZERO) return;
    if (p.dv_kind == PIMC_DV_IDENTITY) { *dx = x; *dy = y; return; }
    PotDev q = p; (void)q;
    switch (p.kind) {
    case PIMC_POT_HARMONIC: *dx = p.k * x; *dy = p.k * y; break;
    case PIMC_POT_SIN2_1D: { double ph = 6.283185307179586 * x * p.scale; *dx = p.depth * 2 * sin(ph) * cos(ph) * (6.283185307179586 * p.scale); break; }
    case PIMC_POT_LATTICE: {
        double s = 0, c = 0, sx = 0, sy = 0, cx = 0, cy = 0, f = 6.283185307179586 * p.scale;
        if (dim < 2) y = 0.0;
        for (int i = 0; i < p.nang; ++i) {
            double ph = 6.283185307179586 * (x * p.sn[i] + y * p.cs[i]) * p.scale + (p.helical ? p.ang[i] : 0.0);
            double s1, c1; sincos(ph, &s1, &c1);
            s += s1; c += c1; sx += c1 * f * p.sn[i]; sy += c1 * f * p.cs[i]; cx += -s1 * f * p.sn[i]; cy += -s1 * f * p.cs[i];
        }
        double n = p.nang; s /= n; c /= n; sx /= n; sy /= n; cx /= n; cy /= n;
        *dx = p.sgn * p.depth * 2 * (s * sx + c * cx); *dy = p.sgn * p.depth * 2 * (s * sy + c * cy); break; }
    default: break;
    }
}
__device__ __forceinline__ double d_lnK2(double ax, double ay, double bx, double by, int dim, double tau, double lambda, double L) // propagator.jl:16-19
{
    double dr = d_distance(ax, bx, L);
    double d2 = dr * dr;
    if (dim > 1) { dr = d_distance(ay, by, L); d2 = d2 + dr * dr; }
    return -d2 / (4 * lambda * tau);
}
__device__ __forceinline__ double d_norm2(double dx, double dy, int dim)
{
    double s = dx * dx; if (dim > 1) s = s + dy * dy; return sqrt(s);
}

// ---------------- pair propagator lookup (system.jl:17-34, propagator.jl:73-86) ----------------
__device__ __forceinline__ double d_tab_lookup(const DevSys &S, double x, double y)
{
    int n = S.tab_n;
    double h = (S.tab_hi - S.tab_lo) / (n - 1);
    double tx = (x - S.tab_lo) / h, ty = (y - S.tab_lo) / h;
    double fx0 = floor(tx), fy0 = floor(ty);
    if (fx0 < 0) fx0 = 0; if (fx0 > n - 2) fx0 = n - 2;
    if (fy0 < 0) fy0 = 0; if (fy0 > n - 2) fy0 = n - 2;
    int ix = (int)fx0, iy = (int)fy0;
    double fx = tx - fx0, fy = ty - fy0;
    const double *A = S.tab;
    double a00 = A[ix + (size_t)n * iy], a10 = A[ix + 1 + (size_t)n * iy];
    double a01 = A[ix + (size_t)n * (iy + 1)], a11 = A[ix + 1 + (size_t)n * (iy + 1)];
    double c0 = (1 - fx) * a00 + fx * a10;
    double c1 = (1 - fx) * a01 + fx * a11;
    return (1 - fy) * c0 + fy * c1;
}
__device__ __forceinline__ double d_lnU(const DevSys &S, double r1x, double r1y, double r2x, double r2y)
{
    if (!S.interactions || !S.tab) return 0.0;
    double dx = r1x - r2x, dy = r1y - r2y;
    double d2 = dx * dx; if (S.dim > 1) d2 = d2 + dy * dy;
    double rel0 = exp(-d2 / (4 * S.tau)) / (4 * 3.141592653589793 * S.tau);
    double p = 1 + d_tab_lookup(S, d_norm2(r1x, r1y, S.dim), d_norm2(r2x, r2y, S.dim)) / rel0;
    return p < 0.0 ? -S.mu : log(p);
}

// ---------------- cell list (nearest_neighbours.jl) ----------------
// floor((x + L) / w) through the reciprocal, re-evaluated with the true division whenever the product lies within a few ulp of an
// integer (the only case in which the two floors can differ)
__device__ __forceinline__ int d_floor_div(double s, double w, double inv)
{
    const double t = s * inv;
    double f = floor(t);
    const double fr = t - f, thr = (fabs(t) + 1.0) * 4e-15;
    if (!(fr >= thr && 1.0 - fr >= thr)) f = floor(s / w);
    return (int)f;
}
__device__ __forceinline__ int d_bin(const DevSys &S, double x, double y) // :26-33, 0-based, clamped
{
    const double inv = 1.0 / S.cellw;
    int ix = d_floor_div(x + S.L, S.cellw, inv);
    ix = ix < 0 ? 0 : (ix > S.nbins - 1 ? S.nbins - 1 : ix);
    if (S.dim == 1) return ix;
    int iy = d_floor_div(y + S.L, S.cellw, inv);
    iy = iy < 0 ? 0 : (iy > S.nbins - 1 ? S.nbins - 1 : iy);
    return ix + S.nbins * iy;
}
__device__ __forceinline__ int d_imod(int x, int n) { int m = x % n; return m < 0 ? m + n : m; }
// periodic wrap of a cell coordinate that left [0, n) by at most one cell (== d_imod for -1 <= x <= n)
__device__ __forceinline__ int d_wrap1(int x, int n) { return x < 0 ? x + n : (x >= n ? x - n : x); }
// stencil cell q of cell b, same order as bin_neighbors (:55-65)
__device__ __forceinline__ int d_stencil(const DevSys &S, int b, int q)
{
    if (S.dim == 2) {
        const int dx[9] = { 0, -1, 0, 1, -1, 1, -1, 0, 1 };
        const int dy[9] = { 0, 1, 1, 1, 0, 0, -1, -1, -1 };
        int y = b / S.nbins, x = b - y * S.nbins;
        return d_wrap1(x + dx[q], S.nbins) + S.nbins * d_wrap1(y + dy[q], S.nbins);
    }
    const int d1[3] = { 0, -1, 1 };
    return d_wrap1(b % S.nbins + d1[q], S.nbins);
}
// Distances.PeriodicEuclidean on the +L shifted coordinates.  mod(s1, p) = s1 - p * floor(s1 / p) is s1 itself whenever s1 < p
// (floor of a quotient below one is zero), which holds for every pair of teleported positions: no division on that path.
__device__ __forceinline__ double d_peuclid(const DevSys &S, double ax, double ay, double bx, double by)
{
    double p = 2 * S.L;
    double s1 = fabs((ax + S.L) - (bx + S.L));
    double s2 = s1 < p ? s1 : s1 - p * floor(s1 / p);
    double s3 = s2 < p - s2 ? s2 : p - s2;
    double acc = s3 * s3;
    if (S.dim > 1) {
        s1 = fabs((ay + S.L) - (by + S.L));
        s2 = s1 < p ? s1 : s1 - p * floor(s1 / p);
        s3 = s2 < p - s2 ? s2 : p - s2;
        acc = acc + s3 * s3;
    }
    return sqrt(acc);
}
__device__ __forceinline__ void d_cell_remove(const DevSys &S, int c, int j, int n)
{
    int b = S.bins[VIDX(S, c, n, j)];
    int *head = S.cell_head + HIDX(S, c, j, b);
    int *nxt = S.cell_next + NIDX(S, c, j, 0);          // successor of particle p: nxt[(size_t)p * M]
    const size_t M = (size_t)S.M;
    int p = *head, prev = -1;
    while (p >= 0 && p != n) { prev = p; p = nxt[p * M]; }
    if (p < 0) return;
    if (prev < 0) *head = nxt[p * M]; else nxt[prev * M] = nxt[p * M];
}
__device__ __forceinline__ void d_cell_insert(const DevSys &S, int c, int j, int n, int b)
{
    int *head = S.cell_head + HIDX(S, c, j, b);
    S.cell_next[NIDX(S, c, j, n)] = *head; *head = n;
    S.bins[VIDX(S, c, n, j)] = b;
    S.mult[VIDX(S, c, n, j)] = 1;
}
__device__ __forceinline__ void d_cell_update(const DevSys &S, int c, int j, int n, double x, double y) // update_nn_bead! :198-209
{
    if (!S.need_cells) return;
    d_cell_remove(S, c, j, n);
    d_cell_insert(S, c, j, n, d_bin(S, x, y));
}
// Stencil prefetch: the (up to) nine list heads of a query are independent loads, and so are the position / successor of the first
// occupant of every non-empty cell.  Issuing them together turns the 9..18 dependent HBM round trips of a naive stencil walk into
// two; longer lists (rare at the cell widths of the examples) continue with dependent loads.  Visiting order is unchanged.
struct NbFirst { int h[9], n[9]; double x[9], y[9]; };
__device__ __forceinline__ void d_nb_first(const DevSys &S, int c, int j, int b, NbFirst &F)
{
    const int nst = S.dim == 2 ? 9 : 3;
#pragma unroll
    for (int q = 0; q < 9; ++q) F.h[q] = q < nst ? S.cell_head[HIDX(S, c, j, d_stencil(S, b, q))] : -1;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        F.n[q] = -1; F.x[q] = 0.0; F.y[q] = 0.0;
        if (F.h[q] >= 0) {
            F.n[q] = S.cell_next[NIDX(S, c, j, F.h[q])];
            F.x[q] = S.r[RIDX(S, c, F.h[q], 0, j)];
            if (S.dim > 1) F.y[q] = S.r[RIDX(S, c, F.h[q], 1, j)];
        }
    }
}
// iterate the stencil occupants in bin_neighbors order (:55-65): BODY sees o (occupant), ox, oy (its position at slice j)
#define PIMC_FOR_STENCIL(S_, c_, j_, F_, ...)                                                                       \
    _Pragma("unroll 1") for (int q_ = 0; q_ < 9; ++q_) {   /* one copy of the body: code size (instruction fetch) over registers */ \
        int o = (F_).h[q_], on_ = (F_).n[q_]; double ox = (F_).x[q_], oy = (F_).y[q_];                                \
        while (o >= 0) {                                                                                             \
            __VA_ARGS__                                                                                              \
            o = on_;                                                                                                 \
            if (o >= 0) {                                                                                            \
                on_ = (S_).cell_next[NIDX(S_, c_, j_, o)];                                                           \
                ox = (S_).r[RIDX(S_, c_, o, 0, j_)]; oy = (S_).dim > 1 ? (S_).r[RIDX(S_, c_, o, 1, j_)] : 0.0;       \
            }                                                                                                        \
        }                                                                                                            \
    }
// find_nn (:156-179): nearest stencil occupant (periodic metric on +L shifted coordinates), -1 if none
static __device__ __noinline__ int d_find_nn(const DevSys &S, int c, double x, double y, int j, int exc)
{
    int best = -1;
    double bd = 0.0;
    NbFirst F; d_nb_first(S, c, j, d_bin(S, x, y), F);
    PIMC_FOR_STENCIL(S, c, j, F, {
        if (o != exc) {
            double d = d_peuclid(S, ox, oy, x, y);
            if (best < 0 || d < bd) { best = o; bd = d; }
        }
    })
    return best;
}
// hard-core test used by hardspherelevy! (helper.jl:167-170) and move_polymer! (helper.jl:385-390)
__device__ __forceinline__ bool d_hardcore_hit(const DevSys &S, int c, double x, double y, int j, int exc)
{
    int nn = d_find_nn(S, c, x, y, j, exc);
    if (nn < 0) return false;
    double dx = d_distance(x, S.r[RIDX(S, c, nn, 0, j)], S.L);
    double dy = S.dim > 1 ? d_distance(y, S.r[RIDX(S, c, nn, 1, j)], S.L) : 0.0;
    return d_norm2(dx, dy, S.dim) < S.a;
}

// ---------------- Gaussian source ----------------
struct GSrc { const double *xi; pimc_stream st; uint32_t slot, kind; const double *tab; };
__device__ __forceinline__ void d_gauss(const GSrc &g, int dim, int bead, int retry, double &g0, double &g1)
{
    if (g.xi) { g0 = g.xi[(size_t)(bead - 1) * dim]; g1 = dim > 1 ? g.xi[(size_t)(bead - 1) * dim + 1] : 0.0; return; }
    pimc_gauss_pair_t(pimc_draw(g.st, g.slot, g.kind, (uint32_t)retry, (uint32_t)bead), g.tab, &g0, &g1);
}

// hardspherelevy! (helper.jl:141-181) == levy! (helper.jl:118-139) when a == 0.
// Endpoints by value; writes the TELEPORTED rows 0..rows-1 to px/py and V(row) to pv. slice j0 is 1-based.
__device__ __forceinline__ int d_bridge(const DevSys &S, int c, double bx, double by, double ex, double ey, int rows, int j0,
                                        int exc, const GSrc &g, double *px, double *py, double *pv)
{
    const double L = S.L; const int dim = S.dim;
    if (fabs(bx - ex) > L) ex += d_sign(bx) * (2 * L);
    if (dim > 1 && fabs(by - ey) > L) ey += d_sign(by) * (2 * L);
    int m = rows - 2;
    double qx = bx, qy = by;
    double tx = d_teleport(bx, L), ty = dim > 1 ? d_teleport(by, L) : 0.0;
    px[0] = tx; if (dim > 1) py[0] = ty;
    if (pv) pv[0] = d_pot(S.pot, tx, ty, dim);
    for (int j = 1; j <= m; ++j) {
        double alpha = (double)(m + 1 - j) / (double)(m + 2 - j);
        double sig = sqrt(2 * S.lambda * alpha * S.tau);
        double om = 1 - alpha;
        double nx = 0.0, ny = 0.0; long long ctr = 0; bool pass = true;
        while (pass) {
            pass = false; ctr += 1;
            if (ctr > S.ctr) { pass = true; break; }
            double g0, g1; d_gauss(g, dim, j, (int)(ctr - 1), g0, g1);
            nx = alpha * qx + om * ex + g0 * sig;
            if (dim > 1) ny = alpha * qy + om * ey + g1 * sig;
            tx = d_teleport(nx, L); ty = dim > 1 ? d_teleport(ny, L) : 0.0;
            if (S.a > 0.0) {
                int sl = (j0 + j - 1) % S.M; // mod1(j0 + j, M) - 1
                if (d_hardcore_hit(S, c, tx, ty, sl, exc)) pass = true;
            }
        }
        if (pass) return 0;
        qx = nx; qy = ny;
        px[j] = tx; if (dim > 1) py[j] = ty;
        if (pv) pv[j] = d_pot(S.pot, tx, ty, dim);
    }
    tx = d_teleport(ex, L); ty = dim > 1 ? d_teleport(ey, L) : 0.0;
    px[rows - 1] = tx; if (dim > 1) py[rows - 1] = ty;
    if (pv) pv[rows - 1] = d_pot(S.pot, tx, ty, dim);
    return 1;
}

__device__ __forceinline__ bool d_metropolis(double delta, double u) { return (delta >= 1.0) || (delta > u); } // helper.jl:3-5
