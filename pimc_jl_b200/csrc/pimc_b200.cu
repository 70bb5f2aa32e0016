// pimc_b200.cu -- kernels and C ABI of libpimc_b200.so (sm_100a).  See include/pimc_b200.h and DESIGN.md.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cstdio>
#include "pimc_moves.cuh"
#include "pimc_launch.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <new>

// =====================================================================================================
// kernels
// =====================================================================================================

// init_world (system.jl:36-78): one CTA per chain.  a == 0: worldlines are independent -> one thread per particle.
// a > 0: thread 0 walks the particles in order (rejection against earlier particles).
__global__ void k_init_world(DevSys S)
{
    const int c = blockIdx.x, M = S.M, N = S.N, dim = S.dim;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, 0);
    const bool serial = S.a > 0.0;
    for (int n = serial ? 0 : threadIdx.x; n < N; n += serial ? 1 : blockDim.x) {
        if (serial && threadIdx.x != 0) break;
        double *px = S.r + RIDX(S, c, n, 0, 0), *py = px + M;
        GSrc g; g.xi = nullptr; g.st = st; g.slot = (uint32_t)n; g.kind = PIMC_K_INIT; g.tab = S.logtab;
        uint32_t levy_calls = 0;
        auto start = [&](uint32_t attempt) {
            pimc_u4 w = pimc_draw(st, (uint32_t)n, PIMC_K_INIT0, attempt, 0);
            px[0] = 2 * S.L * (pimc_u01_co(w.w[0], w.w[1]) - 0.5); px[M - 1] = px[0];
            if (dim > 1) { py[0] = 2 * S.L * (pimc_u01_co(w.w[2], w.w[3]) - 0.5); py[M - 1] = py[0]; }
        };
        auto levy = [&]() {
            // plain levy! (helper.jl:118-139); the retry field of the draw address carries the levy! call index
            const GSrc &gg = g;
            double bx = px[0], by = dim > 1 ? py[0] : 0.0, ex = px[M - 1], ey = dim > 1 ? py[M - 1] : 0.0;
            const double L = S.L;
            if (fabs(bx - ex) > L) ex += d_sign(bx) * (2 * L);
            if (dim > 1 && fabs(by - ey) > L) ey += d_sign(by) * (2 * L);
            int m = M - 2; double qx = bx, qy = by;
            for (int j = 1; j <= m; ++j) {
                double alpha = (double)(m + 1 - j) / (double)(m + 2 - j);
                double sig = sqrt(2 * S.lambda * alpha * S.tau), om = 1 - alpha, g0, g1;
                pimc_gauss_pair_t(pimc_draw(gg.st, gg.slot, PIMC_K_INIT, levy_calls, (uint32_t)j), S.logtab, &g0, &g1);
                double nx = alpha * qx + om * ex + g0 * sig, ny = 0.0;
                if (dim > 1) ny = alpha * qy + om * ey + g1 * sig;
                qx = nx; qy = ny;
                px[j] = d_teleport(nx, L); if (dim > 1) py[j] = d_teleport(ny, L);
            }
            px[0] = d_teleport(bx, L); px[M - 1] = d_teleport(ex, L);
            if (dim > 1) { py[0] = d_teleport(by, L); py[M - 1] = d_teleport(ey, L); }
            levy_calls += 1;
        };
        bool pass = n != 0; long long ctr = 0;
        while (pass) {
            pass = false; ctr += 1;
            if (ctr > 10000) break; // reference: @error and empty world; here the last attempt is kept
            start((uint32_t)(ctr - 1)); levy();
            for (int mm = 0; mm < M; ++mm)
                for (int i = 0; i < n; ++i) {
                    double dx = d_distance(S.r[RIDX(S, c, i, 0, mm)], px[mm], S.L);
                    double dy = dim > 1 ? d_distance(S.r[RIDX(S, c, i, 1, mm)], py[mm], S.L) : 0.0;
                    if (d_norm2(dx, dy, dim) < S.a) { levy(); pass = true; }
                }
        }
        if (n == 0) { start(0); levy(); }
        for (int mm = 0; mm < M; ++mm) {
            int mn = mm == M - 1 ? 0 : mm + 1;
            S.Vl[VIDX(S, c, n, mm)] = -0.5 * S.tau * (d_pot(S.pot, px[mm], dim > 1 ? py[mm] : 0.0, dim) + d_pot(S.pot, px[mn], dim > 1 ? py[mn] : 0.0, dim));
        }
        S.next[(size_t)c * N + n] = n;
    }
}

// link cache of arbitrary (permuted) worlds: V[n][m] = lnV(r[n][m], next bead)  (system.jl:72-74)
__global__ void k_relink(DevSys S, int c0, int nc)
{
    const int M = S.M, N = S.N, dim = S.dim;
    size_t total = (size_t)nc * N * M;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int c = c0 + (int)(idx / ((size_t)N * M)); int rem = (int)(idx % ((size_t)N * M));
        int n = rem / M, j = rem - n * M;
        int q = j == M - 1 ? S.next[(size_t)c * N + n] : n, jn = j == M - 1 ? 0 : j + 1;
        double a = d_pot(S.pot, S.r[RIDX(S, c, n, 0, j)], dim > 1 ? S.r[RIDX(S, c, n, 1, j)] : 0.0, dim);
        double b = d_pot(S.pot, S.r[RIDX(S, c, q, 0, jn)], dim > 1 ? S.r[RIDX(S, c, q, 1, jn)] : 0.0, dim);
        S.Vl[VIDX(S, c, n, j)] = -0.5 * S.tau * (a + b);
    }
}

// update_nnbins! (nearest_neighbours.jl:182-196): one thread per (chain, slice) rebuilds that slice's lists
__global__ void k_cells_build(DevSys S, int c0, int nc)
{
    const int M = S.M, N = S.N;
    size_t total = (size_t)nc * M;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int c = c0 + (int)(idx / M), j = (int)(idx % M);
        for (int b = 0; b < S.ncell; ++b) S.cell_head[HIDX(S, c, j, b)] = -1;
        for (int n = N - 1; n >= 0; --n) // push-front in reverse keeps ascending particle order in every list
            d_cell_insert(S, c, j, n, d_bin(S, S.r[RIDX(S, c, n, 0, j)], S.dim > 1 ? S.r[RIDX(S, c, n, 1, j)] : 0.0));
    }
}
__global__ void k_bins_export(DevSys S, int c0, int nc, long long *out)
{
    const int M = S.M, N = S.N;
    size_t total = (size_t)nc * N * M;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int c = c0 + (int)(idx / ((size_t)N * M)); int rem = (int)(idx % ((size_t)N * M));
        int n = rem / M, j = rem - n * M;
        out[idx] = d_bin(S, S.r[RIDX(S, c, n, 0, j)], S.dim > 1 ? S.r[RIDX(S, c, n, 1, j)] : 0.0) + 1;
    }
}

// ---- estimator / action hooks ----
__global__ void k_energy_now(DevSys S, double *E, double *Ev, double *parts)
{
    __shared__ double red[96];
    int c = blockIdx.x; double e, ev, p3[3];
    d_energy_block(S, c, red, &e, &ev, p3);
    if (threadIdx.x == 0) { E[c] = e; Ev[c] = ev; if (parts) { parts[3 * c] = p3[0]; parts[3 * c + 1] = p3[1]; parts[3 * c + 2] = p3[2]; } }
}
__global__ void k_density_now(DevSys S, DeDev D) { d_density_block(S, blockIdx.x, D); }
__global__ void k_action(DevSys S, double *cached, double *recomputed)
{
    __shared__ double red[64];
    const int c = blockIdx.x, M = S.M, N = S.N, dim = S.dim;
    double a = 0.0, b = 0.0;
    for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) {
        int n = idx / M, j = idx - n * M;
        int q = j == M - 1 ? S.next[(size_t)c * N + n] : n, jn = j == M - 1 ? 0 : j + 1;
        a += S.Vl[VIDX(S, c, n, j)];
        b += -0.5 * S.tau * (d_pot(S.pot, S.r[RIDX(S, c, n, 0, j)], dim > 1 ? S.r[RIDX(S, c, n, 1, j)] : 0.0, dim) +
                             d_pot(S.pot, S.r[RIDX(S, c, q, 0, jn)], dim > 1 ? S.r[RIDX(S, c, q, 1, jn)] : 0.0, dim));
    }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = a; red[32 + (threadIdx.x >> 5)] = b; }
    __syncthreads();
    if (threadIdx.x == 0) { a = 0; b = 0; for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) { a += red[i]; b += red[32 + i]; } cached[c] = a; recomputed[c] = b; }
}
// mean over chains of the energy series: one CTA per measurement index, coalesced over the chain-fastest layout, fixed
// reduction tree (deterministic)
__global__ void k_energy_chain_mean(const double *E, const double *Ev, int C, long long k0, double *mE, double *mEv)
{
    __shared__ double ra[8], rb[8];
    const long long k = k0 + blockIdx.x;
    double a = 0.0, b = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { a += E[(size_t)k * C + c]; b += Ev[(size_t)k * C + c]; }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) { a = 0.0; b = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += ra[i]; b += rb[i]; } mE[blockIdx.x] = a / C; mEv[blockIdx.x] = b / C; }
}
// chain SUMS of a block of measurements, interleaved (E, Ev) per index: the buffer the ranks all-reduce
__global__ void k_energy_chain_sum(const double *E, const double *Ev, int C, long long k0, double *g)
{
    __shared__ double ra[8], rb[8];
    const long long k = k0 + blockIdx.x;
    double a = 0.0, b = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { a += E[(size_t)k * C + c]; b += Ev[(size_t)k * C + c]; }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) { a = 0.0; b = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += ra[i]; b += rb[i]; } g[2 * k] = a; g[2 * k + 1] = b; }
}
__global__ void k_energy_global_mean(const double *g, long long k0, long long n, double inv /* total number of chains */, double *oE, double *oEv)
{
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    oE[k] = g[2 * (k0 + k)] / inv; oEv[k] = g[2 * (k0 + k) + 1] / inv;
}
__global__ void k_energy_chain_series(const double *E, const double *Ev, int C, int c, long long k0, long long n, double *oE, double *oEv)
{
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    oE[k] = E[(size_t)(k0 + k) * C + c]; oEv[k] = Ev[(size_t)(k0 + k) * C + c];
}
__global__ void k_dens_to_double(const unsigned long long *d, size_t n, double *o)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = (double)d[i];
}

// ---- pure-function hooks ----
__global__ void k_distance(long long n, const double *a, const double *b, double L, double *o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = d_distance(a[i], b[i], L); }
__global__ void k_teleport(long long n, const double *a, double L, double *o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = d_teleport(a[i], L); }
__global__ void k_lnK(long long n, const double *a, const double *b, int dim, double tau, double lambda, double L, double *o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      o[i] = d_lnK2(a[i * dim], dim > 1 ? a[i * dim + 1] : 0.0, b[i * dim], dim > 1 ? b[i * dim + 1] : 0.0, dim, tau, lambda, L); }
__global__ void k_lnV(long long n, const double *a, const double *b, int dim, double tau, PotDev p, double *o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      o[i] = -0.5 * tau * (d_pot(p, a[i * dim], dim > 1 ? a[i * dim + 1] : 0.0, dim) + d_pot(p, b[i * dim], dim > 1 ? b[i * dim + 1] : 0.0, dim)); }
__global__ void k_pot(long long n, const double *a, int dim, PotDev p, double *V, double *dV)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      double x = a[i * dim], y = dim > 1 ? a[i * dim + 1] : 0.0;
      if (V) V[i] = d_pot(p, x, y, dim);
      if (dV) { double dx, dy; d_grad(p, x, y, dim, &dx, &dy); dV[i * dim] = dx; if (dim > 1) dV[i * dim + 1] = dy; } } }
// levy! (helper.jl:118-139): one thread per bridge, in place
__global__ void k_levy(double *r, int rows, int dim, double tau, double L, double lambda, const double *xi, long long nb)
{
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
        DevSys S; S.dim = dim; S.M = rows; S.tau = tau; S.L = L; S.lambda = lambda; S.a = 0.0; S.ctr = 1; S.pot.kind = PIMC_POT_ZERO;
        double *px = r + (size_t)b * rows * dim, *py = px + rows;
        GSrc g; g.xi = xi + (size_t)b * (rows - 2) * dim; g.slot = 0; g.kind = 0; g.tab = nullptr;
        d_bridge(S, 0, px[0], dim > 1 ? py[0] : 0.0, px[rows - 1], dim > 1 ? py[rows - 1] : 0.0, rows, 1, -1, g, px, py, nullptr);
    }
}
__global__ void k_gauss(unsigned long long seed, uint32_t chain, unsigned long long iter, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead0, long long n, double *g, const double *tab)
{
    pimc_stream st = pimc_stream_make(seed, chain, iter);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        pimc_gauss_pair_t(pimc_draw(st, slot, kind, retry, bead0 + (uint32_t)i), tab, g + 2 * i, g + 2 * i + 1);
}

// ---- explicit single-move hooks (one warp) ----
struct MoveOut { double wi, wu; int acc; };
__global__ void k_reshape_linear_explicit(DevSys S, int c, int n, int j0, int m, const double *xi, double u, int commit, MoveOut *out, double *rp)
{
    if (threadIdx.x != 0) return;
    GSrc g; g.xi = xi; g.slot = 0; g.kind = 0; g.tab = nullptr;
    out->acc = d_reshape_linear(S, c, n, j0, m, g, u, commit, 0, &out->wi, &out->wu);
    if (rp && out->acc >= 0) { // teleported proposal rows (m+1) x dim, column-major
        const double *px = S.prop + RIDX(S, c, 0, 0, 0), *py = px + S.M;
        // row m (last endpoint) was consumed in place only for pv, positions are intact
        for (int j = 0; j <= m; ++j) { rp[j] = px[j]; if (S.dim > 1) rp[(m + 1) + j] = py[j]; }
    }
}
__global__ void k_reshape_swap_explicit(DevSys S, int c, int n1, int n2, int j0, int m, const double *xi1, const double *xi2, double u, int commit, MoveOut *out)
{
    if (threadIdx.x != 0) return;
    GSrc g1, g2; g1.xi = xi1; g1.slot = 0; g1.kind = 0; g1.tab = nullptr; g2 = g1; g2.xi = xi2;
    out->acc = d_reshape_swap(S, c, n1, n2, j0, m, g1, g2, u, commit, &out->wi, &out->wu);
}
__global__ void k_com_explicit(DevSys S, int c, int n, const double *d, double u, int commit, MoveOut *out)
{
    DSrc ds; ds.d = d; ds.slot = 0;
    double wi, wu; int r = d_com_warp(S, c, n, 0.0, ds, u, commit, &wi, &wu, nullptr);
    if (threadIdx.x == 0) { out->acc = r; out->wi = wi; out->wu = wu; }
}
__global__ void k_swap_weights(DevSys S, int c, int n1, int j0, int m, double *w)
{ if (threadIdx.x == 0) d_swap_weights(S, c, n1, j0, m, w); }
__global__ void k_find_nn(DevSys S, int c, double x, double y, int j, int exc, long long *out)
{ if (threadIdx.x == 0) { int nn = d_find_nn(S, c, x, y, j, exc); out[0] = nn < 0 ? -1 : nn + 1; } }
__global__ void k_find_nns(DevSys S, int c, double x, double y, int j, int exc, long long *out, long long cap, long long *count)
{
    if (threadIdx.x != 0) return;
    int b = d_bin(S, x, y), nst = S.dim == 2 ? 9 : 3; long long cnt = 0;
    for (int q = 0; q < nst; ++q)
        for (int o = S.cell_head[HIDX(S, c, j, d_stencil(S, b, q))]; o >= 0; o = S.cell_next[NIDX(S, c, j, o)]) {
            if (o == exc) continue;
            if (d_peuclid(S, S.r[RIDX(S, c, o, 0, j)], S.dim > 1 ? S.r[RIDX(S, c, o, 1, j)] : 0.0, x, y) <= S.cellw) { if (cnt < cap) out[cnt] = o + 1; cnt++; }
        }
    *count = cnt;
}

// =====================================================================================================
// host side: C ABI
// =====================================================================================================
struct UpdHost { bool used; };
struct pimc_handle {
    pimc_config cfg;
    DevSys S;
    DevTables T;        // host mirror
    DevTables *dT;
    DevSys *dS;         // device copy of S
    cudaStream_t stream;
    char err[512];
    unsigned long long iter;
    long long N_MC, Nctr;
    int nupd, nen, nde;
    long long en_cap[PIMC_MAXE];
    long long en_count[PIMC_MAXE];   // samples every Energy object holds (its own count, measurement.jl:119-120)
    unsigned char *mdone;            // [C] per-chain flag: Energy of the current measurement evaluated inside the sweep launch
    int opt_fuse_energy, opt_isweep;
    int isw_backoff;                 // runs left on the sequential kernel before the optimistic sweep is probed again (dense regimes: most proposals replay)
    double isw_replay_frac;          // replays per proposal of the last optimistic run
    int *queue;                      // chain queue of the chain-major persistent kernel
    int *nw_head, *nw_next;          // second cell list over proposed positions (optimistic sweep of interacting worldlines), lazily allocated
    unsigned long long *dstats;
    std::vector<void *> allocs;
    long long de_ndata[PIMC_MAXD];
    double r_a; double vol;
    int opt_sweep_impl, opt_faithful_impl;
    double *dens_out; size_t dens_out_n;   // density read-out staging
    double *fscr;       // HBM scratch for warp-cooperative proposals that do not fit shared memory (lazily allocated)
    cudaEvent_t ev0, ev1;
    int device;
    PcDev pc[PIMC_MAXP]; long long pc_ndata[PIMC_MAXP]; int npc;      // g(r) objects
    WiDev wi[PIMC_MAXW]; long long wi_count[PIMC_MAXW]; int nwi;       // winding-number objects
    SkDev sk[PIMC_MAXS]; long long sk_ndata[PIMC_MAXS]; int nsk;       // structure-factor objects
    double *g_f64; size_t g_f64_n;     // structure-factor read-out staging (sums + ndata, all-reduced over the ranks)
    // multi-GPU: one rank per handle; estimator blocks are all-reduced on a side stream while the next block's moves run
    void *comm;                        // ncclComm_t, null = single GPU
    int nranks, rank; long long chains_total;
    cudaStream_t cstream; cudaEvent_t cev;
    double *g_en[PIMC_MAXE];           // [cap][2] chain sums of (E, Ev) per measurement index, all-reduced over the ranks
    long long g_en_upto[PIMC_MAXE];    // rows of g_en that hold reduced values
    unsigned long long *g_u64; size_t g_u64_n;   // density read-out staging (counters + ndata)
};
static char g_err[512] = "";
static long long g_launches = 0;
#define LAUNCHED() (++g_launches)

#define SETERR(h, ...) do { if (h) snprintf((h)->err, sizeof((h)->err), __VA_ARGS__); else snprintf(g_err, sizeof g_err, __VA_ARGS__); } while (0)
#define CK(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { SETERR(h, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); return PIMC_ERR_CUDA; } } while (0)

template <typename Tp> static int dalloc(pimc_handle *h, Tp **p, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(Tp) > 0 ? n * sizeof(Tp) : 8);
    if (e != cudaSuccess) { SETERR(h, "cudaMalloc of %zu bytes failed: %s", n * sizeof(Tp), cudaGetErrorString(e)); return PIMC_ERR_NOMEM; }
    e = cudaMemset(q, 0, n * sizeof(Tp) > 0 ? n * sizeof(Tp) : 8);
    if (e != cudaSuccess) { SETERR(h, "cudaMemset failed: %s", cudaGetErrorString(e)); return PIMC_ERR_CUDA; }
    h->allocs.push_back(q);
    *p = (Tp *)q;
    return PIMC_OK;
}
static void pot_to_dev(const pimc_potential *p, PotDev *d)
{
    memset(d, 0, sizeof *d);
    d->kind = p->kind; d->dv_kind = p->dv_kind; d->k = p->k; d->depth = p->depth; d->scale = p->scale; d->sgn = p->sgn;
    d->nang = p->nang; d->helical = p->helical;
    for (int i = 0; i < p->nang && i < PIMC_MAX_ANGLES; ++i) { d->ang[i] = p->ang[i]; d->sn[i] = sin(p->ang[i]); d->cs[i] = cos(p->ang[i]); }
    // commensurate projections?  base = (smallest non-zero |projection|) / q for the smallest q in 1..8 that makes every projection an integer
    // multiple <= PIMC_FAST_NMAX of it (to 1e-12); PIMC_NO_FAST_LATTICE=1 keeps the one-sincos-per-beam evaluation (A/B)
    d->fast = 0;
    if (p->kind == PIMC_POT_LATTICE && p->nang > 0 && !getenv("PIMC_NO_FAST_LATTICE")) {
        auto fit = [&](const double *v, signed char *m, double *base, int *nmax) {
            double mn = 0.0;
            for (int i = 0; i < p->nang; ++i) { const double a = fabs(v[i]); if (a > 1e-12 && (mn == 0.0 || a < mn)) mn = a; }
            if (mn == 0.0) { for (int i = 0; i < p->nang; ++i) m[i] = 0; *base = 0.0; *nmax = 0; return true; }
            for (int q = 1; q <= 8; ++q) {
                const double b = mn / q; bool ok = true; int mx = 0;
                for (int i = 0; i < p->nang && ok; ++i) {
                    const double r = v[i] / b, n = nearbyint(r);
                    if (fabs(r - n) > 1e-12 * (fabs(r) + 1.0) || fabs(n) > PIMC_FAST_NMAX) ok = false;
                    else { m[i] = (signed char)n; if (abs((int)n) > mx) mx = abs((int)n); }
                }
                if (ok) { *base = b; *nmax = mx; return true; }
            }
            return false;
        };
        double bx, by; int nx, ny; signed char mx[PIMC_MAX_ANGLES], my[PIMC_MAX_ANGLES];
        if (!p->helical && fit(d->sn, mx, &bx, &nx) && fit(d->cs, my, &by, &ny)) {
            // cos(ma tx + mb ty) = cos|ma| cos|mb| - sgn(ma) sgn(mb) sin|ma| sin|mb| ;  sin(...) = sgn(ma) sin|ma| cos|mb| + sgn(mb) cos|ma| sin|mb|
            int wcc[9][9] = {}, wss[9][9] = {}, wsc[9][9] = {}, wcs[9][9] = {};
            for (int i = 0; i < p->nang; ++i) {
                const int a = abs(mx[i]), b = abs(my[i]), sa = (mx[i] > 0) - (mx[i] < 0), sb = (my[i] > 0) - (my[i] < 0);
                wcc[a][b] += 1; wss[a][b] -= sa * sb; wsc[a][b] += sa; wcs[a][b] += sb;
            }
            bool sym = true;
            for (int a = 0; a < 9; ++a) for (int b = 0; b < 9; ++b) sym = sym && wss[a][b] == 0 && wsc[a][b] == 0 && wcs[a][b] == 0;
            if (sym) {
                d->fast = 1; d->bx = bx * p->scale; d->by = by * p->scale; d->nmx = nx; d->nmy = ny;
                for (int a = 0; a < 9; ++a) for (int b = 0; b < 9; ++b) d->wcc[a * 9 + b] = (double)wcc[a][b];
            }
        }
    }
}
static int grid_for(size_t n, int block) { size_t g = (n + block - 1) / block; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1; return (int)g; }

extern "C" int pimc_version(void) { return 100; }
extern "C" int64_t pimc_launch_count(void) { return g_launches; }

// dependent-free DFMA streams: 8 accumulators per thread, 4096 FMAs each per loop
__global__ void k_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
extern "C" int pimc_measure_fp64_peak(double *tflops)
{
    int nd = 0; if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) { snprintf(g_err, sizeof g_err, "no CUDA device"); return PIMC_ERR_CUDA; }
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 2000;
    double *o; if (cudaMalloc(&o, sizeof(double) * blocks * threads) != cudaSuccess) return PIMC_ERR_NOMEM;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k_fp64_peak<<<blocks, threads>>>(o, iters); LAUNCHED(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * 64 * (double)iters * blocks * threads;
        if (rep > 0 && ms > 0 && fl / ms * 1e-9 > best) best = fl / ms * 1e-9;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(o);
    if (cudaGetLastError() != cudaSuccess) return PIMC_ERR_CUDA;
    *tflops = best; return PIMC_OK;
}
extern "C" const char *pimc_last_error(const pimc_handle *h) { return h ? h->err : g_err; }

static void comm_teardown(pimc_handle *h);
extern "C" void pimc_destroy(pimc_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    comm_teardown(h);
    for (void *p : h->allocs) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
}


// =====================================================================================================
// multi-GPU inside the library: chains shard over ranks with no data-path collective (chain_offset makes the result
// independent of the sharding); the only exchange is the all-reduce of estimator blocks.  NCCL is bound at run time (dlopen of
// libnccl.so.2: the copy a host process -- torch, MPI.jl -- already mapped is reused), so the library loads on boxes without it.
// =====================================================================================================
#include <dlfcn.h>
#include <nccl.h>
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
    bool ok;
};
static NcclApi *nccl_api()
{
    static NcclApi api; static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true; api.ok = false;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);        // already mapped by the host process?
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return nullptr;
#define NSYM(field, name) do { *(void **)(&api.field) = dlsym(lib, name); if (!api.field) return nullptr; } while (0)
    NSYM(GetUniqueId, "ncclGetUniqueId"); NSYM(CommInitRank, "ncclCommInitRank"); NSYM(CommInitAll, "ncclCommInitAll"); NSYM(CommDestroy, "ncclCommDestroy");
    NSYM(AllReduce, "ncclAllReduce"); NSYM(GetErrorString, "ncclGetErrorString"); NSYM(GetVersion, "ncclGetVersion");
#undef NSYM
    api.ok = true; return &api;
}
#define NK(h, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { SETERR(h, "NCCL error %s at %s:%d (%s)", nccl_api()->GetErrorString(r_), __FILE__, __LINE__, #call); return PIMC_ERR_CUDA; } } while (0)

static void comm_teardown(pimc_handle *h)
{
    if (h->comm && nccl_api()) { cudaStreamSynchronize(h->cstream); nccl_api()->CommDestroy((ncclComm_t)h->comm); }
    h->comm = nullptr;
    if (h->cstream) cudaStreamDestroy(h->cstream);
    if (h->cev) cudaEventDestroy(h->cev);
    h->cstream = nullptr; h->cev = nullptr;
}
static int comm_finish_init(pimc_handle *h, ncclComm_t comm, int nranks, int rank)
{
    h->comm = comm; h->nranks = nranks; h->rank = rank;
    CK(h, cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
    CK(h, cudaEventCreateWithFlags(&h->cev, cudaEventDisableTiming));
    // total number of chains over the ranks (shards may be uneven): first collective of the communicator
    if (h->g_u64_n < 4) { int rc = dalloc(h, &h->g_u64, 4); if (rc) return rc; h->g_u64_n = 4; }
    unsigned long long c = (unsigned long long)h->S.C;
    CK(h, cudaMemcpyAsync(h->g_u64, &c, 8, cudaMemcpyHostToDevice, h->cstream));
    NK(h, nccl_api()->AllReduce(h->g_u64, h->g_u64, 1, ncclUint64, ncclSum, comm, h->cstream));
    CK(h, cudaMemcpyAsync(&c, h->g_u64, 8, cudaMemcpyDeviceToHost, h->cstream));
    CK(h, cudaStreamSynchronize(h->cstream));
    h->chains_total = (long long)c;
    for (int i = 0; i < h->nen; ++i) h->g_en_upto[i] = 0;
    return PIMC_OK;
}
extern "C" int pimc_comm_get_unique_id(void *id128)
{
    if (!id128) return PIMC_ERR_INVALID;
    NcclApi *a = nccl_api(); if (!a) { snprintf(g_err, sizeof g_err, "libnccl.so.2 not found (dlopen)"); return PIMC_ERR_UNSUPPORTED; }
    static_assert(sizeof(ncclUniqueId) == PIMC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id; ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) { snprintf(g_err, sizeof g_err, "ncclGetUniqueId: %s", a->GetErrorString(r)); return PIMC_ERR_CUDA; }
    memcpy(id128, &id, sizeof id); return PIMC_OK;
}
extern "C" int pimc_comm_init(pimc_handle *h, int32_t nranks, int32_t rank, const void *id128)
{
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return PIMC_ERR_INVALID;
    if (h->comm) { SETERR(h, "communicator already initialised"); return PIMC_ERR_STATE; }
    NcclApi *a = nccl_api(); if (!a) { SETERR(h, "libnccl.so.2 not found (dlopen)"); return PIMC_ERR_UNSUPPORTED; }
    CK(h, cudaSetDevice(h->device));
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    ncclComm_t comm; NK(h, a->CommInitRank(&comm, nranks, id, rank));
    return comm_finish_init(h, comm, nranks, rank);
}
// one process, n handles on n different devices (the ncclCommInitAll form of SURVEY 8e).  Afterwards every handle must be driven from its
// own host thread (a collective read-out blocks until all ranks have issued it).
extern "C" int pimc_comm_init_all(pimc_handle **hs, int32_t n)
{
    if (!hs || n < 1 || n > 64) return PIMC_ERR_INVALID;
    NcclApi *a = nccl_api(); if (!a) { SETERR(hs[0], "libnccl.so.2 not found (dlopen)"); return PIMC_ERR_UNSUPPORTED; }
    int devs[64]; ncclComm_t comms[64];
    for (int i = 0; i < n; ++i) { if (!hs[i] || hs[i]->comm) return PIMC_ERR_STATE; devs[i] = hs[i]->device; for (int k = 0; k < i; ++k) if (devs[k] == devs[i]) { SETERR(hs[0], "handles %d and %d share device %d", k, i, devs[i]); return PIMC_ERR_INVALID; } }
    NK(hs[0], a->CommInitAll(comms, n, devs));
    // the first collective (chain count) is issued for every rank before any of them is waited for
    for (int i = 0; i < n; ++i) {
        pimc_handle *h = hs[i];
        CK(h, cudaSetDevice(h->device));
        h->comm = comms[i]; h->nranks = n; h->rank = i;
        CK(h, cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
        CK(h, cudaEventCreateWithFlags(&h->cev, cudaEventDisableTiming));
        if (h->g_u64_n < 4) { int rc = dalloc(h, &h->g_u64, 4); if (rc) return rc; h->g_u64_n = 4; }
    }
    long long tot = 0; for (int i = 0; i < n; ++i) tot += hs[i]->S.C;
    for (int i = 0; i < n; ++i) { hs[i]->chains_total = tot; for (int k = 0; k < hs[i]->nen; ++k) hs[i]->g_en_upto[k] = 0; }
    return PIMC_OK;
}
extern "C" int pimc_comm_info(pimc_handle *h, int32_t *nranks, int32_t *rank, int64_t *chains_total, int32_t *nccl_version)
{
    if (!h) return PIMC_ERR_INVALID;
    if (nranks) *nranks = h->nranks; if (rank) *rank = h->rank; if (chains_total) *chains_total = h->chains_total;
    if (nccl_version) { int v = 0; if (h->comm && nccl_api()) nccl_api()->GetVersion(&v); *nccl_version = v; }
    return PIMC_OK;
}
// rows [from, to) of Energy object id: chain sums, then all-reduce over the ranks, on the side stream behind everything queued on the run
// stream so far.  Returns immediately: the next block's moves overlap the reduction.
static int comm_reduce_energy(pimc_handle *h, int id, long long from, long long to)
{
    if (!h->comm || to <= from) return PIMC_OK;
    EnDev &E = h->T.en[id];
    if (!h->g_en[id]) { int rc = dalloc(h, &h->g_en[id], (size_t)2 * E.cap); if (rc) return rc; }
    if (to > E.cap) to = E.cap;
    if (to <= from) return PIMC_OK;
    CK(h, cudaEventRecord(h->cev, h->stream));
    CK(h, cudaStreamWaitEvent(h->cstream, h->cev, 0));
    k_energy_chain_sum<<<(int)(to - from), 256, 0, h->cstream>>>(E.E, E.Ev, h->S.C, from, h->g_en[id]); LAUNCHED();
    CK(h, cudaGetLastError());
    NK(h, nccl_api()->AllReduce(h->g_en[id] + 2 * from, h->g_en[id] + 2 * from, (size_t)(2 * (to - from)), ncclDouble, ncclSum, (ncclComm_t)h->comm, h->cstream));
    h->g_en_upto[id] = to;
    return PIMC_OK;
}

static int sync_tables(pimc_handle *h) { CK(h, cudaMemcpyAsync(h->dT, &h->T, sizeof(DevTables), cudaMemcpyHostToDevice, h->stream)); return PIMC_OK; }

extern "C" int pimc_create(const pimc_config *cfg, pimc_handle **out)
{
    if (!cfg || !out) { SETERR((pimc_handle *)nullptr, "null argument"); return PIMC_ERR_INVALID; }
    *out = nullptr;
    if (cfg->dim < 1 || cfg->dim > 2 || cfg->M < 3 || cfg->N < 1 || cfg->chains < 1 || !(cfg->T > 0) || !(cfg->L > 0) || cfg->Ncycle < 1 ||
        cfg->N > 65535 || cfg->M > 16383 || cfg->pot.nang > PIMC_MAX_ANGLES) {
        SETERR((pimc_handle *)nullptr, "invalid config (dim in {1,2}, 3 <= M <= 16383, 1 <= N <= 65535, chains >= 1, T > 0, L > 0)");
        return PIMC_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        SETERR((pimc_handle *)nullptr, "no CUDA device: libpimc_b200 has no CPU fallback");
        return PIMC_ERR_CUDA;
    }
    pimc_handle *h = new (std::nothrow) pimc_handle();
    if (!h) return PIMC_ERR_NOMEM;
    h->cfg = *cfg; h->err[0] = 0; h->stream = 0; h->iter = 0; h->N_MC = 0; h->Nctr = 0; h->nupd = h->nen = h->nde = 0;
    h->ev0 = h->ev1 = nullptr; h->dT = nullptr; h->dstats = nullptr; h->opt_sweep_impl = 0; h->opt_faithful_impl = 0; h->fscr = nullptr; h->dens_out = nullptr; h->dens_out_n = 0; h->mdone = nullptr; h->opt_fuse_energy = 0; h->opt_isweep = 0; h->nw_head = h->nw_next = nullptr; h->queue = nullptr; h->isw_backoff = 0; h->isw_replay_frac = 0.0;
    memset(h->en_count, 0, sizeof h->en_count);
    memset(&h->T, 0, sizeof h->T);
    h->npc = h->nwi = 0; memset(h->pc_ndata, 0, sizeof h->pc_ndata); memset(h->wi_count, 0, sizeof h->wi_count);
    h->nsk = 0; memset(h->sk_ndata, 0, sizeof h->sk_ndata); h->g_f64 = nullptr; h->g_f64_n = 0;
    h->comm = nullptr; h->nranks = 1; h->rank = 0; h->chains_total = cfg->chains; h->cstream = nullptr; h->cev = nullptr; h->g_u64 = nullptr; h->g_u64_n = 0;
    memset(h->g_en, 0, sizeof h->g_en); memset(h->g_en_upto, 0, sizeof h->g_en_upto);
    int rc = PIMC_OK;
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "CUDA error %s (%s)", cudaGetErrorString(e_), #call); pimc_destroy(h); return PIMC_ERR_CUDA; } } while (0)
#define RCC(call) do { rc = (call); if (rc != PIMC_OK) { snprintf(g_err, sizeof g_err, "%s", h->err); pimc_destroy(h); return rc; } } while (0)
    if (cfg->device >= 0) CKC(cudaSetDevice(cfg->device));
    CKC(cudaGetDevice(&h->device));
    DevSys &S = h->S; memset(&S, 0, sizeof S);
    S.dim = cfg->dim; S.M = cfg->M; S.N = cfg->N; S.C = cfg->chains; S.chain_offset = cfg->chain_offset;
    S.lambda = cfg->lambda; S.L = cfg->L; S.mu = cfg->mu; S.beta = 1.0 / cfg->T; S.tau = S.beta / cfg->M;
    S.a = cfg->interactions ? exp(-2 * M_PI / cfg->g) : 0.0; // system.jl:151
    S.interactions = cfg->interactions; S.compat = cfg->compat; S.ctr = 10000; S.seed = cfg->seed;
    h->vol = pow(2 * cfg->L, cfg->dim);
    pot_to_dev(&cfg->pot, &S.pot);
    h->r_a = cfg->r_a;
    if (h->r_a == 0.0) {
        if (!cfg->interactions) h->r_a = cfg->L / 4; // system.jl:22-24
        else {   // init_int (system.jl:29-31): the cut-off comes from the propagator itself, determine_nnrange(propint, tau, 1e-20, L)
            if (!cfg->tab || cfg->tab_n < 2) { snprintf(g_err, sizeof g_err, "interactions = true needs the pair-propagator table (pimc_build_prop_table)"); pimc_destroy(h); return PIMC_ERR_INVALID; }
            if (pimc_determine_nnrange(cfg->tab, cfg->tab_n, cfg->tab_lo, cfg->tab_hi, S.tau, 1e-20, cfg->L, &h->r_a) != PIMC_OK) {
                snprintf(g_err, sizeof g_err, "determine_nnrange: no sign change of propint - 0.999 on (r_min, L) (Roots.find_zero would throw, system.jl:14)"); pimc_destroy(h); return PIMC_ERR_STATE; }
        }
    }
    S.nbins = (int)floor((2 * cfg->L) / h->r_a); if (S.nbins < 1) S.nbins = 1; // system.jl:81
    S.ncell = cfg->dim == 2 ? S.nbins * S.nbins : S.nbins;
    S.cellw = 2 * cfg->L / S.nbins;
    S.need_cells = (S.a > 0.0 || cfg->interactions) ? 1 : 0;
    size_t nb = (size_t)S.C * S.N * S.M;
    RCC(dalloc(h, &S.r, nb * S.dim)); RCC(dalloc(h, &S.Vl, nb)); RCC(dalloc(h, &S.next, (size_t)S.C * S.N));
    RCC(dalloc(h, &S.prop, nb * S.dim)); RCC(dalloc(h, &S.propV, nb)); RCC(dalloc(h, &S.wtab, (size_t)S.C * S.N));
    if (S.need_cells) {
        RCC(dalloc(h, &S.bins, nb)); RCC(dalloc(h, &S.cell_head, (size_t)S.C * S.M * S.ncell)); RCC(dalloc(h, &S.cell_next, (size_t)S.C * S.M * S.N));
        RCC(dalloc(h, &S.mult, nb));
    }
    if (cfg->tab && cfg->tab_n > 1) {
        double *t; RCC(dalloc(h, &t, (size_t)cfg->tab_n * cfg->tab_n));
        CKC(cudaMemcpy(t, cfg->tab, sizeof(double) * cfg->tab_n * cfg->tab_n, cudaMemcpyHostToDevice));
        S.tab = t; S.tab_n = cfg->tab_n; S.tab_lo = cfg->tab_lo; S.tab_hi = cfg->tab_hi;
    }
    RCC(dalloc(h, &h->dT, 1)); RCC(dalloc(h, &h->dstats, 32)); RCC(dalloc(h, &h->dS, 1));
    {   // table of the division-free log of the Gaussian transform (include/pimc_rng.h), filled with IEEE operations on the host
        double htab[2 * PIMC_LOGTAB_N]; pimc_logtab_fill(htab);
        double *dl; RCC(dalloc(h, &dl, 2 * PIMC_LOGTAB_N));
        CKC(cudaMemcpy(dl, htab, sizeof htab, cudaMemcpyHostToDevice));
        S.logtab = dl;
    }
    {   // staging tables: alpha_k = (k-1)/k, sigma_k = sqrt(((2 lambda) alpha_k) tau), k = 2..M (same IEEE operations as levy!)
        std::vector<double> ta(S.M + 1, 0.0), ts(S.M + 1, 0.0);
        for (int k = 2; k <= S.M; ++k) { volatile double al = (double)(k - 1) / (double)k; volatile double v = 2 * S.lambda; v = v * al; v = v * S.tau; ta[k] = al; ts[k] = sqrt((double)v); }
        double *da, *ds; RCC(dalloc(h, &da, ta.size())); RCC(dalloc(h, &ds, ts.size()));
        CKC(cudaMemcpy(da, ta.data(), ta.size() * sizeof(double), cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(ds, ts.data(), ts.size() * sizeof(double), cudaMemcpyHostToDevice));
        S.tab_alpha = da; S.tab_sig = ds;
    }
    CKC(cudaEventCreate(&h->ev0)); CKC(cudaEventCreate(&h->ev1));
    {   // identity permutation
        std::vector<int> nx((size_t)S.C * S.N);
        for (int c = 0; c < S.C; ++c) for (int n = 0; n < S.N; ++n) nx[(size_t)c * S.N + n] = n;
        CKC(cudaMemcpy(S.next, nx.data(), nx.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (cfg->init) { k_init_world<<<S.C, 128>>>(S); LAUNCHED(); CKC(cudaGetLastError()); }
    else { k_relink<<<grid_for(nb, 256), 256>>>(S, 0, S.C); LAUNCHED(); CKC(cudaGetLastError()); }
    if (S.need_cells) { k_cells_build<<<grid_for((size_t)S.C * S.M, 128), 128>>>(S, 0, S.C); LAUNCHED(); CKC(cudaGetLastError()); }
    CKC(cudaDeviceSynchronize());
    *out = h;
    return PIMC_OK;
}

extern "C" int pimc_set_stream(pimc_handle *h, void *s) { if (!h) return PIMC_ERR_INVALID; h->stream = (cudaStream_t)s; return PIMC_OK; }
extern "C" int pimc_set_option(pimc_handle *h, int32_t option, int64_t value)
{
    if (!h) return PIMC_ERR_INVALID;
    if (option == PIMC_OPT_SWEEP_IMPL && value >= 0 && value <= 3) { h->opt_sweep_impl = (int)value; return PIMC_OK; }
    if (option == PIMC_OPT_FAITHFUL_IMPL && value >= 0 && value <= 1) { h->opt_faithful_impl = (int)value; return PIMC_OK; }
    if (option == PIMC_OPT_FUSE_ENERGY && value >= 0 && value <= 1) { h->opt_fuse_energy = (int)value; return PIMC_OK; }
    if (option == PIMC_OPT_ISWEEP && value >= 0 && value <= 2) { h->opt_isweep = (int)value; h->isw_backoff = 0; return PIMC_OK; }
    SETERR(h, "unknown option %d / value %lld", option, (long long)value); return PIMC_ERR_INVALID;
}
extern "C" int pimc_set_iter(pimc_handle *h, uint64_t iter) { if (!h) return PIMC_ERR_INVALID; h->iter = iter; return PIMC_OK; }
extern "C" int pimc_get_scalars(pimc_handle *h, double *o, int64_t *io)
{
    if (!h) return PIMC_ERR_INVALID;
    o[0] = h->S.beta; o[1] = h->S.tau; o[2] = h->vol; o[3] = h->S.a; o[4] = h->r_a;
    io[0] = h->S.nbins; io[1] = h->N_MC; io[2] = h->Nctr; io[3] = h->S.ctr; io[4] = (int64_t)h->iter;
    return PIMC_OK;
}
static int check_range(pimc_handle *h, int c0, int nc)
{
    if (!h) return PIMC_ERR_INVALID;
    if (c0 < 0 || nc < 0 || c0 + nc > h->S.C) { SETERR(h, "chain range [%d, %d) outside [0, %d)", c0, c0 + nc, h->S.C); return PIMC_ERR_INVALID; }
    return PIMC_OK;
}
extern "C" int pimc_get_paths(pimc_handle *h, int32_t c0, int32_t nc, double *r, double *V, int64_t *bins, int64_t *next)
{
    int rc = check_range(h, c0, nc); if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    DevSys &S = h->S; size_t per = (size_t)S.N * S.M;
    CK(h, cudaStreamSynchronize(h->stream));
    if (r) CK(h, cudaMemcpy(r, S.r + (size_t)c0 * per * S.dim, sizeof(double) * nc * per * S.dim, cudaMemcpyDeviceToHost));
    if (V) CK(h, cudaMemcpy(V, S.Vl + (size_t)c0 * per, sizeof(double) * nc * per, cudaMemcpyDeviceToHost));
    if (bins) {
        long long *tmp; CK(h, cudaMalloc(&tmp, sizeof(long long) * nc * per));
        k_bins_export<<<grid_for(nc * per, 256), 256, 0, h->stream>>>(S, c0, nc, tmp); LAUNCHED();
        cudaError_t e = cudaStreamSynchronize(h->stream);   // h->stream may be non-blocking: the copy below runs on the legacy stream
        if (e == cudaSuccess) e = cudaMemcpy(bins, tmp, sizeof(long long) * nc * per, cudaMemcpyDeviceToHost);
        cudaFree(tmp); CK(h, e);
    }
    if (next) {
        std::vector<int> nx((size_t)nc * S.N);
        CK(h, cudaMemcpy(nx.data(), S.next + (size_t)c0 * S.N, sizeof(int) * nx.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nx.size(); ++i) next[i] = nx[i] + 1;
    }
    return PIMC_OK;
}
extern "C" int pimc_set_paths(pimc_handle *h, int32_t c0, int32_t nc, const double *r, const int64_t *next)
{
    int rc = check_range(h, c0, nc); if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    DevSys &S = h->S; size_t per = (size_t)S.N * S.M;
    if (r) CK(h, cudaMemcpyAsync(S.r + (size_t)c0 * per * S.dim, r, sizeof(double) * nc * per * S.dim, cudaMemcpyHostToDevice, h->stream));
    if (next) {
        std::vector<int> nx((size_t)nc * S.N);
        for (size_t i = 0; i < nx.size(); ++i) {
            if (next[i] < 1 || next[i] > S.N) { SETERR(h, "next[%zu] = %lld outside 1..N", i, (long long)next[i]); return PIMC_ERR_INVALID; }
            nx[i] = (int)next[i] - 1;
        }
        CK(h, cudaMemcpy(S.next + (size_t)c0 * S.N, nx.data(), sizeof(int) * nx.size(), cudaMemcpyHostToDevice));
    }
    k_relink<<<grid_for(nc * per, 256), 256, 0, h->stream>>>(S, c0, nc); LAUNCHED();
    if (S.need_cells) { k_cells_build<<<grid_for((size_t)nc * S.M, 128), 128, 0, h->stream>>>(S, c0, nc); LAUNCHED(); }
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    return PIMC_OK;
}
extern "C" int pimc_update_nnbins(pimc_handle *h)
{
    if (!h) return PIMC_ERR_INVALID;
    if (!h->S.need_cells) return PIMC_OK;
    CK(h, cudaSetDevice(h->device));
    k_cells_build<<<grid_for((size_t)h->S.C * h->S.M, 128), 128, 0, h->stream>>>(h->S, 0, h->S.C); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    return PIMC_OK;
}

// ---- stateless device hooks ----
struct TmpBuf {
    std::vector<void *> v;
    ~TmpBuf() { for (void *p : v) cudaFree(p); }
    template <typename Tp> Tp *up(const Tp *src, size_t n) { void *q = nullptr; if (cudaMalloc(&q, n * sizeof(Tp) + 8) != cudaSuccess) return nullptr; v.push_back(q); if (src) cudaMemcpy(q, src, n * sizeof(Tp), cudaMemcpyHostToDevice); else cudaMemset(q, 0, n * sizeof(Tp)); return (Tp *)q; }
};
#define CKG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "CUDA error %s (%s)", cudaGetErrorString(e_), #call); return PIMC_ERR_CUDA; } } while (0)
#define NEEDGPU() do { int nd_ = 0; if (cudaGetDeviceCount(&nd_) != cudaSuccess || nd_ == 0) { snprintf(g_err, sizeof g_err, "no CUDA device: libpimc_b200 has no CPU fallback"); return PIMC_ERR_CUDA; } } while (0)

extern "C" int pimc_distance(int64_t n, const double *x1, const double *x2, double L, double *out)
{
    NEEDGPU(); TmpBuf t; double *a = t.up(x1, n), *b = t.up(x2, n), *o = t.up((double *)nullptr, n);
    if (!a || !b || !o) return PIMC_ERR_NOMEM;
    k_distance<<<grid_for(n, 256), 256>>>(n, a, b, L, o); LAUNCHED(); CKG(cudaGetLastError());
    CKG(cudaMemcpy(out, o, n * sizeof(double), cudaMemcpyDeviceToHost)); return PIMC_OK;
}
extern "C" int pimc_teleport(int64_t n, const double *x, double L, double *out)
{
    NEEDGPU(); TmpBuf t; double *a = t.up(x, n), *o = t.up((double *)nullptr, n);
    if (!a || !o) return PIMC_ERR_NOMEM;
    k_teleport<<<grid_for(n, 256), 256>>>(n, a, L, o); LAUNCHED(); CKG(cudaGetLastError());
    CKG(cudaMemcpy(out, o, n * sizeof(double), cudaMemcpyDeviceToHost)); return PIMC_OK;
}
extern "C" int pimc_lnK(int64_t n, const double *r1, const double *r2, int32_t dim, double tau, double lambda, double L, double *out)
{
    NEEDGPU(); TmpBuf t; double *a = t.up(r1, n * dim), *b = t.up(r2, n * dim), *o = t.up((double *)nullptr, n);
    if (!a || !b || !o) return PIMC_ERR_NOMEM;
    k_lnK<<<grid_for(n, 256), 256>>>(n, a, b, dim, tau, lambda, L, o); CKG(cudaGetLastError());
    CKG(cudaMemcpy(out, o, n * sizeof(double), cudaMemcpyDeviceToHost)); return PIMC_OK;
}
extern "C" int pimc_lnV(int64_t n, const double *r1, const double *r2, int32_t dim, double tau, const pimc_potential *p, double *out)
{
    NEEDGPU(); TmpBuf t; double *a = t.up(r1, n * dim), *b = t.up(r2, n * dim), *o = t.up((double *)nullptr, n);
    if (!a || !b || !o) return PIMC_ERR_NOMEM;
    PotDev pd; pot_to_dev(p, &pd);
    k_lnV<<<grid_for(n, 256), 256>>>(n, a, b, dim, tau, pd, o); CKG(cudaGetLastError());
    CKG(cudaMemcpy(out, o, n * sizeof(double), cudaMemcpyDeviceToHost)); return PIMC_OK;
}
extern "C" int pimc_potential_eval(int64_t n, const double *r, int32_t dim, const pimc_potential *p, double *V, double *dV)
{
    NEEDGPU(); TmpBuf t; double *a = t.up(r, n * dim), *o = t.up((double *)nullptr, n), *g = t.up((double *)nullptr, n * dim);
    if (!a || !o || !g) return PIMC_ERR_NOMEM;
    PotDev pd; pot_to_dev(p, &pd);
    k_pot<<<grid_for(n, 256), 256>>>(n, a, dim, pd, o, g); LAUNCHED(); CKG(cudaGetLastError());
    if (V) CKG(cudaMemcpy(V, o, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (dV) CKG(cudaMemcpy(dV, g, n * dim * sizeof(double), cudaMemcpyDeviceToHost));
    CKG(cudaDeviceSynchronize()); return PIMC_OK;
}
extern "C" int pimc_levy_bridge(double *r, int32_t rows, int32_t dim, double tau, double L, double lambda, const double *xi, int64_t nb)
{
    NEEDGPU();
    if (rows < 2 || dim < 1 || dim > 2 || nb < 1) { snprintf(g_err, sizeof g_err, "levy_bridge: rows >= 2, dim in {1,2}, nb >= 1"); return PIMC_ERR_INVALID; }
    TmpBuf t; double *a = t.up(r, (size_t)nb * rows * dim), *x = t.up(xi, (size_t)nb * (rows - 2) * dim + 1);
    if (!a || !x) return PIMC_ERR_NOMEM;
    k_levy<<<grid_for(nb, 128), 128>>>(a, rows, dim, tau, L, lambda, x, nb); LAUNCHED(); CKG(cudaGetLastError());
    CKG(cudaMemcpy(r, a, (size_t)nb * rows * dim * sizeof(double), cudaMemcpyDeviceToHost)); return PIMC_OK;
}
extern "C" int pimc_gauss_pairs(uint64_t seed, uint32_t chain, uint64_t iter, uint32_t slot, uint32_t kind, uint32_t retry, uint32_t bead0, int64_t n, double *g)
{
    NEEDGPU(); TmpBuf t; double *o = t.up((double *)nullptr, 2 * n);
    double htab[2 * PIMC_LOGTAB_N]; pimc_logtab_fill(htab);
    double *dtab = t.up(htab, 2 * PIMC_LOGTAB_N);
    if (!o || !dtab) return PIMC_ERR_NOMEM;
    k_gauss<<<grid_for(n, 256), 256>>>(seed, chain, iter, slot, kind, retry, bead0, n, o, dtab); LAUNCHED(); CKG(cudaGetLastError());
    CKG(cudaMemcpy(g, o, 2 * n * sizeof(double), cudaMemcpyDeviceToHost)); return PIMC_OK;
}

// ---- estimators / action ----
extern "C" int pimc_energy_now(pimc_handle *h, double *E, double *Ev, double *parts)
{
    if (!h) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; int C = h->S.C; double *e = t.up((double *)nullptr, C), *ev = t.up((double *)nullptr, C), *p = t.up((double *)nullptr, 3 * (size_t)C);
    if (!e || !ev || !p) return PIMC_ERR_NOMEM;
    k_energy_now<<<C, 256, 0, h->stream>>>(h->S, e, ev, p); LAUNCHED(); CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    if (E) CK(h, cudaMemcpy(E, e, C * sizeof(double), cudaMemcpyDeviceToHost));
    if (Ev) CK(h, cudaMemcpy(Ev, ev, C * sizeof(double), cudaMemcpyDeviceToHost));
    if (parts) CK(h, cudaMemcpy(parts, p, 3 * (size_t)C * sizeof(double), cudaMemcpyDeviceToHost));
    return PIMC_OK;
}
extern "C" int pimc_action(pimc_handle *h, double *cached, double *recomputed)
{
    if (!h) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; int C = h->S.C; double *a = t.up((double *)nullptr, C), *b = t.up((double *)nullptr, C);
    if (!a || !b) return PIMC_ERR_NOMEM;
    k_action<<<C, 256, 0, h->stream>>>(h->S, a, b); LAUNCHED(); CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    if (cached) CK(h, cudaMemcpy(cached, a, C * sizeof(double), cudaMemcpyDeviceToHost));
    if (recomputed) CK(h, cudaMemcpy(recomputed, b, C * sizeof(double), cudaMemcpyDeviceToHost));
    return PIMC_OK;
}
extern "C" int pimc_find_nn(pimc_handle *h, int32_t chain, const double *r, int64_t slice, int64_t exception, int64_t *nn)
{
    int rc = check_range(h, chain, 1); if (rc) return rc;
    if (!h->S.need_cells) { SETERR(h, "cell list not built (a == 0 and no interactions)"); return PIMC_ERR_STATE; }
    if (slice < 1 || slice > h->S.M) { SETERR(h, "slice outside 1..M"); return PIMC_ERR_INVALID; }
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; long long *o = t.up((long long *)nullptr, 1); if (!o) return PIMC_ERR_NOMEM;
    k_find_nn<<<1, 32, 0, h->stream>>>(h->S, chain, r[0], h->S.dim > 1 ? r[1] : 0.0, (int)slice - 1, (int)exception - 1, o); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    long long v; CK(h, cudaMemcpy(&v, o, sizeof v, cudaMemcpyDeviceToHost)); *nn = v; return PIMC_OK;
}
extern "C" int pimc_find_nns(pimc_handle *h, int32_t chain, const double *r, int64_t slice, int64_t exception, int64_t *out, int64_t cap, int64_t *count)
{
    int rc = check_range(h, chain, 1); if (rc) return rc;
    if (!h->S.need_cells) { SETERR(h, "cell list not built (a == 0 and no interactions)"); return PIMC_ERR_STATE; }
    if (slice < 1 || slice > h->S.M) { SETERR(h, "slice outside 1..M"); return PIMC_ERR_INVALID; }
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; long long *o = t.up((long long *)nullptr, cap + 1), *cn = t.up((long long *)nullptr, 1); if (!o || !cn) return PIMC_ERR_NOMEM;
    k_find_nns<<<1, 32, 0, h->stream>>>(h->S, chain, r[0], h->S.dim > 1 ? r[1] : 0.0, (int)slice - 1, (int)exception - 1, o, cap, cn); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    long long v; CK(h, cudaMemcpy(&v, cn, sizeof v, cudaMemcpyDeviceToHost)); *count = v;
    long long ncopy = v < cap ? v : cap;
    if (ncopy > 0) CK(h, cudaMemcpy(out, o, ncopy * sizeof(long long), cudaMemcpyDeviceToHost));
    return PIMC_OK;
}

// ---- explicit moves ----
static int check_move(pimc_handle *h, int chain, int64_t n, int64_t j0, int64_t m)
{
    int rc = check_range(h, chain, 1); if (rc) return rc;
    if (n < 1 || n > h->S.N || j0 < 1 || j0 > h->S.M || m < 2 || m > h->S.M - 2) { SETERR(h, "move arguments out of range (1<=n<=N, 1<=j0<=M, 2<=m<=M-2)"); return PIMC_ERR_INVALID; }
    return PIMC_OK;
}
extern "C" int pimc_reshape_linear_explicit(pimc_handle *h, int32_t chain, int64_t n, int64_t j0, int64_t m, const double *xi, double u,
                                            int32_t commit, double *w_initial, double *w_updated, double *rprime, int32_t *acc)
{
    int rc = check_move(h, chain, n, j0, m); if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; double *x = t.up(xi, (size_t)(m - 1) * h->S.dim + 1), *rp = t.up((double *)nullptr, (size_t)(m + 1) * h->S.dim);
    MoveOut *o = t.up((MoveOut *)nullptr, 1); if (!x || !rp || !o) return PIMC_ERR_NOMEM;
    k_reshape_linear_explicit<<<1, 32, 0, h->stream>>>(h->S, chain, (int)n - 1, (int)j0, (int)m, x, u, commit, o, rp); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    MoveOut mo; CK(h, cudaMemcpy(&mo, o, sizeof mo, cudaMemcpyDeviceToHost));
    if (w_initial) *w_initial = mo.wi; if (w_updated) *w_updated = mo.wu; if (acc) *acc = mo.acc;
    if (rprime) CK(h, cudaMemcpy(rprime, rp, sizeof(double) * (m + 1) * h->S.dim, cudaMemcpyDeviceToHost));
    return PIMC_OK;
}
extern "C" int pimc_reshape_swap_explicit(pimc_handle *h, int32_t chain, int64_t n1, int64_t n2, int64_t j0, int64_t m, const double *xi1,
                                          const double *xi2, double u, int32_t commit, double *w_initial, double *w_updated, int32_t *acc)
{
    int rc = check_move(h, chain, n1, j0, m); if (rc) return rc;
    if (n2 < 1 || n2 > h->S.N || h->S.N < 2) { SETERR(h, "swap needs two particles in 1..N"); return PIMC_ERR_INVALID; }
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; size_t nx = (size_t)(m - 1) * h->S.dim + 1; double *x1 = t.up(xi1, nx), *x2 = t.up(xi2, nx);
    MoveOut *o = t.up((MoveOut *)nullptr, 1); if (!x1 || !x2 || !o) return PIMC_ERR_NOMEM;
    k_reshape_swap_explicit<<<1, 32, 0, h->stream>>>(h->S, chain, (int)n1 - 1, (int)n2 - 1, (int)j0, (int)m, x1, x2, u, commit, o); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    MoveOut mo; CK(h, cudaMemcpy(&mo, o, sizeof mo, cudaMemcpyDeviceToHost));
    if (w_initial) *w_initial = mo.wi; if (w_updated) *w_updated = mo.wu; if (acc) *acc = mo.acc;
    return PIMC_OK;
}
extern "C" int pimc_com_explicit(pimc_handle *h, int32_t chain, int64_t n, int32_t polymer, const double *d, double u, int32_t commit,
                                 double *w_initial, double *w_updated, int32_t *acc)
{
    int rc = check_range(h, chain, 1); if (rc) return rc;
    if (n < 1 || n > h->S.N) { SETERR(h, "n outside 1..N"); return PIMC_ERR_INVALID; }
    (void)polymer; // the cycle of n is moved as a whole in both variants (SingleCenterOfMass only picks n with next == n)
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; double *dd = t.up(d, 2); MoveOut *o = t.up((MoveOut *)nullptr, 1); if (!dd || !o) return PIMC_ERR_NOMEM;
    k_com_explicit<<<1, 32, 0, h->stream>>>(h->S, chain, (int)n - 1, dd, u, commit, o); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    MoveOut mo; CK(h, cudaMemcpy(&mo, o, sizeof mo, cudaMemcpyDeviceToHost));
    if (w_initial) *w_initial = mo.wi; if (w_updated) *w_updated = mo.wu; if (acc) *acc = mo.acc;
    return PIMC_OK;
}
extern "C" int pimc_swap_weights(pimc_handle *h, int32_t chain, int64_t n1, int64_t j0, int64_t m, double *w)
{
    int rc = check_move(h, chain, n1, j0, m); if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    TmpBuf t; double *o = t.up((double *)nullptr, h->S.N); if (!o) return PIMC_ERR_NOMEM;
    k_swap_weights<<<1, 32, 0, h->stream>>>(h->S, chain, (int)n1 - 1, (int)j0, (int)m, o); LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(w, o, sizeof(double) * h->S.N, cudaMemcpyDeviceToHost)); return PIMC_OK;
}

// ---- update objects ----
static int alloc_ring(pimc_handle *h, UpdDev &U)
{
    U.ring_words = (int)((U.range + 1 + 31) / 32);
    return dalloc(h, &U.ring, (size_t)h->S.C * U.ring_words);
}
extern "C" int pimc_update_create(pimc_handle *h, int32_t kind, double var0, int32_t *id)
{
    if (!h || !id) return PIMC_ERR_INVALID;
    if (kind < 0 || kind > 3) { SETERR(h, "unknown update kind %d", kind); return PIMC_ERR_INVALID; }
    if (h->nupd >= PIMC_MAXU) { SETERR(h, "at most %d update objects per handle", PIMC_MAXU); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    UpdDev &U = h->T.upd[h->nupd]; memset(&U, 0, sizeof U);
    U.kind = kind; U.adj = 10; U.range = 10000;
    double v0;
    if (kind == PIMC_UPD_RESHAPE_LINEAR || kind == PIMC_UPD_RESHAPE_SWAP) { // reshape.jl:12-28,104-120
        U.vmin = 2; U.vmax = h->S.M - 2; U.minacc = 0.6; U.maxacc = 0.8;
        v0 = floor(var0) < U.vmax ? floor(var0) : U.vmax;
    } else { // com.jl:12-27,117-132
        U.vmin = 1e-1; U.vmax = h->S.L / 2; U.minacc = 0.4; U.maxacc = 0.6; v0 = var0;
    }
    int rc; size_t C = h->S.C;
    if ((rc = dalloc(h, &U.var, C)) || (rc = dalloc(h, &U.tries, C)) || (rc = dalloc(h, &U.accepted, C)) || (rc = dalloc(h, &U.tries_var, C)) ||
        (rc = dalloc(h, &U.bead_moves, C)) || (rc = dalloc(h, &U.ring_head, C)) || (rc = dalloc(h, &U.ring_len, C)) || (rc = dalloc(h, &U.ring_sum, C)) ||
        (rc = alloc_ring(h, U))) return rc;
    std::vector<double> v(C, v0);
    CK(h, cudaMemcpy(U.var, v.data(), C * sizeof(double), cudaMemcpyHostToDevice));
    *id = h->nupd++;
    return sync_tables(h);
}
extern "C" int pimc_update_configure(pimc_handle *h, int32_t id, double vmin, double vmax, double minacc, double maxacc, int64_t adj, int64_t range)
{
    if (!h || id < 0 || id >= h->nupd) return PIMC_ERR_INVALID;
    if (adj < 1 || range < 1) { SETERR(h, "adj and range must be >= 1"); return PIMC_ERR_INVALID; }
    CK(h, cudaSetDevice(h->device));
    UpdDev &U = h->T.upd[id]; size_t C = h->S.C;
    U.vmin = vmin; U.vmax = vmax; U.minacc = minacc; U.maxacc = maxacc; U.adj = adj;
    if (range != U.range) { U.range = range; int rc = alloc_ring(h, U); if (rc) return rc;
        CK(h, cudaMemset(U.ring_head, 0, C * sizeof(int))); CK(h, cudaMemset(U.ring_len, 0, C * sizeof(int))); CK(h, cudaMemset(U.ring_sum, 0, C * sizeof(int))); }
    if (U.kind == PIMC_UPD_RESHAPE_LINEAR || U.kind == PIMC_UPD_RESHAPE_SWAP) { // NumbOfSlices(min(slices, maxslices), ...)
        std::vector<double> v(C); CK(h, cudaMemcpy(v.data(), U.var, C * sizeof(double), cudaMemcpyDeviceToHost));
        for (auto &x : v) if (x > vmax) x = vmax;
        CK(h, cudaMemcpy(U.var, v.data(), C * sizeof(double), cudaMemcpyHostToDevice));
    }
    return sync_tables(h);
}
extern "C" int pimc_update_get(pimc_handle *h, int32_t id, int32_t chain, double *var, int64_t *tries, int64_t *tries_var,
                               double *acc_window, int64_t *accepted, int64_t *bead_moves)
{
    if (!h || id < 0 || id >= h->nupd || chain < -1 || chain >= h->S.C) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    UpdDev &U = h->T.upd[id]; size_t C = h->S.C;
    std::vector<double> v(C); std::vector<long long> tr(C), ac(C), tv(C), bm(C); std::vector<int> rl(C), rs(C);
    CK(h, cudaMemcpy(v.data(), U.var, C * 8, cudaMemcpyDeviceToHost)); CK(h, cudaMemcpy(tr.data(), U.tries, C * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(ac.data(), U.accepted, C * 8, cudaMemcpyDeviceToHost)); CK(h, cudaMemcpy(tv.data(), U.tries_var, C * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(bm.data(), U.bead_moves, C * 8, cudaMemcpyDeviceToHost)); CK(h, cudaMemcpy(rl.data(), U.ring_len, C * 4, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(rs.data(), U.ring_sum, C * 4, cudaMemcpyDeviceToHost));
    if (chain >= 0) {
        if (var) *var = v[chain]; if (tries) *tries = tr[chain]; if (tries_var) *tries_var = tv[chain];
        if (acc_window) *acc_window = (double)rs[chain] / (double)rl[chain];
        if (accepted) *accepted = ac[chain]; if (bead_moves) *bead_moves = bm[chain];
    } else {
        double sv = 0, sa = 0; long long t1 = 0, t2 = 0, t3 = 0, t4 = 0; size_t na = 0;
        for (size_t c = 0; c < C; ++c) { sv += v[c]; if (rl[c] > 0) { sa += (double)rs[c] / rl[c]; na++; } t1 += tr[c]; t2 += tv[c]; t3 += ac[c]; t4 += bm[c]; }
        if (var) *var = sv / C; if (acc_window) *acc_window = na ? sa / na : NAN;
        if (tries) *tries = t1; if (tries_var) *tries_var = t2; if (accepted) *accepted = t3; if (bead_moves) *bead_moves = t4;
    }
    return PIMC_OK;
}

// ---- measurement objects ----
extern "C" int pimc_energy_create(pimc_handle *h, int64_t cap, int32_t *id)
{
    if (!h || !id || cap < 1) return PIMC_ERR_INVALID;
    if (h->nen >= PIMC_MAXE) { SETERR(h, "at most %d Energy objects per handle", PIMC_MAXE); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    EnDev &E = h->T.en[h->nen]; E.cap = cap; int rc;
    if ((rc = dalloc(h, &E.E, (size_t)cap * h->S.C)) || (rc = dalloc(h, &E.Ev, (size_t)cap * h->S.C)) || (rc = dalloc(h, &E.acc, (size_t)h->S.C * 5))) return rc;
    h->en_cap[h->nen] = cap;
    *id = h->nen++;
    return sync_tables(h);
}
static int energy_count(pimc_handle *h, int id, long long *n) { *n = h->en_count[id]; return PIMC_OK; }
extern "C" int pimc_energy_read_range(pimc_handle *h, int32_t id, int32_t chain, int64_t start, int64_t count, double *E, double *Ev, int64_t *n)
{
    if (!h || id < 0 || id >= h->nen || chain < -1 || chain >= h->S.C || start < 0 || count < 0) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    long long cnt; int rc = energy_count(h, id, &cnt); if (rc) return rc;
    if (n) *n = cnt;
    long long avail = cnt < h->T.en[id].cap ? cnt : h->T.en[id].cap;
    long long m = avail - start; if (m > count) m = count;
    if (m <= 0 || (!E && !Ev)) return PIMC_OK;
    TmpBuf t; double *a = t.up((double *)nullptr, m), *b = t.up((double *)nullptr, m); if (!a || !b) return PIMC_ERR_NOMEM;
    if (chain < 0 && h->comm) {   // mean over the chains of ALL ranks (collective: every rank reads the same range); blocks reduced at the end of pimc_run are ready
        if (h->g_en_upto[id] < avail) { rc = comm_reduce_energy(h, id, h->g_en_upto[id], avail); if (rc) return rc; }
        CK(h, cudaStreamSynchronize(h->cstream));
        k_energy_global_mean<<<(int)((m + 127) / 128), 128, 0, h->stream>>>(h->g_en[id], start, m, (double)h->chains_total, a, b);
    }
    else if (chain < 0) k_energy_chain_mean<<<(int)m, 256, 0, h->stream>>>(h->T.en[id].E, h->T.en[id].Ev, h->S.C, start, a, b);
    else k_energy_chain_series<<<(int)((m + 127) / 128), 128, 0, h->stream>>>(h->T.en[id].E, h->T.en[id].Ev, h->S.C, chain, start, m, a, b);
    LAUNCHED();
    CK(h, cudaGetLastError()); CK(h, cudaStreamSynchronize(h->stream));
    if (E) CK(h, cudaMemcpy(E, a, m * sizeof(double), cudaMemcpyDeviceToHost));
    if (Ev) CK(h, cudaMemcpy(Ev, b, m * sizeof(double), cudaMemcpyDeviceToHost));
    return PIMC_OK;
}
extern "C" int pimc_energy_read(pimc_handle *h, int32_t id, int32_t chain, double *E, double *Ev, int64_t cap, int64_t *n)
{
    return pimc_energy_read_range(h, id, chain, 0, cap, E, Ev, n);
}
extern "C" int pimc_energy_stats(pimc_handle *h, int32_t id, double *out)
{
    if (!h || id < 0 || id >= h->nen || !out) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(out, h->T.en[id].acc, sizeof(double) * 5 * h->S.C, cudaMemcpyDeviceToHost));
    return PIMC_OK;
}
extern "C" int pimc_density_create(pimc_handle *h, int64_t nbins, int32_t *id)
{
    if (!h || !id || nbins < 1) return PIMC_ERR_INVALID;
    if (h->nde >= PIMC_MAXD) { SETERR(h, "at most %d Density objects per handle", PIMC_MAXD); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    DeDev &D = h->T.de[h->nde]; D.nbins = nbins; D.bin = (2 * h->S.L) / nbins; // measurement.jl:37
    size_t sz = h->S.dim == 2 ? (size_t)nbins * nbins : (size_t)nbins;
    int rc = dalloc(h, &D.dens, sz); if (rc) return rc;
    h->de_ndata[h->nde] = 0;
    *id = h->nde++;
    return sync_tables(h);
}
extern "C" int pimc_density_measure(pimc_handle *h, int32_t id)
{
    if (!h || id < 0 || id >= h->nde) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    k_density_now<<<h->S.C, 256, 0, h->stream>>>(h->S, h->T.de[id]); LAUNCHED(); CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    h->de_ndata[id] += (long long)h->S.M * h->S.C;
    return PIMC_OK;
}
extern "C" int pimc_density_read(pimc_handle *h, int32_t id, double *dens, int64_t *ndata, double *bin)
{
    if (!h || id < 0 || id >= h->nde) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    DeDev &D = h->T.de[id]; size_t sz = h->S.dim == 2 ? (size_t)D.nbins * D.nbins : (size_t)D.nbins;
    if (dens) {
        // read-out buffer kept with the handle: a cudaMalloc / cudaFree pair per block costs a device-wide synchronisation each
        if (h->dens_out_n < sz) { double *o2 = nullptr; int rc = dalloc(h, &o2, sz); if (rc) return rc; h->dens_out = o2; h->dens_out_n = sz; }
        double *o = h->dens_out;
        const unsigned long long *src = D.dens;
        if (h->comm) {   // counters summed over the ranks (collective); the local counters stay local
            if (h->g_u64_n < sz + 1) { int rc = dalloc(h, &h->g_u64, sz + 1); if (rc) return rc; h->g_u64_n = sz + 1; }
            CK(h, cudaMemcpyAsync(h->g_u64, D.dens, sz * 8, cudaMemcpyDeviceToDevice, h->cstream));
            const unsigned long long nd = (unsigned long long)h->de_ndata[id];
            CK(h, cudaMemcpyAsync(h->g_u64 + sz, &nd, 8, cudaMemcpyHostToDevice, h->cstream));
            NK(h, nccl_api()->AllReduce(h->g_u64, h->g_u64, sz + 1, ncclUint64, ncclSum, (ncclComm_t)h->comm, h->cstream));
            CK(h, cudaStreamSynchronize(h->cstream));
            src = h->g_u64;
        }
        k_dens_to_double<<<grid_for(sz, 256), 256, 0, h->stream>>>(src, sz, o); LAUNCHED(); CK(h, cudaGetLastError());
        CK(h, cudaStreamSynchronize(h->stream));
        CK(h, cudaMemcpy(dens, o, sz * sizeof(double), cudaMemcpyDeviceToHost));
        if (h->comm && ndata) { unsigned long long nd; CK(h, cudaMemcpy(&nd, h->g_u64 + sz, 8, cudaMemcpyDeviceToHost)); *ndata = (int64_t)nd; ndata = nullptr; }
    }
    else if (h->comm && ndata) {   // ndata alone: still the global count
        if (h->g_u64_n < 4) { int rc = dalloc(h, &h->g_u64, 4); if (rc) return rc; h->g_u64_n = 4; }
        unsigned long long nd = (unsigned long long)h->de_ndata[id];
        CK(h, cudaMemcpyAsync(h->g_u64, &nd, 8, cudaMemcpyHostToDevice, h->cstream));
        NK(h, nccl_api()->AllReduce(h->g_u64, h->g_u64, 1, ncclUint64, ncclSum, (ncclComm_t)h->comm, h->cstream));
        CK(h, cudaMemcpyAsync(&nd, h->g_u64, 8, cudaMemcpyDeviceToHost, h->cstream));
        CK(h, cudaStreamSynchronize(h->cstream));
        *ndata = (int64_t)nd; ndata = nullptr;
    }
    if (ndata) *ndata = h->de_ndata[id];
    if (bin) *bin = D.bin;
    return PIMC_OK;
}

// ---- run! ----
// force_cadence: the measurement cadence (N_MC / Nctr, simulation.jl:31-32 retry cap) runs although neither Energy nor Density is listed --
// pimc_run_ex drives the remaining estimators between segments
static int run_core(pimc_handle *h, int64_t n, const int32_t *update_ids, const int64_t *every, int32_t nupd,
                    const int32_t *energy_ids, int32_t nen, const int32_t *density_ids, int32_t nde, int32_t sched, bool force_cadence, pimc_run_stats *stats)
{
    if (!h) return PIMC_ERR_INVALID;
    if (n < 0 || nupd < 1 || nupd > PIMC_MAXU || nen < 0 || nen > PIMC_MAXE || nde < 0 || nde > PIMC_MAXD || !update_ids || !every) { SETERR(h, "pimc_run: bad arguments"); return PIMC_ERR_INVALID; }
    if (sched != PIMC_SCHED_FAITHFUL && sched != PIMC_SCHED_SWEEP) { SETERR(h, "unknown schedule %d", sched); return PIMC_ERR_INVALID; }
    DevSys &S = h->S;
    // sweep of interacting worldlines: sequential per chain inside the persistent kernel (warp- / CTA-cooperative bodies only)
    if (sched == PIMC_SCHED_SWEEP && S.need_cells && h->opt_faithful_impl != 0) { SETERR(h, "the sweep schedule of interacting worldlines needs the cooperative proposals (PIMC_OPT_FAITHFUL_IMPL = 0)"); return PIMC_ERR_UNSUPPORTED; }
    CK(h, cudaSetDevice(h->device));
    RunParams P; memset(&P, 0, sizeof P);
    P.n = n; P.iter0 = h->iter; P.nupd = nupd; P.nen = nen; P.nde = nde; P.sched = sched;
    for (int i = 0; i < nupd; ++i) {
        if (update_ids[i] < 0 || update_ids[i] >= h->nupd || every[i] < 1) { SETERR(h, "bad update id / every"); return PIMC_ERR_INVALID; }
        P.upd_id[i] = update_ids[i]; P.w[i] = 1.0 / (double)every[i];
    }
    for (int i = 0; i < nen; ++i) {
        if (energy_ids[i] < 0 || energy_ids[i] >= h->nen) { SETERR(h, "bad energy id"); return PIMC_ERR_INVALID; }
        for (int k = 0; k < i; ++k) if (energy_ids[k] == energy_ids[i]) { SETERR(h, "Energy object %d listed twice", energy_ids[i]); return PIMC_ERR_INVALID; }
        P.en_id[i] = energy_ids[i]; P.en_k0[i] = h->en_count[energy_ids[i]];
    }
    for (int i = 0; i < nde; ++i) { if (density_ids[i] < 0 || density_ids[i] >= h->nde) { SETERR(h, "bad density id"); return PIMC_ERR_INVALID; } P.de_id[i] = density_ids[i]; }
    P.Nctr0 = h->Nctr; P.N_MC0 = h->N_MC; P.Ncycle = h->cfg.Ncycle; P.stats = h->dstats;
    const bool cadence = force_cadence || nen + nde > 0;
    S.ctr = cadence ? 1000 : 10000; // simulation.jl:31-32
    // energy overflow: the reference errors when the pre-sized vector is full (measurement.jl:119-120); every Energy object counts its own samples
    const long long nmeas = cadence ? (h->Nctr + n) / h->cfg.Ncycle : 0;
    for (int i = 0; i < nen; ++i) {
        const long long cnt = h->en_count[P.en_id[i]];
        if (cnt + nmeas > h->T.en[P.en_id[i]].cap) { SETERR(h, "Energy buffer of %lld entries would overflow (%lld + %lld)", (long long)h->T.en[P.en_id[i]].cap, cnt, nmeas); return PIMC_ERR_STATE; }
    }
    CK(h, cudaMemsetAsync(h->dstats, 0, 32 * sizeof(unsigned long long), h->stream));
    // FAITHFUL: warp 0 owns the proposal; large chains get three more warps for the estimators (Energy / Density stream N*M beads)
    int threads = sched == PIMC_SCHED_SWEEP ? 64 : ((size_t)S.N * S.M >= 2048 && (S.need_cells || (nen + nde > 0 && S.C <= 1184)) ? (S.need_cells ? PIMC_CELLS_THREADS : 64) : 32);
    if (sched == PIMC_SCHED_SWEEP) { while (threads < S.N && threads < 256) threads *= 2; }
    if (sched == PIMC_SCHED_SWEEP && S.N == 1 && !S.need_cells) threads = 32;   // one-particle systems (the shipped trapped example): one warp per chain, twice the resident chains
    if (S.need_cells && threads > PIMC_CELLS_THREADS) threads = PIMC_CELLS_THREADS;   // launch bound of k_run_cells
    size_t smem = 96 * sizeof(double) + (size_t)S.N + 16;
    P.fimpl = h->opt_faithful_impl; P.fscr = nullptr;
    P.prof = getenv("PIMC_PROF") ? h->dstats + 4 : nullptr;   // per-phase cycle counters of k_run, printed to stderr after the run
    if (P.fimpl == 0) {
        const size_t fs = faithful_scratch_doubles(S.N, S.M) * sizeof(double);
        if (smem + fs <= 160 * 1024) smem += fs;
        else { // rows of one chain do not fit shared memory: scratch in HBM, one slab per CTA
            if (!h->fscr) { int rc = dalloc(h, &h->fscr, (size_t)S.C * faithful_scratch_doubles(S.N, S.M)); if (rc) return rc; }
            P.fscr = h->fscr;
        }
    }
    int launches = 0;
    // per-iteration sweep kernels (pimc_sweep.cuh) for large batches, the persistent kernel otherwise
    const int pk = S.pot.kind;
    // staging half: xs, ys, (pv) and the row map take BCAP rows; BCAP fills what four CTAs per SM leave of the shared memory
    const size_t rs_fixed = (size_t)SWEEP_THREADS * sizeof(double) + 2 * SWEEP_THREADS * sizeof(int) + (((size_t)S.N + 15) & ~(size_t)15) + (size_t)(S.M + 1 + 2 * PIMC_LOGTAB_N) * sizeof(double) + 32;
    const size_t rs_row = (size_t)(pk == PIMC_POT_ZERO ? 2 : 3) * sizeof(double) + 1;
    const size_t rs_budget256 = ((S.M + 31) / 32 <= 4 ? 56000 : 74000);   // launch bounds: four (KM <= 4) or three CTAs of 256 threads per SM
    const size_t rs_budget = SWEEP_THREADS == 256 ? rs_budget256 : rs_budget256 * SWEEP_THREADS / 256 - 1200;   // (smaller CTAs: 1 KiB reserved + static shared memory per CTA weigh more)
    const int bcap = rs_fixed + 512 * rs_row < rs_budget ? (int)(((rs_budget - rs_fixed) / rs_row) & ~(size_t)15) : 512;
    const size_t smem_rs = rs_fixed + (size_t)bcap * rs_row;
    // COM half: flags, then (TMA path, even M) two mbarriers per warp and two stages of (dim + 1) rows per warp
    const size_t com_flag = (((size_t)S.N + 127) & ~(size_t)127);
    const bool com_tma = (S.M % 2) == 0 && getenv("PIMC_NO_TMA") == nullptr;
    bool has_pcom = false;
    for (int i = 0; i < nupd; ++i) has_pcom |= h->T.upd[update_ids[i]].kind == PIMC_UPD_POLYMER_COM;
    auto com_bytes = [&](int threads) {   // flags | mbarriers | max(TMA stages of every warp, scratch of the exchange-cycle moves)
        const size_t stage = com_tma ? (size_t)(threads / 32) * 2 * (S.dim + 1) * S.M * sizeof(double) : 0, cyc = has_pcom ? pcom_smem_bytes(S.N) : 0;
        return com_flag + (size_t)(threads / 32) * 16 + (stage > cyc ? stage : cyc) + 16;
    };
    const size_t smem_cs = com_bytes(SWEEP_THREADS);
    const bool batched_ok = sched == PIMC_SCHED_SWEEP && !S.need_cells && S.M <= 256 && smem_rs <= 200 * 1024 && smem_cs <= 200 * 1024;
    bool batched = batched_ok && (h->opt_sweep_impl >= 2 || (h->opt_sweep_impl == 0 && (size_t)S.C * S.N * S.M >= (size_t)1 << 20));
    if (h->opt_sweep_impl >= 2 && sched == PIMC_SCHED_SWEEP && !batched_ok) { SETERR(h, "per-iteration sweep kernels need independent worldlines and M <= %d", 256); return PIMC_ERR_UNSUPPORTED; }
    // optimistic-parallel sweep of interacting worldlines (pimc_isweep.cuh): the hard core is the only coupling between the proposals of a
    // sweep when the pair action does not enter ReshapeLinear / centre-of-mass moves (the reference as shipped, compat PAIR_BYVALUE)
    const bool pair_in_moves = S.interactions && !(S.compat & PIMC_COMPAT_PAIR_BYVALUE);
    const bool isweep_ok = sched == PIMC_SCHED_SWEEP && S.need_cells && S.a > 0.0 && !pair_in_moves && h->opt_isweep != 0 && h->opt_faithful_impl == 0 &&
                           S.M <= 256 && S.N <= 8192 && isw_rs_smem_bytes(S.N, S.M) <= 200 * 1024;
    // both paths execute the same sequential definition bit for bit, so the choice is free: option 1 = adaptive (after a run in which more
    // than 30 % of the proposals had to be replayed serially the next 7 runs take the persistent sequential kernel), option 2 = always optimistic
    bool isweep = isweep_ok;
    if (isweep_ok && h->opt_isweep == 1 && n > 0 && h->isw_backoff > 0) { isweep = false; h->isw_backoff -= 1; }
    CK(h, cudaEventRecord(h->ev0, h->stream));
    if (n > 0 && isweep) {
        if (!h->nw_head) {
            int rc = dalloc(h, &h->nw_head, (size_t)S.C * S.ncell * S.M); if (rc) return rc;
            rc = dalloc(h, &h->nw_next, (size_t)S.C * S.N * S.M); if (rc) return rc;
            CK(h, cudaMemset(h->nw_head, 0xFF, sizeof(int) * (size_t)S.C * S.ncell * S.M));
        }
        bool has_rs = false, has_com = false, has_swap = false;
        for (int i = 0; i < nupd; ++i) { int k = h->T.upd[update_ids[i]].kind; has_rs |= k == PIMC_UPD_RESHAPE_LINEAR; has_swap |= k == PIMC_UPD_RESHAPE_SWAP; has_com |= (k == PIMC_UPD_SINGLE_COM || k == PIMC_UPD_POLYMER_COM); }
        ISweepParams IP; memset(&IP, 0, sizeof IP);
        SweepParams &SP = IP.sp;
        SP.nupd = nupd; SP.stats = h->dstats; SP.Sg = h->dS;
        pimc_roundkeys_make(S.seed, &SP.rk);
        for (int i = 0; i < nupd; ++i) { SP.kind[i] = h->T.upd[update_ids[i]].kind; SP.vmax[i] = h->T.upd[update_ids[i]].vmax; SP.upd_id[i] = P.upd_id[i]; SP.w[i] = P.w[i]; IP.upd[i] = h->T.upd[update_ids[i]]; }
        IP.nw_head = h->nw_head; IP.nw_next = h->nw_next; IP.prof = P.prof;
        Sweep2Params SW; memset(&SW, 0, sizeof SW);
        for (int i = 0; i < nupd; ++i) SW.upd[i] = h->T.upd[update_ids[i]];
        if (has_swap && faithful_scratch_doubles(S.N, S.M) * sizeof(double) > 160 * 1024) {
            if (!h->fscr) { int rc = dalloc(h, &h->fscr, (size_t)S.C * faithful_scratch_doubles(S.N, S.M)); if (rc) return rc; }
            SW.fscr = h->fscr;
        }
        MeasParams MP; memset(&MP, 0, sizeof MP);
        MP.nen = nen; MP.nde = nde;
        for (int i = 0; i < nen; ++i) { MP.en_id[i] = P.en_id[i]; MP.en_k0[i] = P.en_k0[i]; }
        for (int i = 0; i < nde; ++i) MP.de_id[i] = P.de_id[i];
        for (long long it = 0; it < n; ++it) {
            SP.iter = h->iter + (unsigned long long)it;
            if (has_rs || has_com) { CK(h, pimc_launch_isweep(S.C, h->stream, S, IP, has_rs, has_com)); LAUNCHED(); launches += (has_rs ? 1 : 0) + (has_com ? 1 : 0); if (has_rs && has_com) LAUNCHED(); }
            if (has_swap) { SW.sp = SP; CK(h, pimc_launch_iswap(S.C, h->stream, S, h->dT, SW)); LAUNCHED(); launches++; }
            const long long ctrv = h->Nctr + it + 1;
            if (nen + nde > 0 && ctrv % h->cfg.Ncycle == 0) {
                MP.ord = ctrv / h->cfg.Ncycle - 1;
                CK(h, pimc_launch_measure(S.C, h->stream, S, h->dT, MP, nullptr)); LAUNCHED(); launches++;
            }
        }
    }
    if (n > 0 && !batched && !isweep) {
        CK(h, pimc_launch_run(S.need_cells != 0, S.C, threads, smem, h->stream, S, h->dT, P));
        LAUNCHED(); launches++;
    }
    if (n > 0 && batched) {
        bool has_rs = false, has_com = false, has_swap = false;
        for (int i = 0; i < nupd; ++i) { int k = h->T.upd[update_ids[i]].kind; has_rs |= k == PIMC_UPD_RESHAPE_LINEAR; has_swap |= k == PIMC_UPD_RESHAPE_SWAP; has_com |= (k == PIMC_UPD_SINGLE_COM || k == PIMC_UPD_POLYMER_COM); }
        SweepParams SP; memset(&SP, 0, sizeof SP);
        SP.nupd = nupd; SP.stats = h->dstats; SP.Sg = h->dS; const char *padenv = getenv("PIMC_EXP_SMEM_PAD"); const size_t smem_pad = padenv ? (size_t)atol(padenv) : 0;
        pimc_roundkeys_make(S.seed, &SP.rk);
        SP.com_stage_off = com_tma ? (int)com_flag : 0;
        for (int i = 0; i < nupd; ++i) { SP.kind[i] = h->T.upd[update_ids[i]].kind; SP.vmax[i] = h->T.upd[update_ids[i]].vmax; }
        CK(h, cudaMemcpyAsync(h->dS, &S, sizeof(DevSys), cudaMemcpyHostToDevice, h->stream));
        for (int i = 0; i < nupd; ++i) { SP.upd_id[i] = P.upd_id[i]; SP.w[i] = P.w[i]; }
        MeasParams MP; memset(&MP, 0, sizeof MP);
        MP.nen = nen; MP.nde = nde;
        for (int i = 0; i < nen; ++i) { MP.en_id[i] = P.en_id[i]; MP.en_k0[i] = P.en_k0[i]; }
        for (int i = 0; i < nde; ++i) MP.de_id[i] = P.de_id[i];
        Sweep2Params SP2; memset(&SP2, 0, sizeof SP2); SP2.cap = bcap;
        for (int i = 0; i < nupd; ++i) SP2.upd[i] = h->T.upd[update_ids[i]];
        // Energy fused into the sweep launch of a measurement iteration (chains whose centre-of-mass sweep streams every worldline anyway)
        const bool fuse_ok = nen > 0 && has_com && h->opt_fuse_energy != 0;
        const bool separate_swap = getenv("PIMC_SEPARATE_SWAP") != nullptr;
        if (fuse_ok && !h->mdone) { int rc = dalloc(h, &h->mdone, (size_t)S.C); if (rc) return rc; }
        // chain-major persistent kernel (pimc_chain.cuh): one launch for the whole call, every CTA takes a chain through all n iterations
        // Automatic choice, from the measurements of profiles/r02_summary.md: the persistent kernel wins when the chains fill at most two
        // rounds of CTA slots (C3 1024 chains: 3.1e10 vs 2.7e10; C5 512 chains: 4.3e10 vs 3.7e10 -- no partly filled waves, no launch gaps),
        // the per-iteration kernels win for large batches (C2 4096 chains: 6.0e10 vs 5.2e10 -- the resident chains' 116 MB do not stay in the
        // two 63 MB L2 partitions, and the in-kernel estimator pass loses the all-SMs streaming pattern of k_measure; C4 2.3 rounds: 1.37e10 vs 1.28e10)
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const int slots = sms * ((S.M + 31) / 32 <= 4 ? 4 : 3);
        const bool chain_major = h->opt_sweep_impl == 3 || (h->opt_sweep_impl == 0 && (getenv("PIMC_CHAIN_MAJOR") != nullptr || (S.C <= 2 * slots && getenv("PIMC_NO_CHAIN_MAJOR") == nullptr)));
        if (chain_major) {
            if (!h->queue) { int rc = dalloc(h, &h->queue, 4); if (rc) return rc; }
            CK(h, cudaMemsetAsync(h->queue, 0, sizeof(int), h->stream));
            ChainParams Q; memset(&Q, 0, sizeof Q);
            SP.iter = h->iter;
            SP2.sp = SP; SP2.fuse = 0; SP2.mp = MP; SP2.mdone = nullptr;
            Q.sw = SP2; Q.mp = MP; Q.mp.ord = 0;
            Q.mp.tma = (nen > 0 && (S.M + 31) / 32 <= 8 && (S.M % 2) == 0 && getenv("PIMC_NO_TMA") == nullptr) ? 1 : 0;
            Q.n = n; Q.Nctr0 = h->Nctr; Q.Ncycle = h->cfg.Ncycle; Q.measure = nen + nde > 0 ? 1 : 0; Q.queue = h->queue;
            size_t smem_ch = smem_rs > smem_cs ? smem_rs : smem_cs;
            const size_t smem_sw = has_swap ? swap_smem_bytes(S.N, S.M) : 0, smem_me = Q.mp.tma ? meas_smem_bytes(SWEEP_THREADS / 32, S.dim, S.M) : 0;
            if (smem_sw > smem_ch) smem_ch = smem_sw;
            if (smem_me > smem_ch) smem_ch = smem_me;
            CK(h, pimc_launch_chain(smem_ch + smem_pad, h->stream, S, h->dT, Q, nullptr)); LAUNCHED(); launches++;
        }
        for (long long it = 0; it < n && !chain_major; ++it) {
            SP.iter = h->iter + (unsigned long long)it;
            const long long ctrv = h->Nctr + it + 1;
            const bool meas_now = nen + nde > 0 && ctrv % h->cfg.Ncycle == 0;
            MP.ord = ctrv / h->cfg.Ncycle - 1;
            SP2.sp = SP; SP2.fuse = meas_now && fuse_ok; SP2.mp = MP; SP2.mdone = h->mdone;
            if (SP2.fuse) CK(h, cudaMemsetAsync(h->mdone, 0, (size_t)S.C, h->stream));
            // chains that picked the swap move run their one proposal on warp 0 of their k_sweep CTA; the stand-alone swap kernel is
            // launched only when no sweep family is listed (or for the A/B: PIMC_SEPARATE_SWAP=1)
            size_t smem_sw = smem_rs > smem_cs ? smem_rs : smem_cs;
            if (has_swap && swap_smem_bytes(S.N, S.M) > smem_sw) smem_sw = swap_smem_bytes(S.N, S.M);
            SP2.swap_in_sweep = (has_swap && (has_com || has_rs) && smem_sw <= 200 * 1024 && !separate_swap) ? 1 : 0;
            if (has_com || has_rs) { CK(h, pimc_launch_sweep(S.C, smem_sw + smem_pad, h->stream, S, h->dT, SP2)); LAUNCHED(); launches++; }
            if (has_swap && !SP2.swap_in_sweep) { CK(h, pimc_launch_swap_iter(S.C, h->stream, S, h->dT, SP2)); LAUNCHED(); launches++; }
            if (meas_now) { CK(h, pimc_launch_measure(S.C, h->stream, S, h->dT, MP, SP2.fuse ? h->mdone : nullptr)); LAUNCHED(); launches++; }
        }
    }
    CK(h, cudaEventRecord(h->ev1, h->stream));
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    float ms = 0; CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    unsigned long long st[32]; CK(h, cudaMemcpy(st, h->dstats, sizeof st, cudaMemcpyDeviceToHost));
    if (n > 0 && isweep && st[13] > 0) {
        h->isw_replay_frac = (double)st[12] / (double)st[13];
        if (h->opt_isweep == 1 && h->isw_replay_frac > 0.30) h->isw_backoff = 7;
    }
    if (P.prof && n > 0 && isweep) {
        const double sw = (double)st[4 + 9] / S.N > 0 ? (double)st[4 + 9] / S.N : 1.0;   // chain-sweeps executed
        fprintf(stderr, "[pimc prof] isweep %.3f ms, %lld iterations x %d chains, %.0f chain-sweeps, %.2f replays per chain-sweep; cycles per chain-sweep", ms, (long long)n, S.C, sw, (double)st[4 + 8] / sw);
        const char *ph[6] = { "setup", "P1", "P3", "P4", "P5", "book" };
        fprintf(stderr, " reshape:"); for (int i = 0; i < 6; ++i) fprintf(stderr, " %s %.0f", ph[i], (double)st[4 + i] / sw);
        fprintf(stderr, " | com:"); for (int i = 0; i < 6; ++i) fprintf(stderr, " %s %.0f", ph[i], (double)st[4 + 10 + i] / sw);
        fprintf(stderr, "  (each divided by ALL chain-sweeps)\n");
    }
    if (P.prof && n > 0 && !isweep) {
        const char *kn[4] = { "ReshapeLinear", "ReshapeSwap", "SingleCOM", "PolymerCOM" };
        fprintf(stderr, "[pimc prof] k_run %.3f ms, %lld iterations x %d chains; mean cycles per proposal:", ms, (long long)n, S.C);
        for (int k = 0; k < 4; ++k) if (st[4 + 4 + k]) fprintf(stderr, " %s %.0f (x%llu)", kn[k], (double)st[4 + k] / (double)st[4 + 4 + k], st[4 + 4 + k]);
        fprintf(stderr, "; per iteration: bookkeeping %.0f, estimators %.0f\n", (double)st[4 + 8] / ((double)n * S.C), (double)st[4 + 9] / ((double)n * S.C));
    }
    h->iter += (unsigned long long)n;
    if (cadence) {
        h->N_MC += nmeas; h->Nctr = (h->Nctr + n) % h->cfg.Ncycle;
        // per-block all-reduce of the Energy accumulators, queued on the side stream: it overlaps the next block's moves
        for (int i = 0; i < nen && h->comm; ++i) if (h->g_en_upto[P.en_id[i]] == h->en_count[P.en_id[i]]) { int rc = comm_reduce_energy(h, P.en_id[i], h->en_count[P.en_id[i]], h->en_count[P.en_id[i]] + nmeas); if (rc) return rc; }
        for (int i = 0; i < nen; ++i) h->en_count[P.en_id[i]] += nmeas;
        for (int i = 0; i < nde; ++i) h->de_ndata[P.de_id[i]] += nmeas * (long long)S.M * S.C;
    }
    if (stats) {
        stats->iterations = n; stats->proposals = (int64_t)st[0]; stats->accepted = 0; stats->bead_moves = (int64_t)st[2];
        stats->measurements = nmeas; stats->launches = launches; stats->kernel_ms = ms;
        long long acc = 0; // accepted totals come from the per-update counters
        for (int i = 0; i < nupd; ++i) { int64_t a = 0; bool dup = false; for (int k = 0; k < i; ++k) dup |= update_ids[k] == update_ids[i]; if (dup) continue;
            pimc_update_get(h, update_ids[i], -1, nullptr, nullptr, nullptr, nullptr, &a, nullptr); acc += a; }
        stats->accepted = acc; // cumulative over the life of the update objects
    }
    return PIMC_OK;
}
extern "C" int pimc_run(pimc_handle *h, int64_t n, const int32_t *update_ids, const int64_t *every, int32_t nupd,
                        const int32_t *energy_ids, int32_t nen, const int32_t *density_ids, int32_t nde, int32_t sched, pimc_run_stats *stats)
{
    return run_core(h, n, update_ids, every, nupd, energy_ids, nen, density_ids, nde, sched, false, stats);
}

// ---- estimators the reference lists as TODO (measurement.jl:125-127): g(r) and winding number ----
extern "C" int pimc_paircorr_create(pimc_handle *h, int64_t nbins, double rmax, int32_t *id)
{
    if (!h || !id || nbins < 1 || !(rmax > 0)) return PIMC_ERR_INVALID;
    if (h->npc >= PIMC_MAXP) { SETERR(h, "at most %d pair-correlation objects per handle", PIMC_MAXP); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    PcDev &G = h->pc[h->npc]; G.nbins = nbins; G.rmax = rmax; G.bin = rmax / (double)nbins;
    int TS, sh; if (pimc_paircorr_smem(h->S, G, &TS, &sh) > 200 * 1024) { SETERR(h, "pair correlation: N = %d does not fit the shared-memory tile", h->S.N); return PIMC_ERR_UNSUPPORTED; }
    int rc = dalloc(h, &G.hist, (size_t)nbins); if (rc) return rc;
    h->pc_ndata[h->npc] = 0;
    *id = h->npc++;
    return PIMC_OK;
}
extern "C" int pimc_paircorr_measure(pimc_handle *h, int32_t id)
{
    if (!h || id < 0 || id >= h->npc) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, pimc_launch_paircorr(h->S.C, h->stream, h->S, h->pc[id])); LAUNCHED();
    CK(h, cudaStreamSynchronize(h->stream));
    h->pc_ndata[id] += (long long)h->S.M * h->S.C;
    return PIMC_OK;
}
// hist: nbins pair counts summed over this handle's chains (over all ranks with a communicator attached); ndata likewise
extern "C" int pimc_paircorr_read(pimc_handle *h, int32_t id, double *hist, int64_t *ndata, double *bin)
{
    if (!h || id < 0 || id >= h->npc) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    PcDev &G = h->pc[id]; const size_t sz = (size_t)G.nbins;
    long long nd = h->pc_ndata[id];
    std::vector<unsigned long long> tmp(sz + 1);
    if (h->comm) {
        if (h->g_u64_n < sz + 1) { int rc = dalloc(h, &h->g_u64, sz + 1); if (rc) return rc; h->g_u64_n = sz + 1; }
        CK(h, cudaMemcpyAsync(h->g_u64, G.hist, sz * 8, cudaMemcpyDeviceToDevice, h->cstream));
        const unsigned long long ndu = (unsigned long long)nd;
        CK(h, cudaMemcpyAsync(h->g_u64 + sz, &ndu, 8, cudaMemcpyHostToDevice, h->cstream));
        NK(h, nccl_api()->AllReduce(h->g_u64, h->g_u64, sz + 1, ncclUint64, ncclSum, (ncclComm_t)h->comm, h->cstream));
        CK(h, cudaMemcpyAsync(tmp.data(), h->g_u64, (sz + 1) * 8, cudaMemcpyDeviceToHost, h->cstream));
        CK(h, cudaStreamSynchronize(h->cstream));
        nd = (long long)tmp[sz];
    } else CK(h, cudaMemcpy(tmp.data(), G.hist, sz * 8, cudaMemcpyDeviceToHost));
    if (hist) for (size_t i = 0; i < sz; ++i) hist[i] = (double)tmp[i];
    if (ndata) *ndata = nd;
    if (bin) *bin = G.bin;
    return PIMC_OK;
}
extern "C" int pimc_winding_create(pimc_handle *h, int64_t cap, int32_t *id)
{
    if (!h || !id || cap < 1) return PIMC_ERR_INVALID;
    if (h->nwi >= PIMC_MAXW) { SETERR(h, "at most %d winding objects per handle", PIMC_MAXW); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    WiDev &W = h->wi[h->nwi]; W.cap = cap;
    int rc = dalloc(h, &W.W, (size_t)cap * h->S.dim * h->S.C); if (rc) return rc;
    h->wi_count[h->nwi] = 0;
    *id = h->nwi++;
    return PIMC_OK;
}
// winding number of the CURRENT configuration of every chain: W[chains][dim]
extern "C" int pimc_winding_now(pimc_handle *h, double *W)
{
    if (!h || !W) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    const int C = h->S.C, dim = h->S.dim;
    TmpBuf t; WiDev tmp; tmp.cap = 1; tmp.W = t.up((double *)nullptr, (size_t)dim * C); if (!tmp.W) return PIMC_ERR_NOMEM;
    CK(h, pimc_launch_winding(C, h->stream, h->S, tmp, 0)); LAUNCHED();
    CK(h, cudaStreamSynchronize(h->stream));
    std::vector<double> w((size_t)dim * C);
    CK(h, cudaMemcpy(w.data(), tmp.W, sizeof(double) * w.size(), cudaMemcpyDeviceToHost));
    for (int c = 0; c < C; ++c) for (int k = 0; k < dim; ++k) W[(size_t)c * dim + k] = w[(size_t)k * C + c];
    return PIMC_OK;
}
// series of one chain: W[n][dim] (chain >= 0), or with chain = -1 the chain-mean of W^2 = sum_k W_k^2 per measurement: W2[n] (local chains)
extern "C" int pimc_winding_read(pimc_handle *h, int32_t id, int32_t chain, double *out, int64_t cap, int64_t *n)
{
    if (!h || id < 0 || id >= h->nwi || chain < -1 || chain >= h->S.C || cap < 0) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    const long long cnt = h->wi_count[id]; if (n) *n = cnt;
    long long m = cnt < h->wi[id].cap ? cnt : h->wi[id].cap; if (m > cap) m = cap;
    if (m <= 0 || !out) return PIMC_OK;
    const int C = h->S.C, dim = h->S.dim;
    std::vector<double> w((size_t)m * dim * C);
    CK(h, cudaMemcpy(w.data(), h->wi[id].W, sizeof(double) * w.size(), cudaMemcpyDeviceToHost));
    for (long long k = 0; k < m; ++k) {
        if (chain >= 0) for (int d = 0; d < dim; ++d) out[k * dim + d] = w[((size_t)k * dim + d) * C + chain];
        else { double s2 = 0.0; for (int c = 0; c < C; ++c) for (int d = 0; d < dim; ++d) { const double v = w[((size_t)k * dim + d) * C + c]; s2 += v * v; } out[k] = s2 / C; }
    }
    return PIMC_OK;
}

// ---- static structure factor and compressibility (`#TODO Compressibilty`, measurement.jl:127) ----
extern "C" int pimc_structure_create(pimc_handle *h, int32_t kmax, int32_t *id)
{
    if (!h || !id) return PIMC_ERR_INVALID;
    if (kmax < 1 || kmax > PIMC_SK_KMAX) { SETERR(h, "structure factor: kmax must lie in 1..%d", PIMC_SK_KMAX); return PIMC_ERR_INVALID; }
    if (h->nsk >= PIMC_MAXS) { SETERR(h, "at most %d structure-factor objects per handle", PIMC_MAXS); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    int TS; if (pimc_structure_smem(h->S, &TS) > 200 * 1024) { SETERR(h, "structure factor: N = %d does not fit the shared-memory tile", h->S.N); return PIMC_ERR_UNSUPPORTED; }
    SkDev &K = h->sk[h->nsk]; K.kmax = kmax;
    int rc = dalloc(h, &K.S, (size_t)h->S.C * (kmax + 1) * (2 * kmax + 1)); if (rc) return rc;
    h->sk_ndata[h->nsk] = 0;
    *id = h->nsk++;
    return PIMC_OK;
}
extern "C" int pimc_structure_measure(pimc_handle *h, int32_t id)
{
    if (!h || id < 0 || id >= h->nsk) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, pimc_launch_structure(h->S.C, h->stream, h->S, h->sk[id])); LAUNCHED();
    CK(h, cudaStreamSynchronize(h->stream));
    h->sk_ndata[id] += (long long)h->S.M * h->S.C;
    return PIMC_OK;
}
// sums[(kmax + 1) * (2 kmax + 1)]: sum of |rho_k|^2 over slices, measurements and this handle's chains in chain order (over all ranks with a
// communicator attached), entry (a, b + kmax) for k = (pi / L)(a, b); ndata likewise.  S(k) = sums / (ndata * N).
extern "C" int pimc_structure_read(pimc_handle *h, int32_t id, double *sums, int64_t *ndata, int32_t *kmax)
{
    if (!h || id < 0 || id >= h->nsk) return PIMC_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    const SkDev &K = h->sk[id]; const size_t nv = (size_t)(K.kmax + 1) * (2 * K.kmax + 1), C = (size_t)h->S.C;
    std::vector<double> all(C * nv), tot(nv + 1, 0.0);
    CK(h, cudaMemcpy(all.data(), K.S, all.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < C; ++c) for (size_t v = 0; v < nv; ++v) tot[v] += all[c * nv + v];
    tot[nv] = (double)h->sk_ndata[id];
    if (h->comm) {
        if (h->g_f64_n < nv + 1) { int rc = dalloc(h, &h->g_f64, nv + 1); if (rc) return rc; h->g_f64_n = nv + 1; }
        CK(h, cudaMemcpyAsync(h->g_f64, tot.data(), (nv + 1) * sizeof(double), cudaMemcpyHostToDevice, h->cstream));
        NK(h, nccl_api()->AllReduce(h->g_f64, h->g_f64, nv + 1, ncclDouble, ncclSum, (ncclComm_t)h->comm, h->cstream));
        CK(h, cudaMemcpyAsync(tot.data(), h->g_f64, (nv + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->cstream));
        CK(h, cudaStreamSynchronize(h->cstream));
    }
    if (sums) for (size_t v = 0; v < nv; ++v) sums[v] = tot[v];
    if (ndata) *ndata = (int64_t)tot[nv];
    if (kmax) *kmax = K.kmax;
    return PIMC_OK;
}
// isothermal compressibility from the long-wavelength limit S(k -> 0) = rho k_B T kappa_T, estimated on the smallest shell of the box
// (|k| = pi / L: the mean of S over (1, 0) and (0, 1); 1-D: (1)):  kappa_T = beta S(k_min) / rho, rho = N / (2L)^dim, beta = M tau.
extern "C" int pimc_compressibility(pimc_handle *h, int32_t id, double *kappa, double *s_kmin)
{
    if (!h || id < 0 || id >= h->nsk || !kappa) return PIMC_ERR_INVALID;
    const int kmax = h->sk[id].kmax, nb = 2 * kmax + 1;
    std::vector<double> sums((size_t)(kmax + 1) * nb); int64_t nd = 0;
    int rc = pimc_structure_read(h, id, sums.data(), &nd, nullptr); if (rc) return rc;
    if (nd <= 0) { SETERR(h, "compressibility: the structure-factor object holds no measurement"); return PIMC_ERR_STATE; }
    const DevSys &S = h->S;
    double s0 = sums[(size_t)1 * nb + kmax];                         // (1, 0)
    if (S.dim > 1) s0 = 0.5 * (s0 + sums[(size_t)0 * nb + kmax + 1]);   // (0, 1)
    s0 /= (double)nd * S.N;
    double vol = 1.0; for (int k = 0; k < S.dim; ++k) vol *= 2 * S.L;
    *kappa = (S.M * S.tau) * s0 / ((double)S.N / vol);
    if (s_kmin) *s_kmin = s0;
    return PIMC_OK;
}

// run! with the full estimator list.  Energy / Density are evaluated inside the run kernels; the estimators above are launched between
// segments of the run that end on a measurement event (same cadence: measurement_Z_sector, measurement.jl:1-17).
extern "C" int pimc_run_ex(pimc_handle *h, int64_t n, const int32_t *update_ids, const int64_t *every, int32_t nupd,
                           const pimc_measurements *meas, int32_t sched, pimc_run_stats *stats)
{
    if (!h) return PIMC_ERR_INVALID;
    pimc_measurements none; memset(&none, 0, sizeof none);
    const pimc_measurements &Z = meas ? *meas : none;
    if (Z.npaircorr < 0 || Z.npaircorr > PIMC_MAXP || Z.nwinding < 0 || Z.nwinding > PIMC_MAXW || Z.nstructure < 0 || Z.nstructure > PIMC_MAXS) { SETERR(h, "pimc_run_ex: bad estimator counts"); return PIMC_ERR_INVALID; }
    for (int i = 0; i < Z.nstructure; ++i) if (Z.structure_ids[i] < 0 || Z.structure_ids[i] >= h->nsk) { SETERR(h, "bad structure-factor id"); return PIMC_ERR_INVALID; }
    for (int i = 0; i < Z.npaircorr; ++i) if (Z.paircorr_ids[i] < 0 || Z.paircorr_ids[i] >= h->npc) { SETERR(h, "bad pair-correlation id"); return PIMC_ERR_INVALID; }
    for (int i = 0; i < Z.nwinding; ++i) if (Z.winding_ids[i] < 0 || Z.winding_ids[i] >= h->nwi) { SETERR(h, "bad winding id"); return PIMC_ERR_INVALID; }
    const int nextra = Z.npaircorr + Z.nwinding + Z.nstructure;
    if (nextra == 0) return run_core(h, n, update_ids, every, nupd, Z.energy_ids, Z.nenergy, Z.density_ids, Z.ndensity, sched, false, stats);
    {   // winding series overflow: like Energy, the pre-sized vector must hold every sample of this run
        const long long nmeas = (h->Nctr + n) / h->cfg.Ncycle;
        for (int i = 0; i < Z.nwinding; ++i) if (h->wi_count[Z.winding_ids[i]] + nmeas > h->wi[Z.winding_ids[i]].cap) { SETERR(h, "winding buffer would overflow"); return PIMC_ERR_STATE; }
    }
    pimc_run_stats tot; memset(&tot, 0, sizeof tot);
    long long remaining = n;
    while (remaining > 0) {
        const long long to_event = h->cfg.Ncycle - h->Nctr, seg = remaining < to_event ? remaining : to_event;
        pimc_run_stats st; memset(&st, 0, sizeof st);
        int rc = run_core(h, seg, update_ids, every, nupd, Z.energy_ids, Z.nenergy, Z.density_ids, Z.ndensity, sched, true, &st);
        if (rc) return rc;
        tot.iterations += st.iterations; tot.proposals += st.proposals; tot.bead_moves += st.bead_moves; tot.measurements += st.measurements;
        tot.launches += st.launches; tot.kernel_ms += st.kernel_ms; tot.accepted = st.accepted;
        if (seg == to_event) {          // a measurement event ends this segment
            cudaEvent_t e0 = h->ev0, e1 = h->ev1;
            CK(h, cudaEventRecord(e0, h->stream));
            for (int i = 0; i < Z.npaircorr; ++i) { CK(h, pimc_launch_paircorr(h->S.C, h->stream, h->S, h->pc[Z.paircorr_ids[i]])); LAUNCHED(); tot.launches++; h->pc_ndata[Z.paircorr_ids[i]] += (long long)h->S.M * h->S.C; }
            for (int i = 0; i < Z.nwinding; ++i) { const int id = Z.winding_ids[i]; CK(h, pimc_launch_winding(h->S.C, h->stream, h->S, h->wi[id], h->wi_count[id])); LAUNCHED(); tot.launches++; h->wi_count[id] += 1; }
            for (int i = 0; i < Z.nstructure; ++i) { const int id = Z.structure_ids[i]; CK(h, pimc_launch_structure(h->S.C, h->stream, h->S, h->sk[id])); LAUNCHED(); tot.launches++; h->sk_ndata[id] += (long long)h->S.M * h->S.C; }
            CK(h, cudaEventRecord(e1, h->stream));
            CK(h, cudaStreamSynchronize(h->stream));
            float ms = 0; CK(h, cudaEventElapsedTime(&ms, e0, e1)); tot.kernel_ms += ms;
        }
        remaining -= seg;
    }
    if (stats) *stats = tot;
    return PIMC_OK;
}

// ---- checkpoint / resume: the complete chain state as one host blob ----
// The reference's savetools (examples/tools/savetools.jl:4-34) write the paths only, so a reloaded Julia run restarts its step adaptation and
// counters.  A resumed run here continues bit for bit: positions, permutation, the CACHED link actions (stale links of compat B14 included),
// cell lists with their list order and multiplicities (B13), iteration counter of the addressed RNG, measurement cadence, every update
// object's adaptive variable / counters / acceptance window, every estimator's accumulators.
struct StateHeader {
    unsigned long long magic; int version, dim, M, N, C, need_cells, ncell, nupd, nen, nde, npc, nwi; unsigned chain_offset; unsigned long long seed;
    unsigned long long iter; long long N_MC, Nctr;
    long long en_count[PIMC_MAXE], en_cap[PIMC_MAXE], de_ndata[PIMC_MAXD], de_nbins[PIMC_MAXD];
    long long pc_ndata[PIMC_MAXP], pc_nbins[PIMC_MAXP], wi_count[PIMC_MAXW], wi_cap[PIMC_MAXW];
    int nsk, sk_kmax[PIMC_MAXS]; long long sk_ndata[PIMC_MAXS];
    int upd_kind[PIMC_MAXU], upd_ring_words[PIMC_MAXU]; long long upd_adj[PIMC_MAXU], upd_range[PIMC_MAXU];
    double upd_vmin[PIMC_MAXU], upd_vmax[PIMC_MAXU], upd_minacc[PIMC_MAXU], upd_maxacc[PIMC_MAXU];
};
#define PIMC_STATE_MAGIC 0x50494d4342323030ull   /* "PIMCB200" */
struct Seg { void *p; size_t bytes; };
static std::vector<Seg> state_segments(pimc_handle *h)
{
    DevSys &S = h->S; const size_t C = S.C, nb = C * S.N * S.M;
    std::vector<Seg> v;
    v.push_back({ S.r, nb * S.dim * sizeof(double) }); v.push_back({ S.Vl, nb * sizeof(double) }); v.push_back({ S.next, C * S.N * sizeof(int) });
    if (S.need_cells) {
        v.push_back({ S.bins, nb * sizeof(*S.bins) }); v.push_back({ S.cell_head, C * S.M * S.ncell * sizeof(int) });
        v.push_back({ S.cell_next, C * S.M * S.N * sizeof(int) }); v.push_back({ S.mult, nb * sizeof(*S.mult) });
    }
    for (int i = 0; i < h->nupd; ++i) {
        UpdDev &U = h->T.upd[i];
        v.push_back({ U.var, C * 8 }); v.push_back({ U.tries, C * 8 }); v.push_back({ U.accepted, C * 8 }); v.push_back({ U.tries_var, C * 8 });
        v.push_back({ U.bead_moves, C * 8 }); v.push_back({ U.ring_head, C * 4 }); v.push_back({ U.ring_len, C * 4 }); v.push_back({ U.ring_sum, C * 4 });
        v.push_back({ U.ring, C * (size_t)U.ring_words * 4 });
    }
    for (int i = 0; i < h->nen; ++i) {
        EnDev &E = h->T.en[i]; const size_t rows = (size_t)(h->en_count[i] < E.cap ? h->en_count[i] : E.cap);
        v.push_back({ E.E, rows * C * 8 }); v.push_back({ E.Ev, rows * C * 8 }); v.push_back({ E.acc, C * 5 * 8 });
    }
    for (int i = 0; i < h->nde; ++i) { DeDev &D = h->T.de[i]; v.push_back({ D.dens, (S.dim == 2 ? (size_t)D.nbins * D.nbins : (size_t)D.nbins) * 8 }); }
    for (int i = 0; i < h->npc; ++i) v.push_back({ h->pc[i].hist, (size_t)h->pc[i].nbins * 8 });
    for (int i = 0; i < h->nwi; ++i) { const size_t rows = (size_t)(h->wi_count[i] < h->wi[i].cap ? h->wi_count[i] : h->wi[i].cap); v.push_back({ h->wi[i].W, rows * S.dim * C * 8 }); }
    for (int i = 0; i < h->nsk; ++i) v.push_back({ h->sk[i].S, C * (size_t)(h->sk[i].kmax + 1) * (2 * h->sk[i].kmax + 1) * 8 });
    return v;
}
static void state_header(pimc_handle *h, StateHeader *H)
{
    memset(H, 0, sizeof *H);
    DevSys &S = h->S;
    H->magic = PIMC_STATE_MAGIC; H->version = 3; H->npc = h->npc; H->nwi = h->nwi; H->nsk = h->nsk;
    for (int i = 0; i < h->nsk; ++i) { H->sk_ndata[i] = h->sk_ndata[i]; H->sk_kmax[i] = h->sk[i].kmax; }
    for (int i = 0; i < h->npc; ++i) { H->pc_ndata[i] = h->pc_ndata[i]; H->pc_nbins[i] = h->pc[i].nbins; }
    for (int i = 0; i < h->nwi; ++i) { H->wi_count[i] = h->wi_count[i]; H->wi_cap[i] = h->wi[i].cap; } H->dim = S.dim; H->M = S.M; H->N = S.N; H->C = S.C; H->need_cells = S.need_cells; H->ncell = S.ncell;
    H->nupd = h->nupd; H->nen = h->nen; H->nde = h->nde; H->chain_offset = S.chain_offset; H->seed = S.seed;
    H->iter = h->iter; H->N_MC = h->N_MC; H->Nctr = h->Nctr;
    for (int i = 0; i < h->nen; ++i) { H->en_count[i] = h->en_count[i]; H->en_cap[i] = h->T.en[i].cap; }
    for (int i = 0; i < h->nde; ++i) { H->de_ndata[i] = h->de_ndata[i]; H->de_nbins[i] = h->T.de[i].nbins; }
    for (int i = 0; i < h->nupd; ++i) {
        const UpdDev &U = h->T.upd[i];
        H->upd_kind[i] = U.kind; H->upd_ring_words[i] = U.ring_words; H->upd_adj[i] = U.adj; H->upd_range[i] = U.range;
        H->upd_vmin[i] = U.vmin; H->upd_vmax[i] = U.vmax; H->upd_minacc[i] = U.minacc; H->upd_maxacc[i] = U.maxacc;
    }
}
extern "C" int pimc_state_size(pimc_handle *h, int64_t *bytes)
{
    if (!h || !bytes) return PIMC_ERR_INVALID;
    size_t n = sizeof(StateHeader);
    for (const Seg &s : state_segments(h)) n += (s.bytes + 7) & ~(size_t)7;
    *bytes = (int64_t)n; return PIMC_OK;
}
extern "C" int pimc_get_state(pimc_handle *h, void *buf, int64_t cap)
{
    if (!h || !buf) return PIMC_ERR_INVALID;
    int64_t need; pimc_state_size(h, &need);
    if (cap < need) { SETERR(h, "state buffer of %lld bytes, %lld needed", (long long)cap, (long long)need); return PIMC_ERR_INVALID; }
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    StateHeader H; state_header(h, &H);
    char *p = (char *)buf; memcpy(p, &H, sizeof H); p += sizeof H;
    for (const Seg &s : state_segments(h)) {
        if (s.bytes) CK(h, cudaMemcpy(p, s.p, s.bytes, cudaMemcpyDeviceToHost));
        p += (s.bytes + 7) & ~(size_t)7;
    }
    return PIMC_OK;
}
extern "C" int pimc_set_state(pimc_handle *h, const void *buf, int64_t bytes)
{
    if (!h || !buf || bytes < (int64_t)sizeof(StateHeader)) return PIMC_ERR_INVALID;
    StateHeader H; memcpy(&H, buf, sizeof H);
    DevSys &S = h->S;
    if (H.magic != PIMC_STATE_MAGIC || H.version != 3) { SETERR(h, "not a pimc_b200 state blob (magic / version)"); return PIMC_ERR_INVALID; }
    if (H.dim != S.dim || H.M != S.M || H.N != S.N || H.C != S.C || H.need_cells != S.need_cells || H.ncell != S.ncell || H.seed != S.seed || H.chain_offset != S.chain_offset) {
        SETERR(h, "state blob belongs to another System (dim/M/N/chains/cells/seed/chain_offset differ)"); return PIMC_ERR_STATE; }
    if (H.npc != h->npc || H.nwi != h->nwi) { SETERR(h, "state blob holds %d/%d pair-correlation/winding objects, the handle %d/%d", H.npc, H.nwi, h->npc, h->nwi); return PIMC_ERR_STATE; }
    for (int i = 0; i < h->npc; ++i) if (H.pc_nbins[i] != h->pc[i].nbins) { SETERR(h, "pair-correlation object %d differs in nbins", i); return PIMC_ERR_STATE; }
    for (int i = 0; i < h->nwi; ++i) if (H.wi_cap[i] != h->wi[i].cap) { SETERR(h, "winding object %d differs in capacity", i); return PIMC_ERR_STATE; }
    if (H.nsk != h->nsk) { SETERR(h, "state blob holds %d structure-factor objects, the handle %d", H.nsk, h->nsk); return PIMC_ERR_STATE; }
    for (int i = 0; i < h->nsk; ++i) if (H.sk_kmax[i] != h->sk[i].kmax) { SETERR(h, "structure-factor object %d differs in kmax", i); return PIMC_ERR_STATE; }
    if (H.nupd != h->nupd || H.nen != h->nen || H.nde != h->nde) { SETERR(h, "state blob holds %d/%d/%d update/Energy/Density objects, the handle %d/%d/%d: create the same objects in the same order first", H.nupd, H.nen, H.nde, h->nupd, h->nen, h->nde); return PIMC_ERR_STATE; }
    for (int i = 0; i < h->nupd; ++i) if (H.upd_kind[i] != h->T.upd[i].kind || H.upd_range[i] != h->T.upd[i].range) { SETERR(h, "update object %d differs in kind or window range", i); return PIMC_ERR_STATE; }
    for (int i = 0; i < h->nen; ++i) if (H.en_cap[i] != h->T.en[i].cap) { SETERR(h, "Energy object %d differs in capacity", i); return PIMC_ERR_STATE; }
    for (int i = 0; i < h->nde; ++i) if (H.de_nbins[i] != h->T.de[i].nbins) { SETERR(h, "Density object %d differs in nbins", i); return PIMC_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    h->iter = H.iter; h->N_MC = H.N_MC; h->Nctr = H.Nctr;
    for (int i = 0; i < h->nen; ++i) { h->en_count[i] = H.en_count[i]; h->g_en_upto[i] = 0; }
    for (int i = 0; i < h->nde; ++i) h->de_ndata[i] = H.de_ndata[i];
    for (int i = 0; i < h->npc; ++i) h->pc_ndata[i] = H.pc_ndata[i];
    for (int i = 0; i < h->nwi; ++i) h->wi_count[i] = H.wi_count[i];
    for (int i = 0; i < h->nsk; ++i) h->sk_ndata[i] = H.sk_ndata[i];
    for (int i = 0; i < h->nupd; ++i) { UpdDev &U = h->T.upd[i]; U.adj = H.upd_adj[i]; U.vmin = H.upd_vmin[i]; U.vmax = H.upd_vmax[i]; U.minacc = H.upd_minacc[i]; U.maxacc = H.upd_maxacc[i]; }
    int64_t need; pimc_state_size(h, &need);   // after en_count is restored: the Energy segments hold the rows taken so far
    if (bytes < need) { SETERR(h, "state blob truncated (%lld of %lld bytes)", (long long)bytes, (long long)need); return PIMC_ERR_INVALID; }
    const char *p = (const char *)buf + sizeof H;
    for (const Seg &s : state_segments(h)) {
        if (s.bytes) CK(h, cudaMemcpy(s.p, p, s.bytes, cudaMemcpyHostToDevice));
        p += (s.bytes + 7) & ~(size_t)7;
    }
    int rc = sync_tables(h); if (rc) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return PIMC_OK;
}
