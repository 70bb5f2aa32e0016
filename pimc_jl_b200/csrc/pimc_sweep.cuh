// pimc_sweep.cuh -- throughput kernels of the SWEEP schedule (one launch per iteration and update family, one CTA per
// chain).  Same draws, same arithmetic per bead and therefore the same trajectories as the persistent k_run;
// only the order of the Delta-U reductions differs (warp shuffles).
//
//   k_reshape_sweep : ReshapeLinear (reshape.jl:31-91) for every worldline of the chain in one time window.
//        phase A  lanes = beads: Philox -> Box-Muller -> xi*sigma_k staged in shared memory          (fp64 / issue bound)
//        phase B  lanes = (task, dim): the serial staging recurrence of levy! (helper.jl:129-135), in place in shared memory
//        phase C  lanes = beads: teleport, potential
//        phase D  warp per task: Delta-U by warp shuffles, Metropolis, coalesced commit of positions + link cache
//   k_com_sweep     : Single/PolymerCenterOfMass (com.jl) for every worldline, one warp per proposal, one pass over HBM.
#pragma once
#include "pimc_moves.cuh"

#define SWEEP_THREADS 128
#define SWEEP_BCAP 1024   // staged rows per batch (xs, ys, pv: 24 B each)
#define SWEEP_TBMAX 256   // tasks per batch

struct SweepParams {
    unsigned long long iter;
    const DevSys *Sg;   // device copy of the system descriptor (for out-of-line slow paths)
    int nupd; int upd_id[PIMC_MAXU]; double w[PIMC_MAXU];
    int kind[PIMC_MAXU]; double vmax[PIMC_MAXU];   // copies of the update descriptors' constants (no global load on the prologue path)
    unsigned long long *stats;
    pimc_roundkeys rk;  // Philox round keys of the seed (constant-bank operands)
};

// teleport (propagator.jl:30-32) without the IEEE division on the fast path: q = x * (1/2L) differs from x / 2L by
// <= 1 ulp, so floor(q + 0.5) can differ only when q + 0.5 sits within a few ulp of an integer; that case takes the exact path.
__device__ __forceinline__ double d_teleport_fast(double x, double L, double twoL, double inv2L)
{
    double s = x * inv2L + 0.5;
    double f = floor(s);
    double d = s - f;
    double eps = 1e-9 * fmax(1.0, fabs(s));
    if (d < eps || d > 1.0 - eps) f = floor(x / twoL + 0.5);
    return ((x + L) - f * twoL) - L;
}

template <int POT> __device__ __forceinline__ double d_pot_t(const PotDev &p, double x, double y, int dim)
{
    if (POT == PIMC_POT_ZERO) return 0.0;
    if (POT == PIMC_POT_HARMONIC) { double s = x * x; if (dim > 1) s = s + y * y; return (0.5 * p.k) * s; }
    return d_pot(p, x, y, dim);
}

__device__ __forceinline__ int d_pick_update(const SweepParams &P, const pimc_u4 &di)
{
    return d_sample_weighted(P.w, P.nupd, pimc_u01_co(di.w[0], di.w[1]));
}

// apply! bookkeeping of one sweep (simulation.jl:19-27), one thread.  Same final state as d_ring_push per proposal, but the
// ring words are buffered in registers (one global read/write per 32 proposals instead of a dependent RMW per proposal).
__device__ __forceinline__ void d_bookkeep_sweep(const UpdDev &U, int c, const unsigned char *flag, int ntask, unsigned long long beads,
                                                 unsigned long long *stats)
{
    unsigned *ring = U.ring + (size_t)c * U.ring_words;
    const int cap = (int)U.range + 1, range = (int)U.range;
    int head = U.ring_head[c], len = U.ring_len[c], sum = U.ring_sum[c];
    long long tries = U.tries_var[c];
    const long long tries0 = tries;
    long long tr = U.tries[c], ac = U.accepted[c];
    int cnt = 0, tail = head + len; if (tail >= cap) tail -= cap;
    int cw = -1, ew = -1; unsigned cword = 0, eword = 0;
    for (int slot = 0; slot < ntask; ++slot) {
        const int f = flag[slot];
        if (f == 2) continue;
        cnt += 1; tr += 1;
        if (f == 3) continue;
        ac += f; tries += 1;
        const int w = tail >> 5;
        if (w != cw) { if (cw >= 0) ring[cw] = cword; cw = w; cword = ring[w]; }
        const unsigned bit = 1u << (tail & 31);
        cword = f ? (cword | bit) : (cword & ~bit);
        tail = tail + 1 == cap ? 0 : tail + 1;
        len += 1; sum += f;
        if (len > range) {
            const int hw = head >> 5; unsigned eb;
            if (hw == cw) eb = (cword >> (head & 31)) & 1u;
            else { if (hw != ew) { ew = hw; eword = ring[hw]; } eb = (eword >> (head & 31)) & 1u; }
            sum -= (int)eb; head = head + 1 == cap ? 0 : head + 1; len -= 1;
        }
    }
    if (cw >= 0) ring[cw] = cword;
    RingReg R; R.head = head; R.len = len; R.sum = sum; R.tries = tries;
    bool adj = cnt > 0 && (tries / U.adj) != (tries0 / U.adj);
    U.ring_head[c] = head; U.ring_len[c] = len; U.ring_sum[c] = sum; U.tries_var[c] = tries;
    U.tries[c] = tr; U.accepted[c] = ac; U.bead_moves[c] += (long long)beads;
    if (adj) d_adjust(U, c, R);
    if (stats) { atomicAdd(stats + 0, (unsigned long long)cnt); atomicAdd(stats + 2, beads); }
}

template <int POT>
__global__ void __launch_bounds__(SWEEP_THREADS, 1024 / SWEEP_THREADS) k_reshape_sweep(DevSys S, const DevTables *__restrict__ T, SweepParams P)
{
    extern __shared__ double sm[];
    double *xs = sm, *ys = sm + SWEEP_BCAP, *pv = sm + 2 * SWEEP_BCAP;               // pv only touched when POT != 0
    double *pend = (POT == PIMC_POT_ZERO) ? sm + 2 * SWEEP_BCAP : sm + 3 * SWEEP_BCAP;
    int *t_m = (int *)pend;                    // [TBMAX] links of the task
    int *t_off = t_m + SWEEP_TBMAX;            // [TBMAX] first staged row
    unsigned char *map = (unsigned char *)(t_off + SWEEP_TBMAX);   // [BCAP] row -> task
    unsigned char *flag = map + SWEEP_BCAP;    // [N] outcome per slot
    double *s_alpha = (double *)(flag + ((S.N + 15) & ~15));  // [M+1] staging table alpha_k
    __shared__ int s_scan[SWEEP_THREADS / 32];
    __shared__ unsigned long long s_bead;

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = S.M, N = S.N, dim = S.dim;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter);
    pimc_u4 di = pimc_draw_rk(st, &P.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P, di);
    if (P.kind[pick] != PIMC_UPD_RESHAPE_LINEAR) return;
    const UpdDev &U = T->upd[P.upd_id[pick]];
    const int var = (int)U.var[c], vmax = (int)P.vmax[pick];
    double *s_logtab = s_alpha + (M + 1);     // [2*128]
    for (int i = tid; i < 2 * PIMC_LOGTAB_N; i += SWEEP_THREADS) s_logtab[i] = S.logtab[i];
    const int j0 = 1 + (int)pimc_index(di.w[2], (uint32_t)M);
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    const int nx_stride = N; (void)nx_stride;
    const int *nextc = S.next + (size_t)c * N;
    if (tid == 0) s_bead = 0;
    for (int i = tid; i <= M; i += SWEEP_THREADS) s_alpha[i] = S.tab_alpha[i];
    unsigned long long my_beads = 0;

    for (int t0 = 0; t0 < N;) {
        // ---- batch selection: tasks t0.. as long as their rows (m + 1 each) fit the staging buffer ----
        int slot = t0 + tid, m = 0, cnt = 0;
        if (slot < N) {
            pimc_u4 dt = pimc_draw_rk(st, &P.rk, (uint32_t)slot, PIMC_K_TASK, 0, 0);
            int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
            m = vmax < mm ? vmax : mm;
            cnt = m + 1;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_scan[w];
        incl += wbase;
        const int excl = incl - cnt;
        const bool fits = slot < N && incl <= SWEEP_BCAP && tid < SWEEP_TBMAX;
        const int TB = __syncthreads_count(fits);      // prefix property: the fitting tasks are exactly the first TB
        if (fits) { t_m[tid] = m; t_off[tid] = excl; }
        __syncthreads();
        const int B = t_off[TB - 1] + t_m[TB - 1] + 1;
        if (tid < TB) {
            // endpoints (reshape.jl:56-58) with the boundary shift of levy! (helper.jl:120-125); map rows -> task
            const int n = slot, jm = j0 + m;
            const int pe = jm <= M ? n : nextc[n], je = (jm <= M ? jm : jm - M) - 1;
            double bx = S.r[RIDX(S, c, n, 0, j0 - 1)], ex = S.r[RIDX(S, c, pe, 0, je)];
            if (fabs(bx - ex) > L) ex += d_sign(bx) * twoL;
            xs[excl] = bx; xs[excl + m] = ex;
            if (dim > 1) {
                double by = S.r[RIDX(S, c, n, 1, j0 - 1)], ey = S.r[RIDX(S, c, pe, 1, je)];
                if (fabs(by - ey) > L) ey += d_sign(by) * twoL;
                ys[excl] = by; ys[excl + m] = ey;
            }
            for (int r = 0; r <= m; ++r) map[excl + r] = (unsigned char)tid;
            my_beads += (unsigned long long)(m - 1);
        }
        __syncthreads();
        // ---- phase A: Gaussians of every interior row, lanes = rows ----
        for (int s = tid; s < B; s += SWEEP_THREADS) {
            const int q = map[s], row = s - t_off[q], mq = t_m[q];
            if (row >= 1 && row < mq) {
                double g0, g1;
                pimc_gauss_pair_t(pimc_draw_rk(st, &P.rk, (uint32_t)(t0 + q), PIMC_K_BRIDGE, 0, (uint32_t)row), s_logtab, &g0, &g1);
                const double sig = S.tab_sig[mq + 1 - row];
                xs[s] = g0 * sig;
                if (dim > 1) ys[s] = g1 * sig;
            }
        }
        __syncthreads();
        // ---- phase B: serial recurrence r[j+1] = (alpha r[j] + (1-alpha) r[end]) + xi sigma, lanes = (task, dim) ----
        for (int w = tid; w < TB * dim; w += SWEEP_THREADS) {
            const int q = dim > 1 ? (w >> 1) : w, k = dim > 1 ? (w & 1) : 0;
            double *arr = k ? ys : xs;
            const int base = t_off[q], mq = t_m[q];
            double prev = arr[base];
            const double e = arr[base + mq];
            const double *al = s_alpha + mq + 1;  // alpha of row `row` is al[-row]
            double *ar = arr + base;
#pragma unroll 4
            for (int row = 1; row < mq; ++row) {
                const double a = al[-row];
                prev = a * prev + (1 - a) * e + ar[row];
                ar[row] = prev;
            }
        }
        __syncthreads();
        // ---- phase C: teleport every row (helper.jl:136-138), potential at the new positions ----
        for (int s = tid; s < B; s += SWEEP_THREADS) {
            double x = d_teleport_fast(xs[s], L, twoL, inv2L), y = 0.0;
            xs[s] = x;
            if (dim > 1) { y = d_teleport_fast(ys[s], L, twoL, inv2L); ys[s] = y; }
            if (POT != PIMC_POT_ZERO) pv[s] = d_pot_t<POT>(S.pot, x, y, dim);
        }
        __syncthreads();
        // ---- phase D: Delta-U, Metropolis, commit -- one warp per task ----
        {
            double *rc = S.r + (size_t)c * N * dim * M;   // this chain's positions / link cache: 32-bit indexing below
            double *vc = S.Vl + (size_t)c * N * M;
            for (int q = warp; q < TB; q += SWEEP_THREADS / 32) {
                const int n = t0 + q, mq = t_m[q], base = t_off[q], nx = nextc[n];
                // link jp (1-based) starts at slice j0-1+jp-1 of n, or wrapped on the next particle of the cycle
                const int first = j0 - 1, nfirst = M - first;           // rows 0..nfirst-1 stay on particle n
                double wi = 0.0, wu = 0.0;
                for (int jp = lane; jp < mq; jp += 32) {
                    const int p = jp < nfirst ? n : nx, sl = jp < nfirst ? first + jp : jp - nfirst;
                    wi += vc[p * M + sl];
                    wu += (POT == PIMC_POT_ZERO) ? mht * (0.0 + 0.0) : mht * (pv[base + jp] + pv[base + jp + 1]);
                }
                wi = 0.0 + warp_sum(wi); wu = 0.0 + warp_sum(wu);
                int acc = 0;
                if (lane == 0) {
                    const double dw = wu - wi;     // exp(dw) >= 1 for dw >= 0: accepted without the exponential or the uniform
                    if (dw >= 0.0) acc = 1;
                    else {
                        const double delta = pimc_exp(dw);
                        if (delta >= 1.0) acc = 1;
                        else { pimc_u4 dm = pimc_draw_rk(st, &P.rk, (uint32_t)n, PIMC_K_TASK, 0, 1); acc = delta > pimc_u01_co(dm.w[0], dm.w[1]); }
                    }
                    flag[n] = (unsigned char)acc;
                }
                acc = __shfl_sync(0xffffffffu, acc, 0);
                if (acc) {
                    for (int jp = lane; jp < mq; jp += 32) {
                        const int p = jp < nfirst ? n : nx, sl = jp < nfirst ? first + jp : jp - nfirst;
                        rc[(p * dim) * M + sl] = xs[base + jp];
                        if (dim > 1) rc[(p * dim + 1) * M + sl] = ys[base + jp];
                        vc[p * M + sl] = (POT == PIMC_POT_ZERO) ? mht * (0.0 + 0.0) : mht * (pv[base + jp] + pv[base + jp + 1]);
                    }
                }
            }
        }
        __syncthreads();
        t0 += TB;
    }
    if (my_beads) atomicAdd(&s_bead, my_beads);
    __syncthreads();
    if (tid == 0) d_bookkeep_sweep(U, c, flag, N, s_bead, P.stats);
}

// generic centre-of-mass proposal for a permutation cycle of several worldlines (rare in a sweep): kept out of line so that it
// does not cost the fast path registers
__device__ __noinline__ int d_com_cycle_generic(const DevSys *Sg, int c, int n, double maxd, pimc_stream st, int *npol)
{
    const DevSys &S = *Sg;
    pimc_u4 dm = pimc_draw(st, (uint32_t)n, PIMC_K_TASK, 0, 1);
    DSrc ds; ds.d = nullptr; ds.st = st; ds.slot = (uint32_t)n;
    return d_com_warp(S, c, n, maxd, ds, pimc_u01_co(dm.w[0], dm.w[1]), 1, nullptr, nullptr, npol);
}

// KM = ceil(M / 32) <= 8: the worldline lives in registers (KM beads per lane), one read and one write of HBM per bead.
template <int POT, int KM>
__global__ void __launch_bounds__(SWEEP_THREADS, (KM <= 4 ? 1024 : 768) / SWEEP_THREADS) k_com_sweep(DevSys S, const DevTables *__restrict__ T, SweepParams P)
{
    extern __shared__ double sm[];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = SWEEP_THREADS / 32;
    const int M = S.M, N = S.N, dim = S.dim;
    unsigned char *flag = (unsigned char *)sm;
    __shared__ unsigned long long s_bead;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter);
    pimc_u4 di = pimc_draw_rk(st, &P.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P, di);
    if (P.kind[pick] != PIMC_UPD_SINGLE_COM && P.kind[pick] != PIMC_UPD_POLYMER_COM) return;
    const UpdDev &U = T->upd[P.upd_id[pick]];
    const bool polymer = P.kind[pick] == PIMC_UPD_POLYMER_COM;
    const double maxd = U.var[c];
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    const int *nextc = S.next + (size_t)c * N;
    if (tid == 0) s_bead = 0;
    for (int i = tid; i < N; i += SWEEP_THREADS) flag[i] = 2;
    __syncthreads();
    unsigned long long my_beads = 0;
    // groups of 32 proposals per warp: lane l draws the displacement of the l-th proposal of the group (one Philox per lane
    // instead of one per warp and proposal); the proposals themselves run one after the other, lanes striding the slices
    for (int g0 = warp; g0 < N; g0 += NW * 32) {
        const int nl = g0 + lane * NW;
        double dxl = 0.0, dyl = 0.0;
        if (nl < N) {
            pimc_u4 w = pimc_draw_rk(st, &P.rk, (uint32_t)nl, PIMC_K_COM, 0, 0);
            dxl = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5);
            dyl = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        }
        for (int t = 0; t < 32; ++t) {
            const int n = g0 + t * NW;
            if (n >= N) break;
            const double dx = __shfl_sync(0xffffffffu, dxl, t), dy = __shfl_sync(0xffffffffu, dyl, t);
            const int nx = nextc[n];
            const bool single = nx == n;
            bool run_it = single;
            if (polymer && !single) { run_it = true; int p = nx, cnt = 0; while (p != n && cnt <= N) { if (p < n) run_it = false; p = nextc[p]; cnt++; } }
            if (!run_it) continue;
            if (!single) {
                int npol = 1;
                int r = d_com_cycle_generic(P.Sg, c, n, maxd, st, &npol);
                if (lane == 0) { flag[n] = r == 1 ? 1 : 0; my_beads += (unsigned long long)M * npol; }
                continue;
            }
            double *rx = S.r + RIDX(S, c, n, 0, 0), *ry = rx + M, *vl = S.Vl + VIDX(S, c, n, 0);
            double x[KM], y[KM], v[KM], wi = 0.0, wu = 0.0;
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                const int j = lane + 32 * k;
                x[k] = j < M ? rx[j] : 0.0;
                y[k] = (dim > 1 && j < M) ? ry[j] : 0.0;
                v[k] = j < M ? vl[j] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                wi += (lane + 32 * k < M) ? v[k] : 0.0;
                x[k] = d_teleport_fast(x[k] + dx, L, twoL, inv2L);
                if (dim > 1) y[k] = d_teleport_fast(y[k] + dy, L, twoL, inv2L);
                v[k] = (POT == PIMC_POT_ZERO) ? 0.0 : d_pot_t<POT>(S.pot, x[k], y[k], dim);
            }
            // link j -> j+1: the next bead's potential sits one lane up (lane 31: lane 0 of the next register); last bead -> bead 0
            const double v00 = __shfl_sync(0xffffffffu, v[0], 0);
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                const int j = lane + 32 * k;
                double up = 0.0;
                if (POT != PIMC_POT_ZERO) {
                    up = __shfl_down_sync(0xffffffffu, v[k], 1);
                    const double nextreg = __shfl_sync(0xffffffffu, (k + 1 < KM) ? v[(k + 1 < KM) ? k + 1 : k] : 0.0, 0);
                    if (lane == 31) up = nextreg;
                    if (j == M - 1) up = v00;
                }
                const double lk = mht * (v[k] + up);
                v[k] = lk;                      // v now holds the new link action
                if (j < M) wu += lk;
            }
            wi = 0.0 + warp_sum(wi); wu = 0.0 + warp_sum(wu);
            int acc = 0;
            if (lane == 0) {
                const double dw = wu - wi;
                if (dw >= 0.0) acc = 1;
                else {
                    const double delta = pimc_exp(dw);
                    if (delta >= 1.0) acc = 1;
                    else { pimc_u4 dm = pimc_draw_rk(st, &P.rk, (uint32_t)n, PIMC_K_TASK, 0, 1); acc = delta > pimc_u01_co(dm.w[0], dm.w[1]); }
                }
                flag[n] = (unsigned char)acc;
                my_beads += (unsigned long long)M;
            }
            acc = __shfl_sync(0xffffffffu, acc, 0);
            if (acc) {
#pragma unroll
                for (int k = 0; k < KM; ++k) {
                    const int j = lane + 32 * k;
                    if (j < M) { rx[j] = x[k]; if (dim > 1) ry[j] = y[k]; vl[j] = v[k]; }
                }
            }
        }
    }
    if (my_beads) atomicAdd(&s_bead, my_beads);
    __syncthreads();
    if (tid == 0) d_bookkeep_sweep(U, c, flag, N, s_bead, P.stats);
}

// the swap move stays one proposal per chain and iteration (reshape.jl:123-283), thread 0 of a one-warp CTA
__global__ void k_swap_iter(DevSys S, const DevTables *__restrict__ T, SweepParams P)
{
    const int c = blockIdx.x, N = S.N, M = S.M;
    if (threadIdx.x != 0) return;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter);
    pimc_u4 di = pimc_draw(st, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const UpdDev &U = T->upd[P.upd_id[d_pick_update(P, di)]];
    if (U.kind != PIMC_UPD_RESHAPE_SWAP) return;
    unsigned char f = 3; unsigned long long beads = 0;
    if (N > 1) {
        const int var = (int)U.var[c];
        pimc_u4 dt = pimc_draw(st, 0, PIMC_K_TASK, 0, 0);
        pimc_u4 dm = pimc_draw(st, 0, PIMC_K_TASK, 0, 1);
        pimc_u4 dsw = pimc_draw(st, 0, PIMC_K_SWAP, 0, 0);
        int j0 = 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
        int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
        int m = (int)U.vmax < mm ? (int)U.vmax : mm;
        int n1 = (int)pimc_index(dsw.w[0], (uint32_t)N);
        double *w = S.wtab + (size_t)c * N;
        d_swap_weights(S, c, n1, j0, m, w);
        double norm = w[0]; for (int i = 1; i < N; ++i) norm = norm + w[i];
        for (int i = 0; i < N; ++i) w[i] = w[i] / norm;
        int n2 = d_sample_weighted(w, N, pimc_u01_co(dsw.w[2], dsw.w[3]));
        if (n1 != n2) {
            GSrc g1, g2; g1.xi = nullptr; g1.st = st; g1.slot = 0; g1.kind = PIMC_K_BRIDGE; g1.tab = S.logtab; g2 = g1; g2.kind = PIMC_K_BRIDGE2;
            int r = d_reshape_swap(S, c, n1, n2, j0, m, g1, g2, pimc_u01_co(dm.w[0], dm.w[1]), 1, nullptr, nullptr);
            f = r == 1 ? 1 : 0; beads = 2ull * (unsigned long long)(m - 1);
        }
    }
    // faithful-style bookkeeping of a single proposal (apply!, simulation.jl:19-27)
    RingReg R; R.head = U.ring_head[c]; R.len = U.ring_len[c]; R.sum = U.ring_sum[c]; R.tries = U.tries_var[c];
    U.tries[c] += 1;
    if (f != 3) { U.accepted[c] += f; d_ring_push(U, c, R, f); }
    U.ring_head[c] = R.head; U.ring_len[c] = R.len; U.ring_sum[c] = R.sum; U.tries_var[c] = R.tries;
    U.bead_moves[c] += (long long)beads;
    if ((R.tries % U.adj) == 0) d_adjust(U, c, R);
    if (P.stats) { atomicAdd(P.stats + 0, 1ull); atomicAdd(P.stats + 2, beads); }
}

// measurement_Z_sector (measurement.jl:1-17) for every chain at one cadence hit; k = 0-based measurement index
struct MeasParams { int nen; int en_id[PIMC_MAXE]; int nde; int de_id[PIMC_MAXD]; long long k; };
// Energy functor (measurement.jl:92-122) with one warp per worldline (lanes stride the slices: coalesced, no index division)
template <int POT>
__device__ __forceinline__ void d_energy_block_fast(const DevSys &S, int c, double *red, double *E, double *Ev)
{
    const int M = S.M, N = S.N, dim = S.dim, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double L = S.L, twoL = 2 * S.L;
    const double *rc = S.r + (size_t)c * N * dim * M;
    const int *nextc = S.next + (size_t)c * N;
    double link = 0.0, pot = 0.0, vkin = 0.0;
    for (int n = warp; n < N; n += nw) {
        const double *rx = rc + (size_t)(n * dim) * M, *ry = rx + M;
        const int nx = nextc[n];
        const double *qx = rc + (size_t)(nx * dim) * M, *qy = qx + M;
        for (int j = lane; j < M; j += 32) {
            const double ax = rx[j], ay = dim > 1 ? ry[j] : 0.0;
            const double bx = j == M - 1 ? qx[0] : rx[j + 1], by = dim > 1 ? (j == M - 1 ? qy[0] : ry[j + 1]) : 0.0;
            double dx = fabs(ax - bx); { const double alt = twoL - dx; dx = alt < dx ? alt : dx; }
            double d2 = dx * dx;
            if (dim > 1) { double dy = fabs(ay - by); const double alt = twoL - dy; dy = alt < dy ? alt : dy; d2 = d2 + dy * dy; }
            link += d2;
            if (POT != PIMC_POT_ZERO) pot += d_pot_t<POT>(S.pot, ax, ay, dim) + d_pot_t<POT>(S.pot, bx, by, dim);
            vkin += d_rdv(S.pot, ax, ay, dim);
        }
    }
    (void)L;
    link = warp_sum(link); pot = warp_sum(pot); vkin = warp_sum(vkin);
    __syncthreads();
    if (lane == 0) { red[warp] = link; red[32 + warp] = pot; red[64 + warp] = vkin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        link = 0.0; pot = 0.0; vkin = 0.0;
        for (int i = 0; i < nw; ++i) { link += red[i]; pot += red[32 + i]; vkin += red[64 + i]; }
        *E = (double)(S.dim * S.N) / (2 * S.tau) - 1 / (4 * S.lambda * (S.tau * S.tau) * S.M) * link + 1.0 / (2 * S.M) * pot;
        *Ev = 1.0 / (2 * S.M) * vkin + 1.0 / (2 * S.M) * pot;
    }
}
template <int POT>
__global__ void __launch_bounds__(256) k_measure(DevSys S, const DevTables *__restrict__ T, MeasParams P)
{
    __shared__ double red[96];
    const int c = blockIdx.x;
    for (int e = 0; e < P.nen; ++e) {
        const EnDev &En = T->en[P.en_id[e]];
        double E, Ev;
        d_energy_block_fast<POT>(S, c, red, &E, &Ev);
        if (threadIdx.x == 0) {
            if (P.k < En.cap) { En.E[(size_t)P.k * S.C + c] = E; En.Ev[(size_t)P.k * S.C + c] = Ev; }
            double *a = En.acc + (size_t)c * 5;
            a[0] += 1.0; a[1] += E; a[2] += E * E; a[3] += Ev; a[4] += Ev * Ev;
        }
        __syncthreads();
    }
    for (int d = 0; d < P.nde; ++d) d_density_block(S, c, T->de[P.de_id[d]]);
}
