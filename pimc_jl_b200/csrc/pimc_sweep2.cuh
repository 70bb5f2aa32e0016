// pimc_sweep2.cuh -- second-generation staging sweep (ReshapeLinear for every worldline of a chain, reshape.jl:31-91).
//
// What bounds the first-generation kernel (pimc_sweep.cuh) is the serial recurrence of levy! (helper.jl:129-135): per batch it
// runs for max(m) dependent steps with two warps busy and six waiting at a barrier, and a chain needed three batches per sweep.
// Here the whole chain is ONE batch (512-thread CTA, two CTAs per SM, ~100 KB of staging each) and the staged rows are laid
// out step-major over tasks ranked by segment length:
//
//     slot(row, rank) = off[row] + rank,   rank 0 = longest segment,   off[row+1] - off[row] = alive(row) rounded up to odd
//
// so that   phase A (lanes = slots)        Philox -> Box-Muller -> xi*sigma, dense and conflict-free,
//           phase B (lanes = (rank, dim))  the recurrence reads and writes consecutive words of one row per step (conflict-free),
//           phase D (half-warp per task, lanes = rows) teleport, potential, Delta-U, Metropolis, coalesced commit; the odd row
//                                          stride keeps its strided shared-memory reads spread over the banks.
// Same draws and the same arithmetic per bead as generation one and as the oracle: trajectories are bit-identical.
#pragma once
#include "pimc_sweep.cuh"

#define SW2_THREADS 512       // largest CTA of this generation (array sizes); the kernel is templated on the actual count
#define SW2_MAXR 256          // rows of a segment: m + 1 <= M - 1 <= 255 (the batched path needs M <= 256)
#define SW2_KR 4              // rows per lane staged in registers by phase D (chunks of 16 * SW2_KR rows)


// smem carve-up shared by host (size) and device (pointers); cap = staged slots, a multiple of 16
__host__ __device__ inline int sw2_mp(int M) { return (M + 2) & ~1; }   // table stride: even, so that 16-byte async copies stay aligned
__host__ __device__ inline size_t sw2_smem_bytes(int pot_kind, int cap, int N, int M)
{
    size_t b = (size_t)(pot_kind == PIMC_POT_ZERO ? 2 : 3) * cap * sizeof(double);       // xs, ys, (pv)
    b += (size_t)(2 * sw2_mp(M) + 2 * PIMC_LOGTAB_N + SW2_THREADS) * sizeof(double);      // alpha, sigma, log table, s_wi
    b += (size_t)(2 * (SW2_MAXR + 2)) * sizeof(int);                                      // alive[], off[]
    b += (size_t)SW2_THREADS * sizeof(unsigned short) + SW2_THREADS;                      // r_task[], r_m[]
    b += (size_t)cap;                                                                     // rowof[]
    b += ((size_t)N + 15) & ~(size_t)15;                                                  // flag[]
    return b + 64;
}

template <int POT, int TH>
__device__ __forceinline__ void d_reshape_sweep2_body(const DevSys &S, const Sweep2Params &P2, const pimc_stream &st, const pimc_u4 &di, const int pick)
{
    extern __shared__ double sm[];
    constexpr int NW = TH / 32;
    const SweepParams &P = P2.sp;
    const UpdDev &U = P2.upd[pick];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double varf = U.var[c];                     // the only global load the segment lengths wait for: issued first
    const int M = S.M, N = S.N, dim = S.dim, cap = P2.cap, MP = sw2_mp(M);
    double *xs = sm, *ys = sm + cap;
    double *pv = sm + 2 * cap;                                                       // potential at the new rows, only when POT != 0
    double *s_alpha = (POT == PIMC_POT_ZERO) ? sm + 2 * cap : sm + 3 * cap;          // [MP] alpha_k
    double *s_sig = s_alpha + MP;                                                    // [MP] sigma_k
    double *s_logtab = s_sig + MP;                                                   // [2*128]
    double *s_wi = s_logtab + 2 * PIMC_LOGTAB_N;                                     // [TH] cached action of a task's links, by rank
    int *alive = (int *)(s_wi + TH);                                                 // [MAXR+2] histogram of m, then #tasks with m >= row
    int *off = alive + (SW2_MAXR + 2);                                               // [MAXR+2] first slot of a row
    unsigned short *r_task = (unsigned short *)(off + (SW2_MAXR + 2));               // [TH] rank -> task of the batch
    unsigned char *r_m = (unsigned char *)(r_task + TH);                             // [TH] rank -> m
    unsigned char *rowof = r_m + TH;                                                 // [cap] slot -> row
    unsigned char *flag = rowof + cap;                                               // [N] outcome per worldline
    __shared__ int s_scan[NW];
    __shared__ int s_first, s_B, s_mmax, s_rows;
    __shared__ unsigned long long s_bead;
    __shared__ BookPre s_pre;
#ifdef EXP_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64(), tstart = tlast; int nbatch = 0;
#endif

    // ---- prologue: nothing here waits for global memory ----
    for (int i = tid; i < SW2_MAXR + 2; i += TH) alive[i] = 0;
    if (tid == 0) { s_bead = 0; s_rows = 0; }
    if (tid == 32) d_book_prefetch_async1(U, c, &s_pre);   // the thread that issues a cp.async is the one that waits for it
    for (int i = tid; i < PIMC_LOGTAB_N; i += TH) d_cp_async16(s_logtab + 2 * i, S.logtab + 2 * i);
    for (int i = tid; i < (M + 1) / 2; i += TH) { d_cp_async16(s_alpha + 2 * i, S.tab_alpha + 2 * i); d_cp_async16(s_sig + 2 * i, S.tab_sig + 2 * i); }
    if (tid == TH - 1 && ((M + 1) & 1)) { d_cp_async8(s_alpha + M, S.tab_alpha + M); d_cp_async8(s_sig + M, S.tab_sig + M); }
    const int j0 = 1 + (int)pimc_index(di.w[2], (uint32_t)M);
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    const int *nextc = S.next + (size_t)c * N;
    double *rc = S.r + (size_t)c * N * dim * M;
    double *vc = S.Vl + (size_t)c * N * M;
    const int first = j0 - 1, nfirst = M - first;   // a strand's rows 0..nfirst-1 lie on its own particle, the rest on the next one
    unsigned long long my_beads = 0;
    __syncthreads();                                // histogram zeroed
    const int var = (int)varf, vmax = (int)P.vmax[pick];
    const int rowslack = (vmax < var + 1 ? vmax : var + 1) + 2;   // padding slots: at most one per row
    TICK(0);

    for (int sb0 = 0; sb0 < N; sb0 += TH) {        // super-batch: one task per thread (one pass for N <= 512)
        if (sb0 > 0) { for (int i = tid; i < SW2_MAXR + 2; i += TH) alive[i] = 0; if (tid == 0) s_rows = 0; __syncthreads(); }
        const int n = sb0 + tid;
        const int nsb = N - sb0 < TH ? N - sb0 : TH;
        int m = 0, cnt = 0, pos = 0;
        double bx = 0.0, by = 0.0, ex = 0.0, ey = 0.0;
        if (n < N) {
            pimc_u4 dt = pimc_draw_rk(st, &P.rk, (uint32_t)n, PIMC_K_TASK, 0, 0);
            const int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
            m = vmax < mm ? vmax : mm;
            cnt = m + 1;
            // endpoints (reshape.jl:56-58); the loads are consumed after the ranking below
            const int pe = m < nfirst ? n : nextc[n], je = m < nfirst ? first + m : m - nfirst;
            bx = rc[(n * dim) * M + first]; ex = rc[(pe * dim) * M + je];
            if (dim > 1) { by = rc[(n * dim + 1) * M + first]; ey = rc[(pe * dim + 1) * M + je]; }
            my_beads += (unsigned long long)(m - 1);
            pos = atomicAdd(&alive[m], 1);          // optimistic: the whole super-batch is one batch
        }
        { const int rw = __reduce_add_sync(0xffffffffu, cnt); if (lane == 0 && rw) atomicAdd(&s_rows, rw); }
        __syncthreads();
        const bool all_fit = s_rows + rowslack <= cap;
        int incl = 0;
        if (!all_fit) {                             // several batches: inclusive prefix of the row counts over the tasks
            incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            for (int w = 0; w < warp; ++w) incl += s_scan[w];
        }

        for (int b0 = 0; b0 < nsb;) {              // batches: consecutive tasks whose rows fit the staging buffer (normally all of them)
            bool fits = tid < nsb;
            int TB = nsb;
            if (!all_fit) {
                __syncthreads();
                if (tid == b0) s_first = incl - cnt;
                for (int i = tid; i < SW2_MAXR + 2; i += TH) alive[i] = 0;
                __syncthreads();
                fits = tid >= b0 && tid < nsb && incl - s_first + rowslack <= cap;
                TB = __syncthreads_count(fits);    // prefix property: the fitting tasks are b0 .. b0+TB-1
                if (fits) pos = atomicAdd(&alive[m], 1);
                __syncthreads();
            }
            // ---- rank the tasks by segment length (counting sort on m, longest first) ----
            if (warp == 0) {
                // alive[row] <- #tasks with m >= row (suffix sums of the histogram); off[row] <- first slot of the row (odd strides)
                constexpr int PER = (SW2_MAXR + 2 + 31) / 32;
                int h[PER], tot = 0;
#pragma unroll
                for (int i = PER - 1; i >= 0; --i) { const int r = lane * PER + i; tot += (r < SW2_MAXR + 2) ? alive[r] : 0; h[i] = tot; }
                int suf = tot;                                        // inclusive suffix over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
                suf -= tot;                                           // tasks counted by the lanes above
                int pre = 0, mmax = 0;
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int a = h[i] + suf; h[i] = a;
                    if (a > 0) mmax = lane * PER + i;
                    pre += a > 0 ? (a | 1) : 0;                       // odd stride: strided readers of phase D spread over the banks
                }
                int incl2 = pre;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl2, o); if (lane >= o) incl2 += v; }
                int run = incl2 - pre;
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int r = lane * PER + i;
                    if (r < SW2_MAXR + 2) { alive[r] = h[i]; off[r] = run; }
                    run += h[i] > 0 ? (h[i] | 1) : 0;
                }
                mmax = __reduce_max_sync(0xffffffffu, mmax);
                if (lane == 31) s_B = run;
                if (lane == 0) s_mmax = mmax;
            } else if (tid == 32 && sb0 == 0 && b0 == 0) {             // second step of the counter prefetch: the scalars have landed
                d_cp_async_wait_all(); d_book_prefetch_async2(U, c, &s_pre);
            }
            __syncthreads();
            const int B = s_B, mmax = s_mmax;
            if (fits) {
                const int rank = alive[m + 1] + pos;                  // tasks with a longer segment come first
                r_task[rank] = (unsigned short)(tid - b0); r_m[rank] = (unsigned char)m;
                // boundary shift of levy! (helper.jl:120-125)
                if (fabs(bx - ex) > L) ex += d_sign(bx) * twoL;
                if (dim > 1 && fabs(by - ey) > L) ey += d_sign(by) * twoL;
                xs[off[0] + rank] = bx; xs[off[m] + rank] = ex;
                if (dim > 1) { ys[off[0] + rank] = by; ys[off[m] + rank] = ey; }
            }
            for (int row = warp; row <= mmax; row += NW)
                for (int s = off[row] + lane, e = off[row + 1]; s < e; s += 32) rowof[s] = (unsigned char)row;
            d_cp_async_wait_all();                                    // tables (first batch only; no-op afterwards)
            __syncthreads();
            TICK(1);
            // ---- phase A: Gaussians of every interior row, lanes = slots ----
            for (int s = tid; s < B; s += TH) {
                const int row = rowof[s], rank = s - off[row];
                if (row >= 1 && rank < alive[row + 1]) {              // interior row of its task: row < m  <=>  the task is alive at row + 1
                    const int mq = r_m[rank];
                    double g0, g1;
                    pimc_gauss_pair_t(pimc_draw_rk(st, &P.rk, (uint32_t)(sb0 + b0 + r_task[rank]), PIMC_K_BRIDGE, 0, (uint32_t)row), s_logtab, &g0, &g1);
                    const double sig = s_sig[mq + 1 - row];
                    xs[s] = g0 * sig;
                    if (dim > 1) ys[s] = g1 * sig;
                }
            }
            __syncthreads();
            TICK(2);
            // ---- phase B: serial recurrence r[j+1] = (alpha r[j] + (1-alpha) r[end]) + xi sigma, lanes = (rank, dim);
            //      the warps it leaves idle sum the cached link actions of every task meanwhile (global loads hidden behind B) ----
            const int TBp = (TB + 31) & ~31, nBw = (TBp * dim) / 32;
            const bool helpers = NW - nBw >= 2;
            if (!helpers || warp < nBw) {
                for (int w = tid; w < TBp * dim; w += TH) {
                    const int k = w >= TBp ? 1 : 0, rank = w - k * TBp;
                    if (rank >= TB) continue;
                    double *arr = (k ? ys : xs) + rank;
                    const int mq = r_m[rank];
                    double prev = arr[off[0]];
                    const double e = arr[off[mq]];
                    const double *al = s_alpha + mq + 1;   // alpha of row `row` is al[-row]
                    // software pipelined by hand: the next step's alpha and xi*sigma (and the row offset one step further) are fetched
                    // before this step's store, so that only DMUL -> DADD -> DADD sits on the serial path
                    int o1 = off[1], o2 = off[2];
                    double a = al[-1], g = arr[o1];
                    for (int row = 1; row < mq; ++row) {
                        const int o3 = off[row + 2];                                 // row + 2 <= mmax + 1: defined
                        const double a_n = al[-(row + 1)], g_n = arr[o2];            // row + 1 <= mq: the end row / alpha_1 slots exist
                        const double t = (1 - a) * e;
                        prev = a * prev + t + g;
                        arr[o1] = prev;
                        a = a_n; g = g_n; o1 = o2; o2 = o3;
                    }
                }
            } else {
                const int hl = lane & 15, nH = (NW - nBw) * 2;
                for (int rk = (warp - nBw) * 2 + (lane >> 4); rk < ((TB + 1) & ~1); rk += nH) {
                    const int rank = rk < TB ? rk : TB - 1;
                    const int nq = sb0 + b0 + r_task[rank], mq = r_m[rank], nxq = nextc[nq];
                    const double wi = d_wi_halfwarp(vc + nq * M + first, vc + nxq * M - nfirst, mq < nfirst ? mq : nfirst, mq, hl);
                    if (hl == 0 && rk < TB) s_wi[rank] = wi;
                }
            }
            __syncthreads();
            TICK(3);
            // ---- phase D: teleport (helper.jl:136-138), potential, Delta-U, Metropolis, coalesced commit -- half a warp per task ----
            {
                const int hl = lane & 15, hq = tid >> 4;
                for (int rk = hq; rk < ((TB + 1) & ~1); rk += TH / 16) {   // both halves of a warp iterate together
                    const bool live = rk < TB;
                    const int rank = live ? rk : TB - 1;
                    const int nq = sb0 + b0 + r_task[rank], mq = live ? r_m[rank] : 0, nxq = nextc[nq];
                    const int mqw = max(mq, __shfl_xor_sync(0xffffffffu, mq, 16));   // warp-uniform trip count
                    const int n1 = mq < nfirst ? mq : nfirst;                 // links / rows 0..n1-1 on particle nq, the rest wrapped on nxq
                    double *x1 = rc + (nq * dim) * M + first, *x2 = rc + (nxq * dim) * M - nfirst;
                    double *w1 = vc + nq * M + first, *w2 = vc + nxq * M - nfirst;
                    double wi = helpers ? s_wi[rank] : d_wi_halfwarp(w1, w2, n1, mq, hl);
                    for (int base = 0; base <= mqw; base += 16 * SW2_KR) {    // rows staged through registers: loads, then arithmetic, then stores
                        double X[SW2_KR], Y[SW2_KR];
#pragma unroll
                        for (int i = 0; i < SW2_KR; ++i) {
                            const int jp = base + hl + 16 * i;
                            if (jp <= mq) { const int s = off[jp] + rank; X[i] = xs[s]; Y[i] = dim > 1 ? ys[s] : 0.0; }
                        }
#pragma unroll
                        for (int i = 0; i < SW2_KR; ++i) {
                            const int jp = base + hl + 16 * i;
                            if (jp <= mq) {
                                X[i] = d_teleport_fast(X[i], L, twoL, inv2L);
                                if (dim > 1) Y[i] = d_teleport_fast(Y[i], L, twoL, inv2L);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < SW2_KR; ++i) {
                            const int jp = base + hl + 16 * i;
                            if (jp <= mq) {
                                const int s = off[jp] + rank;
                                xs[s] = X[i]; if (dim > 1) ys[s] = Y[i];
                                if (POT != PIMC_POT_ZERO) pv[s] = d_pot_t<POT>(S.pot, X[i], Y[i], dim);
                            }
                        }
                    }
                    __syncwarp();
                    double wu = 0.0;
                    if (POT != PIMC_POT_ZERO)
                        for (int jp = hl; jp < mq; jp += 16) wu += mht * (pv[off[jp] + rank] + pv[off[jp + 1] + rank]);
                    else
                        for (int jp = hl; jp < mq; jp += 16) wu += mht * (0.0 + 0.0);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) wu += __shfl_xor_sync(0xffffffffu, wu, o);
                    wi = 0.0 + wi; wu = 0.0 + wu;
                    int acc = 0;
                    if (hl == 0 && live) {
                        const double dw = wu - wi;     // exp(dw) >= 1 for dw >= 0: accepted without the exponential or the uniform
                        if (dw >= 0.0) acc = 1;
                        else {
                            const double delta = pimc_exp(dw);
                            if (delta >= 1.0) acc = 1;
                            else { pimc_u4 dm = pimc_draw_rk(st, &P.rk, (uint32_t)nq, PIMC_K_TASK, 0, 1); acc = delta > pimc_u01_co(dm.w[0], dm.w[1]); }
                        }
                        flag[nq] = (unsigned char)acc;
                    }
                    acc = __shfl_sync(0xffffffffu, acc, lane & 16);
                    if (acc) {
                        for (int jp = hl; jp < mq; jp += 16) {
                            const int s = off[jp] + rank;
                            double *xd = jp < n1 ? x1 : x2, *wd = jp < n1 ? w1 : w2;
                            xd[jp] = xs[s];
                            if (dim > 1) xd[M + jp] = ys[s];
                            wd[jp] = (POT == PIMC_POT_ZERO) ? mht * (0.0 + 0.0) : mht * (pv[s] + pv[off[jp + 1] + rank]);
                        }
                    }
                }
            }
            __syncthreads();
            TICK(4);
#ifdef EXP_TIMING
            nbatch++;
#endif
            b0 += TB;
        }
    }
    {
        const unsigned wsum = __reduce_add_sync(0xffffffffu, (unsigned)my_beads);
        if (lane == 0 && wsum) atomicAdd(&s_bead, (unsigned long long)wsum);
    }
    if (tid == 32) d_cp_async_wait_all();          // ring words of the counter prefetch
    __syncthreads();
    if (warp == 0) d_bookkeep_sweep_warp(U, c, flag, N, s_bead, P.stats, s_pre);
#ifdef EXP_TIMING
    TICK(5);
    if (threadIdx.x == 0 && (blockIdx.x % 512) == 7 && (P.iter % 64) == 3)
        printf("gen2 blk %d iter %llu batches %d cycles: prologue %lld setup %lld A %lld B %lld D %lld book %lld total %lld\n", blockIdx.x, P.iter, nbatch, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], clock64() - tstart);
#endif
}

// One launch per iteration, one 512-thread CTA per chain: the CTA picks its update (simulation.jl:33-37) and runs that family's sweep.
template <int POT, int KM, int TH>
__global__ void __launch_bounds__(TH, 1024 / TH) k_sweep2(const __grid_constant__ DevSys S, const DevTables *__restrict__ T, const __grid_constant__ Sweep2Params P2)
{
    const int c = blockIdx.x;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P2.sp.iter);
    pimc_u4 di = pimc_draw_rk(st, &P2.sp.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P2.sp, di);
    const int kind = P2.sp.kind[pick];
    if (kind == PIMC_UPD_RESHAPE_LINEAR) d_reshape_sweep2_body<POT, TH>(S, P2, st, di, pick);
    else if (kind == PIMC_UPD_SINGLE_COM || kind == PIMC_UPD_POLYMER_COM) d_com_sweep_body<POT, KM, TH>(S, P2.upd[pick], P2.sp, st, pick);
}
