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
#include "pimc_sweep_util.cuh"

// staged rows per batch: runtime (Sweep2Params::cap), sized by the host from the shared-memory budget
#define SWEEP_TBMAX 256   // tasks per batch
#ifndef SWEEP_UNROLL_A
#define SWEEP_UNROLL_A 1  // rows per thread in flight in phase A (Gaussians); A/B knob, see profiles/
#endif
#ifndef SWEEP_UNROLL_C
#define SWEEP_UNROLL_C 1  // rows per thread in flight in phase C (teleport, potential)
#endif
static constexpr int kSweepUnrollA = SWEEP_UNROLL_A, kSweepUnrollC = SWEEP_UNROLL_C;
#ifdef EXP_TIMING
#define TICK(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); tacc[i] += t_ - tlast; tlast = t_; } } while (0)
#else
#define TICK(i) do { } while (0)
#endif

// cached action of the links of one strand, summed by half a warp in the order of the first-generation kernel
__device__ __forceinline__ double d_wi_halfwarp(const double *w1, const double *w2, int n1, int mq, int hl)
{
    double v[8], wi = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int jp = hl + 16 * i; v[i] = jp < mq ? (jp < n1 ? w1[jp] : w2[jp]) : 0.0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) wi += v[i];
    for (int jp = hl + 128; jp < mq; jp += 16) wi += jp < n1 ? w1[jp] : w2[jp];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) wi += __shfl_xor_sync(0xffffffffu, wi, o);
    return wi;
}

template <int POT>
__device__ __forceinline__ void d_reshape_sweep_body(const DevSys &S, const UpdDev &U, const SweepParams &P, const pimc_stream &st,
                                                     const pimc_u4 &di, const int pick, const int BCAP, const int c)
{
    extern __shared__ double sm[];
    double *xs = sm, *ys = sm + BCAP;          // BCAP staged rows per batch (a multiple of 16, sized by the host from the shared-memory budget)
    double *pv = sm + 2 * BCAP;                                                       // potential at the new rows, only when POT != 0
    double *s_wi = (POT == PIMC_POT_ZERO) ? sm + 2 * BCAP : sm + 3 * BCAP;            // [THREADS] cached action of a task's links (w_initial)
    int *t_m = (int *)(s_wi + SWEEP_THREADS);  // [THREADS] links of the batch's tasks
    int *t_off = t_m + SWEEP_THREADS;          // [THREADS] first staged row
    unsigned char *map = (unsigned char *)(t_off + SWEEP_THREADS);   // [BCAP] row -> task of the batch
    unsigned char *flag = map + BCAP;          // [N] outcome per slot
    double *s_logtab = (double *)(flag + ((S.N + 15) & ~15));  // [2*128] log table of the Gaussian transform (16-byte aligned)
    double *s_alpha = s_logtab + 2 * PIMC_LOGTAB_N;            // [M+1] staging table alpha_k
    __shared__ int s_scan[SWEEP_THREADS / 32];
    __shared__ int s_first;
    __shared__ unsigned long long s_bead;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = S.M, N = S.N, dim = S.dim;
    const double varf = U.var[c];             // the only global load the segment lengths wait for: issued first
    // counters of the apply! bookkeeping and the tables arrive by cp.async (no registers, nobody waits): the scalars now, the ring
    // words they point at after the first barrier; the thread that issues a cp.async is the one that waits for it
    __shared__ BookPre s_pre; if (tid == 32) d_book_prefetch_async1(U, c, &s_pre);
    for (int i = tid; i < PIMC_LOGTAB_N; i += SWEEP_THREADS) d_cp_async16(s_logtab + 2 * i, S.logtab + 2 * i);
    for (int i = tid; i < (M + 1) / 2; i += SWEEP_THREADS) d_cp_async16(s_alpha + 2 * i, S.tab_alpha + 2 * i);
    if (tid == SWEEP_THREADS - 1 && ((M + 1) & 1)) d_cp_async8(s_alpha + M, S.tab_alpha + M);
    const int vmax = (int)P.vmax[pick];
    const int j0 = 1 + (int)pimc_index(di.w[2], (uint32_t)M);
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    const int *nextc = S.next + (size_t)c * N;
    double *rc = S.r + (size_t)c * N * dim * M;   // this chain's positions / link cache: 32-bit indexing below
    double *vc = S.Vl + (size_t)c * N * M;
    const int first = j0 - 1, nfirst = M - first;   // a strand's rows 0..nfirst-1 lie on its own particle, the rest on the next one
    if (tid == 0) s_bead = 0;
    unsigned long long my_beads = 0;
    bool tables_pending = true;
#ifdef EXP_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64(), tstart = tlast; int nbatch = 0;
#endif

    // super-batch: SWEEP_THREADS tasks, one per thread; their segment length and endpoints are fetched once (one Philox draw,
    // one round of global loads) and stay in registers until the task's batch comes up
    for (int sb0 = 0; sb0 < N; sb0 += SWEEP_THREADS) {
        const int n = sb0 + tid;
        int m = 0, cnt = 0, nx = 0;
        double bx = 0.0, by = 0.0, ex = 0.0, ey = 0.0;
        if (n < N) {
            nx = nextc[n];                          // independent of the adaptive variable: in flight together with it
            bx = rc[(n * dim) * M + first]; if (dim > 1) by = rc[(n * dim + 1) * M + first];
            pimc_u4 dt = pimc_draw_rk(st, &P.rk, (uint32_t)n, PIMC_K_TASK, 0, 0);
            const int var = (int)varf;
            int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
            m = vmax < mm ? vmax : mm;
            cnt = m + 1;
            // endpoints (reshape.jl:56-58) with the boundary shift of levy! (helper.jl:120-125)
            const int pe = m < nfirst ? n : nx, je = m < nfirst ? first + m : m - nfirst;
            ex = rc[(pe * dim) * M + je];
            if (fabs(bx - ex) > L) ex += d_sign(bx) * twoL;
            if (dim > 1) {
                ey = rc[(pe * dim + 1) * M + je];
                if (fabs(by - ey) > L) ey += d_sign(by) * twoL;
            }
            my_beads += (unsigned long long)(m - 1);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        __syncthreads();                       // previous super-batch done with s_scan / staging
        if (lane == 31) s_scan[warp] = incl;
        if (tables_pending) {                  // first pass: the async copies have had the prologue to land
            d_cp_async_wait_all();
            if (tid == 32) { d_book_prefetch_async2(U, c, &s_pre); }
            tables_pending = false;
        }
        __syncthreads();
        for (int w = 0; w < warp; ++w) incl += s_scan[w];
        const int nsb = N - sb0 < SWEEP_THREADS ? N - sb0 : SWEEP_THREADS;

        for (int b0 = 0; b0 < nsb;) {          // batches: consecutive tasks whose rows (m + 1 each) fit the staging buffer
            TICK(7);
            if (tid == b0) s_first = incl - cnt;
            __syncthreads();
            const int basep = s_first;
            const bool fits = tid >= b0 && tid < nsb && incl - basep <= BCAP;
            const int TB = __syncthreads_count(fits);      // prefix property: the fitting tasks are b0 .. b0+TB-1
            const int q_me = tid - b0, excl = incl - cnt - basep;
            if (fits) {
                t_m[q_me] = m; t_off[q_me] = excl;
                xs[excl] = bx; xs[excl + m] = ex;
                if (dim > 1) { ys[excl] = by; ys[excl + m] = ey; }
                for (int r = 0; r <= m; ++r) map[excl + r] = (unsigned char)q_me;
            }
            __syncthreads();
            const int B = t_off[TB - 1] + t_m[TB - 1] + 1;
            TICK(0);
            // ---- phase A: Gaussians of every interior row, lanes = rows ----
#ifdef EXP_SKIP_A
            for (int s = tid; s < B; s += SWEEP_THREADS) { const int q = map[s], row = s - t_off[q], mq = t_m[q]; if (row >= 1 && row < mq) { xs[s] = 0.0; ys[s] = 0.0; } }
#else
#pragma unroll kSweepUnrollA
            for (int s = tid; s < B; s += SWEEP_THREADS) {
                const int q = map[s], row = s - t_off[q], mq = t_m[q];
                if (row >= 1 && row < mq) {
                    double g0, g1;
                    pimc_gauss_pair_t(pimc_draw_rk(st, &P.rk, (uint32_t)(sb0 + b0 + q), PIMC_K_BRIDGE, 0, (uint32_t)row), s_logtab, &g0, &g1);
                    const double sig = S.tab_sig[mq + 1 - row];
                    xs[s] = g0 * sig;
                    if (dim > 1) ys[s] = g1 * sig;
                }
            }
#endif
            __syncthreads();
            TICK(1);
            // ---- phase B: serial recurrence r[j+1] = (alpha r[j] + (1-alpha) r[end]) + xi sigma, lanes = (task, dim);
            //      the warps it leaves idle sum the cached link actions of the tasks meanwhile (global loads hidden behind B) ----
            const int nBw = (TB * dim + 31) >> 5;
            const bool helpers = SWEEP_THREADS / 32 - nBw >= 2;
            if (helpers && warp >= nBw) {
                const int hl = lane & 15, nH = (SWEEP_THREADS / 32 - nBw) * 2;
                for (int q = (warp - nBw) * 2 + (lane >> 4); q < ((TB + 1) & ~1); q += nH) {
                    const int qq = q < TB ? q : TB - 1;
                    const int nq = sb0 + b0 + qq, mq = t_m[qq], nxq = nextc[nq];
                    const double wi = d_wi_halfwarp(vc + nq * M + first, vc + nxq * M - nfirst, mq < nfirst ? mq : nfirst, mq, hl);
                    if (hl == 0 && q < TB) s_wi[qq] = wi;
                }
            }
#ifndef EXP_SKIP_B
            for (int w = tid; w < TB * dim; w += SWEEP_THREADS) {
                const int q = dim > 1 ? (w >> 1) : w, k = dim > 1 ? (w & 1) : 0;
                double *arr = k ? ys : xs;
                const int base = t_off[q], mq = t_m[q];
                double prev = arr[base];
                const double e = arr[base + mq];
                const double *al = s_alpha + mq + 1;  // alpha of row `row` is al[-row]
                double *ar = arr + base;
                // software pipelined by hand: next step's alpha and xi*sigma are fetched before this step's store, so that only
                // DMUL -> DADD -> DADD sits on the serial path (the compiler cannot hoist the loads over the aliasing store)
                double a = al[-1], g = ar[1];
                // (folding the teleport of every row into this loop was measured: -10 % -- the recurrence runs on TB * dim lanes only, and the
                //  extra dozen fp64 instructions per step lengthen the serial phase by more than the parallel pass below costs)
                for (int row = 1; row < mq; ++row) {
                    const double a_n = al[-(row + 1)], g_n = ar[row + 1];   // row + 1 <= mq: the end row / alpha_1 slots exist
                    const double t = (1 - a) * e;
                    prev = a * prev + t + g;
                    ar[row] = prev;
                    a = a_n; g = g_n;
                }
            }
#endif
            __syncthreads();
            TICK(2);
            // ---- phase C: teleport every row (helper.jl:136-138), potential at the new positions ----
#ifndef EXP_SKIP_C
#pragma unroll kSweepUnrollC
            for (int s = tid; s < B; s += SWEEP_THREADS) {
                double x = d_teleport_fast(xs[s], L, twoL, inv2L), y = 0.0;
                xs[s] = x;
                if (dim > 1) { y = d_teleport_fast(ys[s], L, twoL, inv2L); ys[s] = y; }
                if (POT != PIMC_POT_ZERO) pv[s] = d_pot_t<POT>(S.pot, x, y, dim);
            }
#endif
            __syncthreads();
            TICK(3);
            // ---- phase D: Delta-U from shared memory, Metropolis, coalesced commit -- one warp per task ----
            {   // half a warp per task: twice as many tasks in flight per phase, 4-step shuffle reductions
                const int hl = lane & 15, hq = tid >> 4;
                const unsigned hmask = (lane & 16) ? 0xFFFF0000u : 0x0000FFFFu;
                for (int q = hq; q < ((TB + 1) & ~1); q += SWEEP_THREADS / 16) {   // both halves of a warp iterate together
                    const bool live = q < TB;
                    const int qq = live ? q : TB - 1;
                    const int nq = sb0 + b0 + qq, mq = live ? t_m[qq] : 0, base = t_off[qq], nxq = nextc[nq];
                    double wi = helpers ? s_wi[qq] : d_wi_halfwarp(vc + nq * M + first, vc + nxq * M - nfirst, mq < nfirst ? mq : nfirst, mq, hl);
                    double wu = 0.0;
                    for (int jp = hl; jp < mq; jp += 16)
                        wu += (POT == PIMC_POT_ZERO) ? mht * (0.0 + 0.0) : mht * (pv[base + jp] + pv[base + jp + 1]);
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
                    (void)hmask;
#ifdef EXP_SKIP_D
                    acc = 0;
#endif
                    if (acc) {
                        const int n1 = mq < nfirst ? mq : nfirst;                 // rows 0..n1-1 on particle nq, the rest wrapped on nxq
                        double *x1 = rc + (nq * dim) * M + first, *x2 = rc + (nxq * dim) * M - nfirst;
                        double *w1 = vc + nq * M + first, *w2 = vc + nxq * M - nfirst;
                        for (int jp = hl; jp < n1; jp += 16) {
                            x1[jp] = xs[base + jp];
                            if (dim > 1) x1[M + jp] = ys[base + jp];
                            w1[jp] = (POT == PIMC_POT_ZERO) ? mht * (0.0 + 0.0) : mht * (pv[base + jp] + pv[base + jp + 1]);
                        }
                        for (int jp = nfirst + hl; jp < mq; jp += 16) {
                            x2[jp] = xs[base + jp];
                            if (dim > 1) x2[M + jp] = ys[base + jp];
                            w2[jp] = (POT == PIMC_POT_ZERO) ? mht * (0.0 + 0.0) : mht * (pv[base + jp] + pv[base + jp + 1]);
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
    {   // one shared-memory atomic per warp (a 64-bit ATOMS on one address from every thread costs ~8k cycles)
        const unsigned wsum = __reduce_add_sync(0xffffffffu, (unsigned)my_beads);
        if (lane == 0 && wsum) atomicAdd(&s_bead, (unsigned long long)wsum);
    }
    if (tid == 32) d_cp_async_wait_all();      // ring words of the counter prefetch
    __syncthreads();
    if (warp == 0) d_bookkeep_sweep_warp(U, c, flag, N, s_bead, P.stats, s_pre);
#ifdef EXP_TIMING
    TICK(5);
    if (threadIdx.x == 0 && (blockIdx.x % 512) == 7 && (P.iter % 64) == 3)
        printf("blk %d iter %llu batches %d cycles: setup %lld A %lld B %lld C %lld D %lld book %lld total %lld\n", blockIdx.x, P.iter, nbatch, tacc[0] + tacc[7], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], clock64() - tstart);
#endif
}

// ---- 1-D bulk async copies (TMA, cp.async.bulk + mbarrier complete_tx): HBM -> shared memory without register staging ----
__device__ __forceinline__ void d_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void d_mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void d_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void d_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void d_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}

// One member worldline of a permutation cycle under the displacement (dx, dy) (com.jl:47-100, move_polymer! helper.jl:368-395):
// loads its rows into registers, returns the lane-partial sums of the cached (wi) and new (wu) link actions; on return x, y hold
// the shifted, wrapped positions and v the new link actions.  (fx, fy): first bead of the NEXT member of the cycle (unshifted).
template <int POT, int KM>
__device__ __forceinline__ void d_com_member(const DevSys &S, const double *rx, const double *ry, const double *vl, double dx, double dy, double fx, double fy,
                                             double *x, double *y, double *v, double &wi, double &wu)
{
    const int lane = threadIdx.x & 31, M = S.M, dim = S.dim;
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
        const int j = lane + 32 * k;
        x[k] = j < M ? rx[j] : 0.0; y[k] = (dim > 1 && j < M) ? ry[j] : 0.0; v[k] = j < M ? vl[j] : 0.0;
    }
    wi = 0.0; wu = 0.0;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
        wi += (lane + 32 * k < M) ? v[k] : 0.0;
        x[k] = d_teleport_fast(x[k] + dx, L, twoL, inv2L);
        if (dim > 1) y[k] = d_teleport_fast(y[k] + dy, L, twoL, inv2L);
        v[k] = (POT == PIMC_POT_ZERO) ? 0.0 : d_pot_t<POT>(S.pot, x[k], y[k], dim);
    }
    const double vnext = (POT == PIMC_POT_ZERO) ? 0.0
        : d_pot_t<POT>(S.pot, d_teleport_fast(fx + dx, L, twoL, inv2L), dim > 1 ? d_teleport_fast(fy + dy, L, twoL, inv2L) : 0.0, dim);
#pragma unroll
    for (int k = 0; k < KM; ++k) {
        const int j = lane + 32 * k;
        double up = 0.0;
        if (POT != PIMC_POT_ZERO) {
            up = __shfl_down_sync(0xffffffffu, v[k], 1);
            const double nextreg = __shfl_sync(0xffffffffu, (k + 1 < KM) ? v[(k + 1 < KM) ? k + 1 : k] : 0.0, 0);
            if (lane == 31) up = nextreg;
            if (j == M - 1) up = vnext;         // the link out of the last slice ends on the next member of the cycle
        }
        const double lk = mht * (v[k] + up);
        v[k] = lk;
        if (j < M) wu += lk;
    }
}

// PolymerCenterOfMass for the exchange cycles of a chain (com.jl:31-104 read as "whole cycle"), all warps of the CTA together:
// leaders (smallest index of a cycle) by pointer jumping in shared memory, one warp per MEMBER for the link-action sums, one thread
// per cycle for the sums in cycle order + Metropolis, one warp per member again for the commit.  Same draws as the one-warp-per-cycle
// implementation d_com_warp (slot = leader); the reduction order differs (per-member warp sums, then cycle order).
template <int POT, int KM, int TH>
__device__ __forceinline__ void d_pcom_cycles(const DevSys &S, const SweepParams &P, const pimc_stream &st, const int c, const double maxd,
                                              unsigned char *flag, char *scratch, unsigned long long &my_beads)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = TH / 32, M = S.M, N = S.N, dim = S.dim;
    double *s_wi = (double *)scratch, *s_wu = s_wi + N, *s_fx = s_wu + N, *s_fy = s_fx + N;
    int *s_nx = (int *)(s_fy + N), *nA = s_nx + N, *nB = nA + N, *mA = nB + N, *mB = mA + N;
    unsigned char *s_acc = (unsigned char *)(mB + N);
    const int *nextc = S.next + (size_t)c * N;
    int any = 0;
    for (int i = tid; i < N; i += TH) { const int nx = nextc[i]; s_nx[i] = nx; nA[i] = nx; mA[i] = i; s_acc[i] = 0; any |= nx != i; }
    if (!__syncthreads_or(any)) return;                       // no exchange cycle in this chain
    int *ns = nA, *nd = nB, *ms = mA, *md = mB;
    for (int span = 1; span < N; span <<= 1) {                // after k rounds ms[i] = min over the 2^k successors of i
        for (int i = tid; i < N; i += TH) { const int j = ns[i]; md[i] = ms[i] < ms[j] ? ms[i] : ms[j]; nd[i] = ns[j]; }
        __syncthreads();
        int *t = ns; ns = nd; nd = t; t = ms; ms = md; md = t;
    }
    const double *rc = S.r + (size_t)c * N * dim * M, *vc = S.Vl + (size_t)c * N * M;
    for (int p = warp; p < N; p += NW) {                      // pass 1: link-action sums of every member
        const int pn = s_nx[p];
        if (pn == p) continue;
        const pimc_u4 w = pimc_draw_rk(st, &P.rk, (uint32_t)ms[p], PIMC_K_COM, 0, 0);
        const double dx = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5), dy = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        const double fx = rc[(size_t)(pn * dim) * M], fy = dim > 1 ? rc[(size_t)(pn * dim + 1) * M] : 0.0;
        double x[KM], y[KM], v[KM], wi, wu;
        d_com_member<POT, KM>(S, rc + (size_t)(p * dim) * M, rc + (size_t)(p * dim + 1) * M, vc + (size_t)p * M, dx, dy, fx, fy, x, y, v, wi, wu);
        wi = warp_sum(wi); wu = warp_sum(wu);
        if (lane == 0) { s_wi[p] = wi; s_wu[p] = wu; s_fx[p] = fx; s_fy[p] = fy; }
    }
    __syncthreads();
    for (int n = tid; n < N; n += TH)                         // one thread per cycle: sums in cycle order, Metropolis (helper.jl:3-5)
        if (ms[n] == n && s_nx[n] != n) {
            double wi = 0.0, wu = 0.0; int p = n, npol = 0;
            do { wi += s_wi[p]; wu += s_wu[p]; p = s_nx[p]; npol++; } while (p != n && npol <= N);
            const double dw = (0.0 + wu) - (0.0 + wi);
            int acc = 0;
            if (dw >= 0.0) acc = 1;
            else {
                const double delta = pimc_exp(dw);
                if (delta >= 1.0) acc = 1;
                else { pimc_u4 dm = pimc_draw_rk(st, &P.rk, (uint32_t)n, PIMC_K_TASK, 0, 1); acc = delta > pimc_u01_co(dm.w[0], dm.w[1]); }
            }
            s_acc[n] = (unsigned char)acc; flag[n] = (unsigned char)acc;
            my_beads += (unsigned long long)M * npol;
        }
    __syncthreads();
    double *rw = S.r + (size_t)c * N * dim * M, *vw = S.Vl + (size_t)c * N * M;
    for (int p = warp; p < N; p += NW) {                      // pass 2: commit the members of the accepted cycles
        if (s_nx[p] == p || !s_acc[ms[p]]) continue;
        const pimc_u4 w = pimc_draw_rk(st, &P.rk, (uint32_t)ms[p], PIMC_K_COM, 0, 0);
        const double dx = maxd * 2 * (pimc_u01_co(w.w[0], w.w[1]) - 0.5), dy = maxd * 2 * (pimc_u01_co(w.w[2], w.w[3]) - 0.5);
        double x[KM], y[KM], v[KM], wi, wu;
        double *rx = rw + (size_t)(p * dim) * M, *ry = rw + (size_t)(p * dim + 1) * M, *vl = vw + (size_t)p * M;
        d_com_member<POT, KM>(S, rx, ry, vl, dx, dy, s_fx[p], s_fy[p], x, y, v, wi, wu);
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            if (j < M) { rx[j] = x[k]; if (dim > 1) ry[j] = y[k]; vl[j] = v[k]; }
        }
    }
}

// KM = ceil(M / 32) <= 8: the worldline lives in registers (KM beads per lane), one read and one write of HBM per bead.
// FUSE: this launch belongs to a measurement iteration -- the Energy functor (measurement.jl:92-122) of the chain is accumulated on the
// fly from the final rows every proposal already holds in registers (one read of HBM serves the move and the estimator); chains with
// exchange cycles (whose members this sweep does not stream) leave mdone[c] = 0 and are measured by k_measure.
template <int POT, int KM, bool FUSE, int TH = SWEEP_THREADS>
__device__ __forceinline__ void d_com_sweep_body(const DevSys &S, const UpdDev &U, const SweepParams &P, const pimc_stream &st,
                                                 const int pick, const Sweep2Params &P2, const DevTables *__restrict__ T, const int c)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = TH / 32;
    const int M = S.M, N = S.N, dim = S.dim;
    unsigned char *flag = (unsigned char *)sm;
    __shared__ unsigned long long s_bead;
    __shared__ BookPre s_pre; if (tid == 0) s_pre = d_book_prefetch(U, c);
    const bool polymer = P.kind[pick] == PIMC_UPD_POLYMER_COM;
    const double maxd = U.var[c];
    const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
    const int *nextc = S.next + (size_t)c * N;
    if (tid == 0) s_bead = 0;
    for (int i = tid; i < N; i += TH) flag[i] = 2;
    // TMA pipeline of this warp: two stages of (x row, y row, link-action row); the next proposal's worldline is in flight
    // (cp.async.bulk -> shared memory, completion on an mbarrier) while the current one is being processed
    const bool use_tma = P.com_stage_off > 0;
    const int rows = dim + 1;
    const uint32_t row_bytes = (uint32_t)M * 8u;
    unsigned long long *mbar = (unsigned long long *)((char *)sm + P.com_stage_off) + warp * 2;
    double *stage0 = (double *)((char *)sm + P.com_stage_off + NW * 16) + (size_t)warp * 2 * rows * M;   // after the 2 * NW mbarriers
    const uint32_t bar_u = d_smem_u32(mbar), stage_u = d_smem_u32(stage0);
    const double *rc0 = S.r + (size_t)c * N * dim * M, *vc0 = S.Vl + (size_t)c * N * M;
    auto issue = [&](int n_, int stg) {
        if (lane == 0) {
            const uint32_t b = bar_u + 8u * stg, d0 = stage_u + (uint32_t)stg * rows * row_bytes;
            d_mbar_expect_tx(b, rows * row_bytes);
            d_bulk_g2s(d0, rc0 + (size_t)(n_ * dim) * M, (uint32_t)dim * row_bytes, b);      // x row (and y row: contiguous)
            d_bulk_g2s(d0 + (uint32_t)dim * row_bytes, vc0 + (size_t)n_ * M, row_bytes, b);  // cached link actions
        }
    };
    if (use_tma && lane == 0) { d_mbar_init(bar_u, 1); d_mbar_init(bar_u + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    int it = 0;
    if (use_tma && warp < N) issue(warp, 0);
#if COM_PREFETCH_NEXT
    int nx_pf = warp < N ? nextc[warp] : 0;
#endif
    unsigned long long my_beads = 0;
    double e_link = 0.0, e_pot = 0.0, e_vkin = 0.0; int e_cyc = 0;   // FUSE: lane partials of the Energy sums of this warp's worldlines
    const int dvk = S.pot.dv_kind;
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
            const int stg = it & 1; const uint32_t par = (uint32_t)(it >> 1) & 1u; ++it;
            if (use_tma) {
                __syncwarp();                              // every lane is done with the other stage (read two proposals ago)
                if (n + NW < N) issue(n + NW, stg ^ 1);
                d_mbar_wait(bar_u + 8u * stg, par);        // this proposal's rows have landed
            }
#if COM_PREFETCH_NEXT
            const int nx = nx_pf;
            nx_pf = n + NW < N ? nextc[n + NW] : 0;        // the next proposal's permutation entry: in flight behind this proposal's arithmetic
#else
            const int nx = nextc[n];
#endif
            const bool single = nx == n;
            if (!single) { if (FUSE) e_cyc = 1; continue; }   // members of exchange cycles: PolymerCOM moves them below, SingleCOM never (com.jl:139-141)
            double *rx = S.r + RIDX(S, c, n, 0, 0), *ry = rx + M, *vl = S.Vl + VIDX(S, c, n, 0);
            double x[KM], y[KM], v[KM], wi = 0.0, wu = 0.0;
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                const int j = lane + 32 * k;
                if (use_tma) {
                    const double *sx = stage0 + (size_t)stg * rows * M;
                    x[k] = j < M ? sx[j] : 0.0;
                    y[k] = (dim > 1 && j < M) ? sx[M + j] : 0.0;
                    v[k] = j < M ? sx[dim * M + j] : 0.0;
                } else {
                    x[k] = j < M ? rx[j] : 0.0;
                    y[k] = (dim > 1 && j < M) ? ry[j] : 0.0;
                    v[k] = j < M ? vl[j] : 0.0;
                }
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
            if (FUSE) {
                if (!acc) {                      // rejected: the worldline stays where it was (rows still staged / cached)
#pragma unroll
                    for (int k = 0; k < KM; ++k) {
                        const int j = lane + 32 * k;
                        const double *sx = stage0 + (size_t)stg * rows * M;
                        x[k] = j < M ? (use_tma ? sx[j] : rx[j]) : 0.0;
                        y[k] = (dim > 1 && j < M) ? (use_tma ? sx[M + j] : ry[j]) : 0.0;
                    }
                }
                const double x00 = __shfl_sync(0xffffffffu, x[0], 0), y00 = __shfl_sync(0xffffffffu, y[0], 0);   // closed worldline: last link ends on bead 0
#pragma unroll
                for (int k = 0; k < KM; ++k) {
                    const int j = lane + 32 * k;
                    double bx = __shfl_down_sync(0xffffffffu, x[k], 1), by = __shfl_down_sync(0xffffffffu, y[k], 1);
                    const double nbx = __shfl_sync(0xffffffffu, x[(k + 1 < KM) ? k + 1 : k], 0), nby = __shfl_sync(0xffffffffu, y[(k + 1 < KM) ? k + 1 : k], 0);
                    if (lane == 31) { bx = nbx; by = nby; }
                    if (j == M - 1) { bx = x00; by = y00; }
                    if (j < M) {
                        const double ax = x[k], ay = y[k];
                        double ddx = fabs(ax - bx); { const double alt = twoL - ddx; ddx = alt < ddx ? alt : ddx; }
                        double d2 = ddx * ddx;
                        if (dim > 1) { double ddy = fabs(ay - by); const double alt = twoL - ddy; ddy = alt < ddy ? alt : ddy; d2 = d2 + ddy * ddy; }
                        e_link += d2;
                        if (dvk == PIMC_DV_IDENTITY) { double q = ax * ax; if (dim > 1) q = q + ay * ay; e_vkin += q; }   // r . dV(r), measurement.jl:105
                        else if (dvk != PIMC_DV_ZERO) e_vkin += d_rdv(S.pot, ax, ay, dim);
                    }
                }
                // sum over the links of V(a) + V(b) = (sum of the link actions) / (-tau / 2): the cached ones if rejected, the new ones if accepted
                if (POT != PIMC_POT_ZERO && lane == 0) e_pot += (acc ? wu : wi) / mht;
            }
        }
    }
    if (FUSE) {
        __shared__ double s_en[3 * (TH / 32)];
        e_link = warp_sum(e_link); e_pot = warp_sum(e_pot); e_vkin = warp_sum(e_vkin);
        const int anycyc = __syncthreads_or(e_cyc);
        if (lane == 0) { s_en[warp] = e_link; s_en[NW + warp] = e_pot; s_en[2 * NW + warp] = e_vkin; }
        __syncthreads();
        if (tid == 0 && !anycyc) {
            double link = 0.0, pot = 0.0, vkin = 0.0;
            for (int i = 0; i < NW; ++i) { link += s_en[i]; pot += s_en[NW + i]; vkin += s_en[2 * NW + i]; }
            const double E = (double)(S.dim * S.N) / (2 * S.tau) - 1 / (4 * S.lambda * (S.tau * S.tau) * S.M) * link + 1.0 / (2 * S.M) * pot;
            const double Ev = 1.0 / (2 * S.M) * vkin + 1.0 / (2 * S.M) * pot;
            for (int e = 0; e < P2.mp.nen; ++e) {
                const EnDev &En = T->en[P2.mp.en_id[e]];
                const long long k = P2.mp.en_k0[e] + P2.mp.ord;
                if (k < En.cap) { En.E[(size_t)k * S.C + c] = E; En.Ev[(size_t)k * S.C + c] = Ev; }
                double *a = En.acc + (size_t)c * 5;
                a[0] += 1.0; a[1] += E; a[2] += E * E; a[3] += Ev; a[4] += Ev * Ev;
            }
            P2.mdone[c] = 1;
        }
    }
    if (polymer) {                              // whole-cycle moves of the exchange cycles, all warps together (staging area reused)
        __syncthreads();
        d_pcom_cycles<POT, KM, TH>(S, P, st, c, maxd, flag, (char *)sm + (((size_t)N + 127) & ~(size_t)127) + NW * 16, my_beads);
    }
    if (my_beads) atomicAdd(&s_bead, my_beads);
    __syncthreads();
    if (use_tma && lane == 0) { d_mbar_inval(bar_u); d_mbar_inval(bar_u + 8); }   // the shared memory is reused by whatever runs next in this CTA
    if (warp == 0) d_bookkeep_sweep_warp(U, c, flag, N, s_bead, P.stats, s_pre);
}

// The swap move stays one proposal per chain and iteration (reshape.jl:123-283): one warp per chain.  Everything that touches
// memory or a transcendental runs across the lanes (weight table of sampleparticles, Gaussians of the two bridges, potentials, commit,
// tail exchange); the two staging recurrences run on four lanes (bridge x dim); every SUM keeps the sequential order of the
// one-thread reference implementation d_reshape_swap / d_swap_weights on lane 0, so results are bit-identical to it and to the oracle.
// Shared memory (dynamic): w[N] | 2 bridges x { x[M+1], y[M+1], v[M+1], link[M+1], vold[M+1] }.
__device__ __forceinline__ void d_swap_iter_body(const DevSys &S, const Sweep2Params &P2, const pimc_stream &st, const int pick, const int c, double *sm)
{
    const SweepParams &P = P2.sp;
    const int N = S.N, M = S.M, dim = S.dim, lane = threadIdx.x & 31;
    const UpdDev &U = P2.upd[pick];
    double *w = sm, *br = sm + N;                              // br: bridge b at br + b * 5 * (M + 1)
    const int R1 = M + 1;
    unsigned char f = 3; unsigned long long beads = 0;
    if (N > 1) {
        const int var = (int)U.var[c];
        const pimc_u4 dt = pimc_draw_rk(st, &P.rk, 0, PIMC_K_TASK, 0, 0);
        const pimc_u4 dm = pimc_draw_rk(st, &P.rk, 0, PIMC_K_TASK, 0, 1);
        const pimc_u4 dsw = pimc_draw_rk(st, &P.rk, 0, PIMC_K_SWAP, 0, 0);
        const int j0 = 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
        const int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
        const int m = (int)U.vmax < mm ? (int)U.vmax : mm;
        const int n1 = (int)pimc_index(dsw.w[0], (uint32_t)N);
        int *nextc = S.next + (size_t)c * N;
        double *rc = S.r + (size_t)c * N * dim * M, *vc = S.Vl + (size_t)c * N * M;
        // ---- sampleparticles (helper.jl:224-267): weight table across the lanes, sums on lane 0 ----
        {
            const int jmw = (j0 + m - 1) % M;                 // mod1(j0 + m, M) - 1
            const bool wrapw = j0 + m > M;
            const int n1next = wrapw ? nextc[n1] : n1;
            const double mt = m * S.tau;
            const double ax = rc[(n1 * dim) * M + j0 - 1], ay = dim > 1 ? rc[(n1 * dim + 1) * M + j0 - 1] : 0.0;
            const double cx = rc[(n1next * dim) * M + jmw], cy = dim > 1 ? rc[(n1next * dim + 1) * M + jmw] : 0.0;
            for (int i = lane; i < N; i += 32) {
                const int inext = wrapw ? nextc[i] : i;
                const double t = d_lnK2(ax, ay, rc[(inext * dim) * M + jmw], dim > 1 ? rc[(inext * dim + 1) * M + jmw] : 0.0, dim, S.lambda, mt, S.L);
                const double y = d_lnK2(rc[(i * dim) * M + j0 - 1], dim > 1 ? rc[(i * dim + 1) * M + j0 - 1] : 0.0, cx, cy, dim, S.lambda, mt, S.L);
                w[i] = pimc_exp(t + y);
            }
            __syncwarp();
            double norm = 0.0;
            if (lane == 0) { norm = w[0]; for (int i = 1; i < N; ++i) norm = norm + w[i]; }
            norm = __shfl_sync(0xffffffffu, norm, 0);
            for (int i = lane; i < N; i += 32) w[i] = w[i] / norm;
            __syncwarp();
        }
        int n2 = 0;
        if (lane == 0) n2 = d_sample_weighted(w, N, pimc_u01_co(dsw.w[2], dsw.w[3]));
        n2 = __shfl_sync(0xffffffffu, n2, 0);
        if (n1 != n2) {
            // ---- ReshapeSwapLinear body (reshape.jl:138-279), independent worldlines (the sweep schedule has no pair action) ----
            const int jm = j0 + m, rows = m + 1, mb = m - 1;   // mb interior beads per bridge
            const int x1 = nextc[n1], x2 = nextc[n2];
            const bool wrap = jm > M;
            const int je = (wrap ? jm - M : jm) - 1;
            const int e1 = wrap ? x2 : n2, e2 = wrap ? x1 : n1; // bridge 1 ends on the cycle of n2 and vice versa
            const double L = S.L, twoL = 2 * S.L, inv2L = 1.0 / twoL, mht = -0.5 * S.tau;
            // Gaussians of both bridges across the lanes: xi * sigma_j staged in place of the interior rows
            for (int idx = lane; idx < 2 * mb; idx += 32) {
                const int b = idx >= mb ? 1 : 0, j = 1 + idx - b * mb;
                double g0, g1;
                pimc_gauss_pair_t(pimc_draw_rk(st, &P.rk, 0, b ? PIMC_K_BRIDGE2 : PIMC_K_BRIDGE, 0, (uint32_t)j), S.logtab, &g0, &g1);
                const double alpha = (double)(mb + 1 - j) / (double)(mb + 2 - j);
                const double sig = sqrt(2 * S.lambda * alpha * S.tau);
                double *bb = br + b * 5 * R1;
                bb[j] = g0 * sig; if (dim > 1) bb[R1 + j] = g1 * sig;
            }
            // cached link actions of the old configuration (w_initial), fetched across the lanes
            for (int idx = lane; idx < 2 * m; idx += 32) {
                const int b = idx >= m ? 1 : 0, jp = idx - b * m, j = j0 + jp;
                const int q = j <= M ? (b ? n2 : n1) : (b ? x2 : x1), sl = (j <= M ? j : j - M) - 1;
                br[b * 5 * R1 + 4 * R1 + jp] = vc[q * M + sl];
            }
            __syncwarp();
            // the two staging recurrences of levy! (helper.jl:118-139), lanes = (bridge, dim)
            if (lane < 2 * dim) {
                const int b = lane / dim, k = lane - b * dim;
                const int ns = b ? n2 : n1, ne = b ? e2 : e1;
                double q = rc[(ns * dim + k) * M + j0 - 1], e = rc[(ne * dim + k) * M + je];
                if (fabs(q - e) > L) e += d_sign(q) * twoL;
                double *arr = br + b * 5 * R1 + k * R1;
                arr[0] = d_teleport_fast(q, L, twoL, inv2L);
                for (int j = 1; j <= mb; ++j) {
                    const double alpha = (double)(mb + 1 - j) / (double)(mb + 2 - j), om = 1 - alpha;
                    q = alpha * q + om * e + arr[j];
                    arr[j] = d_teleport_fast(q, L, twoL, inv2L);
                }
                arr[rows - 1] = d_teleport_fast(e, L, twoL, inv2L);
            }
            __syncwarp();
            for (int idx = lane; idx < 2 * rows; idx += 32) {  // potential at every row of both bridges
                const int b = idx >= rows ? 1 : 0, row = idx - b * rows;
                double *bb = br + b * 5 * R1;
                bb[2 * R1 + row] = d_pot(S.pot, bb[row], dim > 1 ? bb[R1 + row] : 0.0, dim);
            }
            __syncwarp();
            for (int idx = lane; idx < 2 * m; idx += 32) {     // lnV of the new links (propagator.jl:26-28)
                const int b = idx >= m ? 1 : 0, jp = idx - b * m;
                double *bb = br + b * 5 * R1;
                bb[3 * R1 + jp] = mht * (bb[2 * R1 + jp] + bb[2 * R1 + jp + 1]);
            }
            __syncwarp();
            int ret = 0;
            if (lane == 0) {                                   // the sums in the reference order
                const double *b1 = br, *b2 = br + 5 * R1;
                double w_initial = 0.0, w_updated = 0.0, s1 = 0.0, s2 = 0.0;
                for (int jp = 0; jp < m; ++jp) w_initial += b1[4 * R1 + jp] + b2[4 * R1 + jp];
                for (int jp = 0; jp < m; ++jp) { s1 = jp == 0 ? b1[3 * R1] : s1 + b1[3 * R1 + jp]; s2 = jp == 0 ? b2[3 * R1] : s2 + b2[3 * R1 + jp]; }
                w_updated += s1 + s2;
                ret = d_metropolis(pimc_exp(w_updated - w_initial), pimc_u01_co(dm.w[0], dm.w[1])) ? 1 : 0;
            }
            ret = __shfl_sync(0xffffffffu, ret, 0);
            if (ret) {
                const double *b1 = br, *b2 = br + 5 * R1;
                if (lane == 0) { nextc[n1] = x2; nextc[n2] = x1; }   // reshape.jl:254
                for (int jr = 2 + lane; jr <= m + 1; jr += 32) {     // new rows 2..m+1: beyond slice M they land on the NEW next particle
                    const int j = j0 + jr - 1;
                    const int q1 = j <= M ? n1 : x2, q2 = j <= M ? n2 : x1, sl = (j <= M ? j : j - M) - 1;
                    rc[(q1 * dim) * M + sl] = b1[jr - 1]; if (dim > 1) rc[(q1 * dim + 1) * M + sl] = b1[R1 + jr - 1];
                    rc[(q2 * dim) * M + sl] = b2[jr - 1]; if (dim > 1) rc[(q2 * dim + 1) * M + sl] = b2[R1 + jr - 1];
                }
                for (int jp = 1 + lane; jp <= m; jp += 32) {
                    const int j = j0 + jp - 1;
                    const int q1 = j <= M ? n1 : x2, q2 = j <= M ? n2 : x1, sl = (j <= M ? j : j - M) - 1;
                    vc[q1 * M + sl] = b1[3 * R1 + jp - 1];
                    vc[q2 * M + sl] = b2[3 * R1 + jp - 1];
                }
                if (jm < M)                                          // tails jm+1..M change owner (reshape.jl:269-275)
                    for (int sl = jm + lane; sl < M; sl += 32) {
                        for (int k = 0; k < dim; ++k) {
                            const double t = rc[(n1 * dim + k) * M + sl]; rc[(n1 * dim + k) * M + sl] = rc[(n2 * dim + k) * M + sl]; rc[(n2 * dim + k) * M + sl] = t;
                        }
                        const double tv = vc[n1 * M + sl]; vc[n1 * M + sl] = vc[n2 * M + sl]; vc[n2 * M + sl] = tv;
                    }
                if (lane == 0 && !(S.compat & PIMC_COMPAT_SWAP_STALE_LINK) && jm <= M) {   // intended: the link leaving slice j_m changes owner too
                    const double tv = vc[n1 * M + jm - 1]; vc[n1 * M + jm - 1] = vc[n2 * M + jm - 1]; vc[n2 * M + jm - 1] = tv;
                }
            }
            f = ret ? 1 : 0; beads = 2ull * (unsigned long long)(m - 1);
        }
    }
    if (lane != 0) return;
    // faithful-style bookkeeping of a single proposal (apply!, simulation.jl:19-27)
    RingReg R; R.head = U.ring_head[c]; R.len = U.ring_len[c]; R.sum = U.ring_sum[c]; R.tries = U.tries_var[c];
    U.tries[c] += 1;
    if (f != 3) { U.accepted[c] += f; d_ring_push(U, c, R, f); }
    U.ring_head[c] = R.head; U.ring_len[c] = R.len; U.ring_sum[c] = R.sum; U.tries_var[c] = R.tries;
    U.bead_moves[c] += (long long)beads;
    if ((R.tries % U.adj) == 0) d_adjust(U, c, R);
    if (P.stats) { atomicAdd(P.stats + 0, 1ull); atomicAdd(P.stats + 2, beads); }
}
static __global__ void __launch_bounds__(32) k_swap_iter(const __grid_constant__ DevSys S, const DevTables *__restrict__ T, const __grid_constant__ Sweep2Params P2)
{
    extern __shared__ double sm[];
    const SweepParams &P = P2.sp;
    const int c = blockIdx.x;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter);
    pimc_u4 di = pimc_draw_rk(st, &P.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P, di);
    if (P.kind[pick] != PIMC_UPD_RESHAPE_SWAP) return;
    d_swap_iter_body(S, P2, st, pick, c, sm);
}

// One launch per iteration: every CTA (= chain) picks its update (simulation.jl:33-37) and runs that family's sweep.
template <int POT, int KM, bool FUSE>
__global__ void __launch_bounds__(SWEEP_THREADS, (KM <= 4 ? 1024 : 768) / SWEEP_THREADS) k_sweep(const __grid_constant__ DevSys S, const DevTables *__restrict__ T,
                                                                                                   const __grid_constant__ Sweep2Params P2)
{
    const int c = blockIdx.x;
    const SweepParams &P = P2.sp;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter);
    pimc_u4 di = pimc_draw_rk(st, &P.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_pick_update(P, di);
    const int kind = P.kind[pick];
    if (kind == PIMC_UPD_RESHAPE_LINEAR) d_reshape_sweep_body<POT>(S, P2.upd[pick], P, st, di, pick, P2.cap, c);
    else if (kind == PIMC_UPD_SINGLE_COM || kind == PIMC_UPD_POLYMER_COM) d_com_sweep_body<POT, KM, FUSE>(S, P2.upd[pick], P, st, pick, P2, T, c);
    else if (kind == PIMC_UPD_RESHAPE_SWAP && P2.swap_in_sweep && (threadIdx.x >> 5) == 0) {   // the one swap proposal of this chain, on warp 0: no second launch per iteration
        extern __shared__ double sm_swap[];
        d_swap_iter_body(S, P2, st, pick, c, sm_swap);
    }
}

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
// register-resident variant for M <= 32*KM: all of a worldline's loads are issued before any arithmetic, the next bead comes
// from the neighbouring lane by shuffle
template <int POT, int KM>
__device__ __forceinline__ void d_energy_block_reg(const DevSys &S, int c, double *red, double *E, double *Ev)
{
    const int M = S.M, N = S.N, dim = S.dim, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double twoL = 2 * S.L;
    const double *rc = S.r + (size_t)c * N * dim * M;
    const int *nextc = S.next + (size_t)c * N;
    const int dvk = S.pot.dv_kind;                                  // hoisted: the virial term r . dV(r) is x^2 + y^2 for dV = identity
    double link = 0.0, pot = 0.0, vkin = 0.0;
    for (int n = warp; n < N; n += nw) {
        const double *rx = rc + (size_t)(n * dim) * M, *ry = rx + M;
        const int nx = nextc[n];
        double x[KM], y[KM];
#pragma unroll
        for (int k = 0; k < KM; ++k) { const int j = lane + 32 * k; x[k] = j < M ? rx[j] : 0.0; y[k] = (dim > 1 && j < M) ? ry[j] : 0.0; }
        const double *qx = rc + (size_t)(nx * dim) * M;
        const double x0n = qx[0], y0n = dim > 1 ? qx[M] : 0.0;      // first bead of the next particle of the cycle
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            const int j = lane + 32 * k;
            double bx = __shfl_down_sync(0xffffffffu, x[k], 1), by = __shfl_down_sync(0xffffffffu, y[k], 1);
            const double nbx = __shfl_sync(0xffffffffu, x[(k + 1 < KM) ? k + 1 : k], 0), nby = __shfl_sync(0xffffffffu, y[(k + 1 < KM) ? k + 1 : k], 0);
            if (lane == 31) { bx = nbx; by = nby; }
            if (j == M - 1) { bx = x0n; by = y0n; }
            if (j < M) {
                const double ax = x[k], ay = y[k];
                double dx = fabs(ax - bx); { const double alt = twoL - dx; dx = alt < dx ? alt : dx; }
                double d2 = dx * dx;
                if (dim > 1) { double dy = fabs(ay - by); const double alt = twoL - dy; dy = alt < dy ? alt : dy; d2 = d2 + dy * dy; }
                link += d2;
                if (POT != PIMC_POT_ZERO) pot += d_pot_t<POT>(S.pot, ax, ay, dim) + d_pot_t<POT>(S.pot, bx, by, dim);
                if (dvk == PIMC_DV_IDENTITY) { double s = ax * ax; if (dim > 1) s = s + ay * ay; vkin += s; }   // r . dV(r), measurement.jl:105
                else if (dvk != PIMC_DV_ZERO) vkin += d_rdv(S.pot, ax, ay, dim);
            }
        }
    }
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
// One staged worldline (rows sx | sy in shared memory) into the lane-partial Energy sums.  DV: 0 zero, 1 identity, 2 general gradient --
// hoisted out of the bead loop (a per-bead runtime branch was a tenth of the kernel's instructions); the successor bead is read from the
// staged row itself (no register tile, no shuffles), only the link that leaves the last slice needs the first bead of the cycle's next member.
template <int POT, int KM, int DV, bool FAST>
__device__ __forceinline__ void d_energy_row(const DevSys &S, const double *sx, const double x0n, const double y0n, const double twoL,
                                             double &link, double &pot, double &vkin)
{
    // FAST: dim == 2 and M == 32 * KM, both checked by the caller -- no bead predicates, the y row at a constant offset, the successor of every
    // bead but the last one an unconditional read.  Same arithmetic per bead in the same order: the sums are bit-identical to the general form.
    const int M = FAST ? 32 * KM : S.M, dim = FAST ? 2 : S.dim, lane = threadIdx.x & 31;
    const double *sy = sx + M;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
        const int j = lane + 32 * k;
        if (FAST || j < M) {
            const double ax = sx[j], ay = dim > 1 ? sy[j] : 0.0;
            double bx, by;
            if (FAST && k + 1 < KM) { bx = sx[j + 1]; by = sy[j + 1]; }
            else { bx = j + 1 < M ? sx[j + 1] : x0n; by = dim > 1 ? (j + 1 < M ? sy[j + 1] : y0n) : 0.0; }
            // distance() of propagator.jl:6-9, squared: min(2L - |d|, |d|)^2; the sign of d does not survive the square, so |d| is never materialised
            const double ddx = ax - bx, altx = twoL - fabs(ddx), tx = altx < fabs(ddx) ? altx : ddx;
            double d2 = tx * tx;
            if (dim > 1) { const double ddy = ay - by, alty = twoL - fabs(ddy), ty = alty < fabs(ddy) ? alty : ddy; d2 = d2 + ty * ty; }
            link += d2;
            if (POT != PIMC_POT_ZERO) pot += d_pot_t<POT>(S.pot, ax, ay, dim) + d_pot_t<POT>(S.pot, bx, by, dim);
            if (DV == 1) { double q = ax * ax; if (dim > 1) q = q + ay * ay; vkin += q; }   // r . dV(r), measurement.jl:105
            else if (DV == 2) vkin += d_rdv(S.pot, ax, ay, dim);
        }
    }
}
// TMA-fed Energy pass (even M): every warp keeps the NEXT worldline's rows in flight (cp.async.bulk into its own two-stage shared-memory ring,
// completion on an mbarrier) while it reduces the current one straight from shared memory; the permutation entry of the next worldline is
// fetched one step ahead.  Same per-lane summation order as d_energy_block_reg.
template <int POT, int KM>
__device__ __forceinline__ void d_energy_block_tma(const DevSys &S, int c, double *red, double *E, double *Ev, char *dyn)
{
    const int M = S.M, N = S.N, dim = S.dim, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double twoL = 2 * S.L;
    const double *rc = S.r + (size_t)c * N * dim * M;
    asm volatile("" : "+l"(rc));                        // keep the chain's base in a register pair: left alone, ptxas rematerialises the 64-bit product per worldline
    const int dvk = S.pot.dv_kind;
    const bool fast = dim == 2 && M == 32 * KM;
    // Everything the loop needs per worldline is carried as a running value (32-bit offsets inside the chain, stage indices): the first form
    // recomputed the 64-bit addresses of rows, stages and barriers per worldline -- more than half of the kernel's instructions, and the
    // kernel was issue-bound (ncu: issue slots 77 % busy, DRAM 65 %).
    const unsigned wl = (unsigned)(dim * M), wl_bytes = wl * 8u;               // one worldline: dim rows of M doubles
    const unsigned step = (unsigned)nw * wl;
    constexpr int NST = MEAS_STAGES;                                           // ring depth: NST - 1 worldlines in flight while one is reduced
    double *stage0 = (double *)(dyn + nw * NST * 8) + (size_t)warp * NST * wl;
    const uint32_t bar0 = d_smem_u32((unsigned long long *)dyn + warp * NST), st0 = d_smem_u32(stage0);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) d_mbar_init(bar0 + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double link = 0.0, pot = 0.0, vkin = 0.0;
    unsigned off_iss = (unsigned)warp * wl;                                    // offset (doubles) of the next worldline to fetch
    int n_iss = warp, s_iss = 0;                                               // its index and the stage it goes to
#pragma unroll
    for (int i = 0; i < NST - 1; ++i) {                                        // prologue: fill all stages but one
        if (n_iss < N && lane == 0) { d_mbar_expect_tx(bar0 + 8u * s_iss, wl_bytes); d_bulk_g2s(st0 + (uint32_t)s_iss * wl_bytes, rc + off_iss, wl_bytes, bar0 + 8u * s_iss); }
        n_iss += nw; off_iss += step; s_iss += 1;
    }
    const int *pnx = S.next + (size_t)c * N + warp;                            // permutation entry, fetched one worldline ahead
    int nx = warp < N ? *pnx : 0;
    int s_cur = 0; uint32_t par = 0u;                                          // current stage and the phase parity of its barrier
    for (int n = warp; n < N; n += nw) {
        __syncwarp();                                   // every lane is done with the stage read in the previous step: it is refilled now
        if (n_iss < N && lane == 0) { d_mbar_expect_tx(bar0 + 8u * s_iss, wl_bytes); d_bulk_g2s(st0 + (uint32_t)s_iss * wl_bytes, rc + off_iss, wl_bytes, bar0 + 8u * s_iss); }
        n_iss += nw; off_iss += step; s_iss = s_iss + 1 == NST ? 0 : s_iss + 1;
        pnx += nw;
        const int nx_next = n + nw < N ? *pnx : 0;
        double x0n = 0.0, y0n = 0.0;                    // first bead of the next particle of the cycle
        if (nx != n) { const double *qx = rc + (unsigned)nx * wl; x0n = qx[0]; y0n = dim > 1 ? qx[M] : 0.0; }
        d_mbar_wait(bar0 + 8u * s_cur, par);
        const double *sx_cur = stage0 + (unsigned)s_cur * wl;
        if (nx == n) { x0n = sx_cur[0]; y0n = dim > 1 ? sx_cur[M] : 0.0; }   // closed on itself: its own bead 0
        if (POT != PIMC_POT_LATTICE && fast) {          // (the lattice bodies are large: one copy)
            if (dvk == PIMC_DV_IDENTITY) d_energy_row<POT, KM, 1, POT != PIMC_POT_LATTICE>(S, sx_cur, x0n, y0n, twoL, link, pot, vkin);
            else if (dvk == PIMC_DV_ZERO) d_energy_row<POT, KM, 0, POT != PIMC_POT_LATTICE>(S, sx_cur, x0n, y0n, twoL, link, pot, vkin);
            else d_energy_row<POT, KM, 2, false>(S, sx_cur, x0n, y0n, twoL, link, pot, vkin);
        }
        else if (dvk == PIMC_DV_IDENTITY) d_energy_row<POT, KM, 1, false>(S, sx_cur, x0n, y0n, twoL, link, pot, vkin);
        else if (dvk == PIMC_DV_ZERO) d_energy_row<POT, KM, 0, false>(S, sx_cur, x0n, y0n, twoL, link, pot, vkin);
        else d_energy_row<POT, KM, 2, false>(S, sx_cur, x0n, y0n, twoL, link, pot, vkin);
        nx = nx_next;
        s_cur += 1; if (s_cur == NST) { s_cur = 0; par ^= 1u; }
    }
    __syncwarp();
    if (lane == 0) {   // the shared memory is reused by whatever runs next in this CTA
#pragma unroll
        for (int i = 0; i < NST; ++i) d_mbar_inval(bar0 + 8u * i);
    }
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
// one measurement event of chain c (measurement.jl:1-17 cadence handled by the caller): Energy objects, then Density objects
template <int POT, int KM>
__device__ __forceinline__ void d_measure_body(const DevSys &S, const DevTables *__restrict__ T, const MeasParams &P, const int c, const bool en_done,
                                               double *red, char *dyn)
{
    for (int e = 0; e < P.nen && !en_done; ++e) {
        const EnDev &En = T->en[P.en_id[e]];
        const long long k = P.en_k0[e] + P.ord; // the object's own count (measurement.jl:119-120)
        double E, Ev;
        if (KM > 0 && P.tma && e == 0) d_energy_block_tma<POT, (KM > 0 ? KM : 1)>(S, c, red, &E, &Ev, dyn);
        else if (KM > 0) d_energy_block_reg<POT, (KM > 0 ? KM : 1)>(S, c, red, &E, &Ev);
        else d_energy_block_fast<POT>(S, c, red, &E, &Ev);
        if (threadIdx.x == 0) {
            if (k < En.cap) { En.E[(size_t)k * S.C + c] = E; En.Ev[(size_t)k * S.C + c] = Ev; }
            double *a = En.acc + (size_t)c * 5;
            a[0] += 1.0; a[1] += E; a[2] += E * E; a[3] += Ev; a[4] += Ev * Ev;
        }
        __syncthreads();
    }
    for (int d = 0; d < P.nde; ++d) d_density_block(S, c, T->de[P.de_id[d]]);
}
template <int POT, int KM>
__global__ void __launch_bounds__(256, MEAS_MINBLOCKS) k_measure(const __grid_constant__ DevSys S, const DevTables *__restrict__ T, const __grid_constant__ MeasParams P,
                                                    const unsigned char *__restrict__ mdone)
{
    __shared__ double red[96];
    extern __shared__ __align__(16) char dyn_meas[];   // TMA ring of the Energy pass (P.tma), else unused
    const int c = blockIdx.x;
    d_measure_body<POT, KM>(S, T, P, c, mdone && mdone[c], red, dyn_meas);   // mdone: Energy of this chain was evaluated inside the sweep launch (fused)
}
