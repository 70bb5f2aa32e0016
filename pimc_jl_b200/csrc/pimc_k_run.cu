// pimc_k_run.cu -- the persistent run! kernels (reference schedule; sequential sweep of interacting worldlines) of libpimc_b200.so.
#include "pimc_moves.cuh"
#include "pimc_faithful.cuh"
#include "pimc_launch.h"

// ---- run! (simulation.jl:29-42): one persistent CTA per chain, all n iterations inside the kernel ----
// dynamic shared memory: 96 doubles (reductions) + scratch of the warp-cooperative proposal (pimc_faithful.cuh; in HBM when
// P.fscr is set) + N bytes (per-task outcome) + control words
// CELLS = false: systems without hard core / pair action / cell list; the compiler is told so and drops every neighbour query
// (they are out-of-line calls that would otherwise push the kernel to the register cap).
template <bool CELLS>
__device__ __forceinline__ void d_run_body(const DevSys &S, const DevTables *__restrict__ T, const RunParams &P)
{
    if (!CELLS) { __builtin_assume(S.need_cells == 0); __builtin_assume(S.interactions == 0); __builtin_assume(!(S.a > 0.0)); }
    extern __shared__ double smem[];
    double *red = smem;
    const size_t fs_doubles = P.fimpl == 0 ? faithful_scratch_doubles(S.N, S.M) : 0;
    double *fscr = P.fscr ? P.fscr + (size_t)blockIdx.x * fs_doubles : smem + 96;
    unsigned char *flag = (unsigned char *)(smem + 96 + (P.fscr ? 0 : fs_doubles));
    __shared__ unsigned long long s_bead;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int M = S.M, N = S.N;
    unsigned long long tot_bead = 0, tot_prop = 0;
    unsigned long long prof_c[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }; // cycles per update kind [0..3], proposals [4..7], bookkeeping, estimators

    for (int c = blockIdx.x; c < S.C; c += gridDim.x) {
        for (long long it = 0; it < P.n; ++it) {
            pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter0 + (unsigned long long)it);
            pimc_u4 di = f_draw(st, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
            int pick = d_sample_weighted(P.w, P.nupd, pimc_u01_co(di.w[0], di.w[1]));
            const UpdDev &U = T->upd[P.upd_id[pick]];
            const double var = U.var[c];
            const int sweep = (P.sched == PIMC_SCHED_SWEEP) && U.kind != PIMC_UPD_RESHAPE_SWAP;
            if (tid == 0) { s_bead = 0; }
            for (int i = tid; i < N; i += blockDim.x) flag[i] = 2; // 2 = slot not proposed
            __syncthreads();
            const long long t_move0 = P.prof ? clock64() : 0;

            if (CELLS && sweep && P.fimpl == 0 && U.kind == PIMC_UPD_RESHAPE_LINEAR) {
                // sweep of INTERACTING worldlines: every worldline proposes once per iteration in one common time window, strictly in
                // order (proposal n sees the committed results of the proposals before it: the oracle's ORA_SCHED_SWEEP_SEQ); each
                // proposal is the warp-cooperative body.  Amortises the iteration overhead and the estimators over N proposals.
                if (warp == 0) {
                    const int j0w = 1 + (int)pimc_index(di.w[2], (uint32_t)M);
                    unsigned long long bm = 0;
                    for (int slot = 0; slot < N; ++slot) {
                        pimc_u4 dt = f_draw(st, (uint32_t)slot, PIMC_K_TASK, 0, 0);
                        pimc_u4 dm = f_draw(st, (uint32_t)slot, PIMC_K_TASK, 0, 1);
                        int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)((int)var - 1));
                        int m = (int)U.vmax < mm ? (int)U.vmax : mm;
                        GSrc g; g.xi = nullptr; g.st = st; g.slot = (uint32_t)slot; g.kind = PIMC_K_BRIDGE; g.tab = S.logtab;
                        int r = d_reshape_linear_w(S, c, slot, j0w, m, g, pimc_u01_co(dm.w[0], dm.w[1]), fscr + ((N + 1) & ~1));
                        if (lane == 0) flag[slot] = r == 1 ? 1 : 0;
                        bm += (unsigned long long)(m - 1);
                        __syncwarp();
                    }
                    if (lane == 0) s_bead = bm;
                }
            } else if (CELLS && sweep && P.fimpl == 0) { // centre-of-mass sweep of interacting worldlines: the whole CTA on one proposal at a time, in order
                const bool polymer = U.kind == PIMC_UPD_POLYMER_COM;
                const int *nextc = S.next + (size_t)c * N;
                unsigned long long bm = 0;
                for (int slot = 0; slot < N; ++slot) {
                    bool run_it;
                    if (!polymer) run_it = nextc[slot] == slot;
                    else { run_it = true; int p = nextc[slot], cnt = 0; while (p != slot && cnt <= N) { if (p < slot) run_it = false; p = nextc[p]; cnt++; } }
                    if (!run_it) continue;
                    pimc_u4 dm = f_draw(st, (uint32_t)slot, PIMC_K_TASK, 0, 1);
                    DSrc ds; ds.d = nullptr; ds.st = st; ds.slot = (uint32_t)slot;
                    int npol = 1;
                    int r = d_com_cta(S, c, slot, var, ds, pimc_u01_co(dm.w[0], dm.w[1]), red, &npol, fscr, fs_doubles);
                    if (tid == 0) flag[slot] = r == 1 ? 1 : 0;
                    bm += (unsigned long long)M * npol;
                    __syncthreads();
                }
                if (tid == 0) s_bead = bm;
            } else if (U.kind == PIMC_UPD_RESHAPE_LINEAR && !sweep && P.fimpl == 0) {
                if (warp == 0) { // one proposal, the whole warp on it (pimc_faithful.cuh)
                    pimc_u4 dt = f_draw(st, 0, PIMC_K_TASK, 0, 0);
                    pimc_u4 dm = f_draw(st, 0, PIMC_K_TASK, 0, 1);
                    int n = (int)pimc_index(dt.w[0], (uint32_t)N);
                    int j0 = 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
                    int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)((int)var - 1));
                    int m = (int)U.vmax < mm ? (int)U.vmax : mm;
                    GSrc g; g.xi = nullptr; g.st = st; g.slot = 0; g.kind = PIMC_K_BRIDGE; g.tab = S.logtab;
                    int r = d_reshape_linear_w(S, c, n, j0, m, g, pimc_u01_co(dm.w[0], dm.w[1]), fscr + ((N + 1) & ~1));
                    if (lane == 0) { flag[0] = r == 1 ? 1 : 0; s_bead = (unsigned long long)(m - 1); }
                }
            } else if (U.kind == PIMC_UPD_RESHAPE_SWAP && P.fimpl == 0) {
                if (warp == 0 && N > 1) {
                    pimc_u4 dt = f_draw(st, 0, PIMC_K_TASK, 0, 0);
                    pimc_u4 dm = f_draw(st, 0, PIMC_K_TASK, 0, 1);
                    pimc_u4 dsw = f_draw(st, 0, PIMC_K_SWAP, 0, 0);
                    int j0 = 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
                    int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)((int)var - 1));
                    int m = (int)U.vmax < mm ? (int)U.vmax : mm;
                    int n1 = (int)pimc_index(dsw.w[0], (uint32_t)N);
                    int n2 = d_sample_partner_w(S, c, n1, j0, m, pimc_u01_co(dsw.w[2], dsw.w[3]), fscr);
                    if (n1 == n2) { if (lane == 0) flag[0] = 3; } // early return without queue!(counter_var) (reshape.jl:134-136)
                    else {
                        GSrc g1, g2; g1.xi = nullptr; g1.st = st; g1.slot = 0; g1.kind = PIMC_K_BRIDGE; g1.tab = S.logtab; g2 = g1; g2.kind = PIMC_K_BRIDGE2;
                        int r = d_reshape_swap_w(S, c, n1, n2, j0, m, g1, g2, pimc_u01_co(dm.w[0], dm.w[1]), fscr + ((N + 1) & ~1));
                        if (lane == 0) { flag[0] = r == 1 ? 1 : 0; s_bead = 2ull * (unsigned long long)(m - 1); }
                    }
                } else if (tid == 0 && N <= 1) flag[0] = 3;
            } else if (U.kind == PIMC_UPD_RESHAPE_LINEAR) {
                const int ntask = sweep ? N : 1;
                const int j0w = 1 + (int)pimc_index(di.w[2], (uint32_t)M);
                unsigned long long bm = 0;
                for (int slot = tid; slot < ntask; slot += blockDim.x) {
                    pimc_u4 dt = pimc_draw(st, (uint32_t)slot, PIMC_K_TASK, 0, 0);
                    pimc_u4 dm = pimc_draw(st, (uint32_t)slot, PIMC_K_TASK, 0, 1);
                    int n = sweep ? slot : (int)pimc_index(dt.w[0], (uint32_t)N);
                    int j0 = sweep ? j0w : 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
                    int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)((int)var - 1));
                    int m = (int)U.vmax < mm ? (int)U.vmax : mm;
                    GSrc g; g.xi = nullptr; g.st = st; g.slot = (uint32_t)slot; g.kind = PIMC_K_BRIDGE; g.tab = S.logtab;
                    int r = d_reshape_linear(S, c, n, j0, m, g, pimc_u01_co(dm.w[0], dm.w[1]), 1, slot, nullptr, nullptr);
                    flag[slot] = r == 1 ? 1 : 0;
                    bm += (unsigned long long)(m - 1);
                }
                if (bm) atomicAdd(&s_bead, bm);
            } else if (U.kind == PIMC_UPD_RESHAPE_SWAP) {
                if (tid == 0 && N > 1) {
                    pimc_u4 dt = pimc_draw(st, 0, PIMC_K_TASK, 0, 0);
                    pimc_u4 dm = pimc_draw(st, 0, PIMC_K_TASK, 0, 1);
                    pimc_u4 dsw = pimc_draw(st, 0, PIMC_K_SWAP, 0, 0);
                    int j0 = 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
                    int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)((int)var - 1));
                    int m = (int)U.vmax < mm ? (int)U.vmax : mm;
                    int n1 = (int)pimc_index(dsw.w[0], (uint32_t)N);
                    double *w = S.wtab + (size_t)c * N;
                    d_swap_weights(S, c, n1, j0, m, w);
                    double norm = w[0]; for (int i = 1; i < N; ++i) norm = norm + w[i];
                    for (int i = 0; i < N; ++i) w[i] = w[i] / norm;
                    int n2 = d_sample_weighted(w, N, pimc_u01_co(dsw.w[2], dsw.w[3]));
                    if (n1 == n2) flag[0] = 3; // early return without queue!(counter_var) (reshape.jl:134-136)
                    else {
                        GSrc g1, g2; g1.xi = nullptr; g1.st = st; g1.slot = 0; g1.kind = PIMC_K_BRIDGE; g1.tab = S.logtab; g2 = g1; g2.kind = PIMC_K_BRIDGE2;
                        int r = d_reshape_swap(S, c, n1, n2, j0, m, g1, g2, pimc_u01_co(dm.w[0], dm.w[1]), 1, nullptr, nullptr);
                        flag[0] = r == 1 ? 1 : 0;
                        s_bead = 2ull * (unsigned long long)(m - 1);
                    }
                } else if (tid == 0) flag[0] = 3;
            } else { // centre-of-mass moves: one warp per proposal
                const bool polymer = U.kind == PIMC_UPD_POLYMER_COM;
                const int *nextc = S.next + (size_t)c * N;
                if (sweep) {
                    for (int slot = warp; slot < N; slot += nwarp) {
                        bool run_it;
                        if (!polymer) run_it = nextc[slot] == slot;
                        else { run_it = true; int p = nextc[slot], cnt = 0; while (p != slot && cnt <= N) { if (p < slot) run_it = false; p = nextc[p]; cnt++; } }
                        if (!run_it) continue;
                        pimc_u4 dm = pimc_draw(st, (uint32_t)slot, PIMC_K_TASK, 0, 1);
                        DSrc ds; ds.d = nullptr; ds.st = st; ds.slot = (uint32_t)slot;
                        int npol = 1;
                        int r = d_com_warp(S, c, slot, var, ds, pimc_u01_co(dm.w[0], dm.w[1]), 1, nullptr, nullptr, &npol);
                        if (lane == 0) { flag[slot] = r == 1 ? 1 : 0; atomicAdd(&s_bead, (unsigned long long)M * npol); }
                    }
                } else if (P.fimpl == 0) { // one proposal, the whole CTA on it (hard-core tests of all beads: pimc_faithful.cuh)
                    pimc_u4 dt = f_draw(st, 0, PIMC_K_TASK, 0, 0);
                    pimc_u4 dm = f_draw(st, 0, PIMC_K_TASK, 0, 1);
                    int n = -1;
                    if (polymer) n = (int)pimc_index(dt.w[0], (uint32_t)N);
                    else { // uniform among particles with next == self (com.jl:144-164)
                        int cnt = 0; for (int i = 0; i < N; ++i) cnt += nextc[i] == i;
                        if (cnt > 0) { int k = (int)pimc_index(dt.w[0], (uint32_t)cnt); for (int i = 0; i < N; ++i) if (nextc[i] == i && k-- == 0) { n = i; break; } }
                    }
                    if (n < 0) { if (tid == 0) flag[0] = 3; }
                    else {
                        DSrc ds; ds.d = nullptr; ds.st = st; ds.slot = 0;
                        int npol = 1;
                        int r = d_com_cta(S, c, n, var, ds, pimc_u01_co(dm.w[0], dm.w[1]), red, &npol, fscr, fs_doubles);
                        if (tid == 0) { flag[0] = r == 1 ? 1 : 0; s_bead = (unsigned long long)M * npol; }
                    }
                } else if (warp == 0) {
                    pimc_u4 dt = pimc_draw(st, 0, PIMC_K_TASK, 0, 0);
                    pimc_u4 dm = pimc_draw(st, 0, PIMC_K_TASK, 0, 1);
                    int n = -1;
                    if (polymer) n = (int)pimc_index(dt.w[0], (uint32_t)N);
                    else { // uniform among particles with next == self (com.jl:144-164)
                        int cnt = 0; for (int i = 0; i < N; ++i) cnt += nextc[i] == i;
                        if (cnt > 0) { int k = (int)pimc_index(dt.w[0], (uint32_t)cnt); for (int i = 0; i < N; ++i) if (nextc[i] == i && k-- == 0) { n = i; break; } }
                    }
                    if (n < 0) { if (lane == 0) flag[0] = 3; }
                    else {
                        DSrc ds; ds.d = nullptr; ds.st = st; ds.slot = 0;
                        int npol = 1;
                        int r = d_com_warp(S, c, n, var, ds, pimc_u01_co(dm.w[0], dm.w[1]), 1, nullptr, nullptr, &npol);
                        if (lane == 0) { flag[0] = r == 1 ? 1 : 0; s_bead = (unsigned long long)M * npol; }
                    }
                }
            }
            __syncthreads();
            const long long t_move1 = P.prof ? clock64() : 0;
            if (P.prof && tid == 0) { prof_c[U.kind & 3] += (unsigned long long)(t_move1 - t_move0); prof_c[4 + (U.kind & 3)] += 1; }

            // apply! bookkeeping (simulation.jl:19-27), replayed in slot order by one thread
            if (tid == 0) {
                RingReg R; R.head = U.ring_head[c]; R.len = U.ring_len[c]; R.sum = U.ring_sum[c]; R.tries = U.tries_var[c];
                const long long tries0 = R.tries; long long tr = U.tries[c], ac = U.accepted[c]; int cnt = 0;
                const int ntask = sweep ? N : 1;
                for (int slot = 0; slot < ntask; ++slot) {
                    int f = flag[slot];
                    if (f == 2) continue;
                    cnt += 1; tr += 1;
                    if (f == 3) continue;
                    ac += f; d_ring_push(U, c, R, f);
                }
                bool adj;
                if (sweep) adj = cnt > 0 && (R.tries / U.adj) != (tries0 / U.adj);
                else adj = (R.tries % U.adj) == 0;
                U.ring_head[c] = R.head; U.ring_len[c] = R.len; U.ring_sum[c] = R.sum; U.tries_var[c] = R.tries;
                U.tries[c] = tr; U.accepted[c] = ac; U.bead_moves[c] += (long long)s_bead;
                if (adj) d_adjust(U, c, R);
                tot_bead += s_bead; tot_prop += (unsigned long long)cnt;
            }
            const long long t_book = P.prof ? clock64() : 0;
            if (P.prof && tid == 0) prof_c[8] += (unsigned long long)(t_book - t_move1);
            // measurement_Z_sector (measurement.jl:1-17): deterministic cadence, identical on every chain
            if (P.nen + P.nde > 0) {
                long long ctrv = P.Nctr0 + it + 1;
                if (ctrv % P.Ncycle == 0) {
                    const long long ord = ctrv / P.Ncycle - 1; // 0-based ordinal of this measurement within the run (Nctr0 < Ncycle)
                    __syncthreads();
                    for (int e = 0; e < P.nen; ++e) {
                        const EnDev &En = T->en[P.en_id[e]];
                        const long long k = P.en_k0[e] + ord;   // the object's own count (measurement.jl:119-120)
                        double E, Ev;
                        d_energy_block(S, c, red, &E, &Ev, nullptr);
                        if (tid == 0) {
                            if (k < En.cap) { En.E[(size_t)k * S.C + c] = E; En.Ev[(size_t)k * S.C + c] = Ev; }
                            double *a = En.acc + (size_t)c * 5;
                            a[0] += 1.0; a[1] += E; a[2] += E * E; a[3] += Ev; a[4] += Ev * Ev;
                        }
                        __syncthreads();
                    }
                    for (int d = 0; d < P.nde; ++d) d_density_block(S, c, T->de[P.de_id[d]]);
                }
            }
            __syncthreads();
            if (P.prof && tid == 0) prof_c[9] += (unsigned long long)(clock64() - t_book);
        }
    }
    if (tid == 0 && P.prof) for (int i = 0; i < 10; ++i) atomicAdd(P.prof + i, prof_c[i]);
    if (tid == 0 && P.stats) { atomicAdd(P.stats + 0, tot_prop); atomicAdd(P.stats + 2, tot_bead); }
}
__global__ void __launch_bounds__(256, 3) k_run(const __grid_constant__ DevSys S, const DevTables *__restrict__ T, const __grid_constant__ RunParams P) { d_run_body<false>(S, T, P); }
// interacting systems: CTAs of at most 64 threads, eight per SM (every chain of the 1024-chain configurations resident at once)
__global__ void __launch_bounds__(PIMC_CELLS_THREADS, 8) k_run_cells(const __grid_constant__ DevSys S, const DevTables *__restrict__ T, const __grid_constant__ RunParams P) { d_run_body<true>(S, T, P); }


// The swap move of the per-iteration path for interacting systems (optimistic sweeps, pimc_isweep.cuh): one proposal per chain and
// iteration (reshape.jl:123-283) by the warp-cooperative body -- sampleparticles table, hard-core bridges, pair sums through the cell
// list, commit with cell-list surgery.  Shared memory: the proposal scratch of pimc_faithful.cuh.
__global__ void __launch_bounds__(32) k_iswap(const __grid_constant__ DevSys S, const DevTables *__restrict__ T, const __grid_constant__ Sweep2Params P2)
{
    extern __shared__ double smem[];
    const SweepParams &P = P2.sp;
    const int c = blockIdx.x, N = S.N, M = S.M, lane = threadIdx.x;
    pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter);
    const pimc_u4 di = f_draw(st, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
    const int pick = d_sample_weighted(P.w, P.nupd, pimc_u01_co(di.w[0], di.w[1]));
    if (P.kind[pick] != PIMC_UPD_RESHAPE_SWAP) return;
    const UpdDev &U = P2.upd[pick];
    double *fscr = P2.fscr ? P2.fscr + (size_t)c * faithful_scratch_doubles(N, M) : smem;
    int f = 3; unsigned long long beads = 0;
    if (N > 1) {
        const int var = (int)U.var[c];
        const pimc_u4 dt = f_draw(st, 0, PIMC_K_TASK, 0, 0), dm = f_draw(st, 0, PIMC_K_TASK, 0, 1), dsw = f_draw(st, 0, PIMC_K_SWAP, 0, 0);
        const int j0 = 1 + (int)pimc_index(dt.w[1], (uint32_t)M);
        const int mm = 2 + (int)pimc_index(dt.w[2], (uint32_t)(var - 1));
        const int m = (int)U.vmax < mm ? (int)U.vmax : mm;
        const int n1 = (int)pimc_index(dsw.w[0], (uint32_t)N);
        const int n2 = d_sample_partner_w(S, c, n1, j0, m, pimc_u01_co(dsw.w[2], dsw.w[3]), fscr);
        if (n1 != n2) {   // n1 == n2: early return without queue!(counter_var) (reshape.jl:134-136)
            GSrc g1, g2; g1.xi = nullptr; g1.st = st; g1.slot = 0; g1.kind = PIMC_K_BRIDGE; g1.tab = S.logtab; g2 = g1; g2.kind = PIMC_K_BRIDGE2;
            const int r = d_reshape_swap_w(S, c, n1, n2, j0, m, g1, g2, pimc_u01_co(dm.w[0], dm.w[1]), fscr + ((N + 1) & ~1));
            f = r == 1 ? 1 : 0; beads = 2ull * (unsigned long long)(m - 1);
        }
    }
    if (lane != 0) return;
    RingReg R; R.head = U.ring_head[c]; R.len = U.ring_len[c]; R.sum = U.ring_sum[c]; R.tries = U.tries_var[c];   // apply! (simulation.jl:19-27)
    U.tries[c] += 1;
    if (f != 3) { U.accepted[c] += f; d_ring_push(U, c, R, f); }
    U.ring_head[c] = R.head; U.ring_len[c] = R.len; U.ring_sum[c] = R.sum; U.tries_var[c] = R.tries;
    U.bead_moves[c] += (long long)beads;
    if ((R.tries % U.adj) == 0) d_adjust(U, c, R);
    if (P.stats) { atomicAdd(P.stats + 0, 1ull); atomicAdd(P.stats + 2, beads); }
}
cudaError_t pimc_launch_iswap(int grid, cudaStream_t st, const DevSys &S, const DevTables *dT, const Sweep2Params &P)
{
    size_t smem = P.fscr ? 16 : faithful_scratch_doubles(S.N, S.M) * sizeof(double);
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k_iswap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    k_iswap<<<grid, 32, smem, st>>>(S, dT, P);
    return cudaGetLastError();
}

cudaError_t pimc_launch_run(bool cells, int grid, int threads, size_t smem, cudaStream_t st, const DevSys &S, const DevTables *dT, const RunParams &P)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cells ? k_run_cells : k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (cells) k_run_cells<<<grid, threads, smem, st>>>(S, dT, P); else k_run<<<grid, threads, smem, st>>>(S, dT, P);
    return cudaGetLastError();
}
