// pimc_chain.cuh -- chain-major execution of the SWEEP schedule for independent worldlines: ONE launch per run!, persistent CTAs, and every CTA
// takes a chain through ALL n iterations of the call before it fetches the next chain from a queue (chains never interact, so no order between
// them has to be kept).  Same device bodies, same draws and therefore the same bits as the per-iteration kernels of pimc_sweep.cuh; what
// changes is the memory traffic: a chain's state (N * M * 24 B, 192 KiB for C2) is read from HBM once per CALL and then lives in L2 (the
// resident CTAs' chains together fit the 126 MB L2), the estimator pass of a measurement iteration re-reads it from L2, and the ~n kernel
// boundaries (tail drain + ramp-up of every launch) disappear.
#pragma once
#include "pimc_sweep.cuh"


template <int POT, int KM>
__global__ void __launch_bounds__(SWEEP_THREADS, (KM <= 4 ? 1024 : 768) / SWEEP_THREADS) k_chain(const __grid_constant__ DevSys S, const DevTables *__restrict__ T,
                                                                                                   const __grid_constant__ ChainParams Q)
{
    extern __shared__ double sm[];
    __shared__ int s_c;
    __shared__ double red[96];
    const SweepParams &P = Q.sw.sp;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_c = atomicAdd(Q.queue, 1);
        __syncthreads();
        const int c = s_c;
        if (c >= S.C) return;
        long long ord = Q.mp.ord;
        for (long long it = 0; it < Q.n; ++it) {
            const pimc_stream st = pimc_stream_make(S.seed, S.chain_offset + c, P.iter + (unsigned long long)it);
            const pimc_u4 di = pimc_draw_rk(st, &P.rk, PIMC_SLOT_CHAIN, PIMC_K_ITER, 0, 0);
            const int pick = d_pick_update(P, di);
            const int kind = P.kind[pick];
            if (kind == PIMC_UPD_RESHAPE_LINEAR) d_reshape_sweep_body<POT>(S, Q.sw.upd[pick], P, st, di, pick, Q.sw.cap, c);
            else if (kind == PIMC_UPD_SINGLE_COM || kind == PIMC_UPD_POLYMER_COM) d_com_sweep_body<POT, KM, false>(S, Q.sw.upd[pick], P, st, pick, Q.sw, T, c);
            else if (warp == 0) d_swap_iter_body(S, Q.sw, st, pick, c, sm);
            // rows written with ordinary stores are read by bulk-async copies (TMA) in later iterations: order the two proxies, then the CTA
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (Q.measure && (Q.Nctr0 + it + 1) % Q.Ncycle == 0) {
                MeasParams M1 = Q.mp; M1.ord = ord; ++ord;
                d_measure_body<POT, KM>(S, T, M1, c, false, red, (char *)sm);
                __syncthreads();
            }
        }
    }
}
