// pimc_k_sweep.cu -- per-iteration throughput kernels of the SWEEP schedule for independent worldlines (k_sweep, k_swap_iter, k_measure).
#include <cstdlib>
#include "pimc_sweep.cuh"

typedef void (*sweep_fn)(const DevSys, const DevTables *, const Sweep2Params);
typedef void (*meas_fn)(const DevSys, const DevTables *, const MeasParams, const unsigned char *);

template <int POT, bool FUSE> static sweep_fn pick_sweep(int KM) { return KM <= 1 ? k_sweep<POT, 1, FUSE> : KM <= 2 ? k_sweep<POT, 2, FUSE> : KM <= 4 ? k_sweep<POT, 4, FUSE> : k_sweep<POT, 8, FUSE>; }
template <int POT> static sweep_fn pick_sweep(int KM, bool fuse) { return fuse ? pick_sweep<POT, true>(KM) : pick_sweep<POT, false>(KM); }
template <int POT> static meas_fn pick_meas(int KM)
{
    return KM <= 1 ? k_measure<POT, 1> : KM <= 2 ? k_measure<POT, 2> : KM <= 4 ? k_measure<POT, 4> : KM <= 8 ? k_measure<POT, 8> : k_measure<POT, 0>;
}

cudaError_t pimc_launch_sweep(int grid, size_t smem, cudaStream_t st, const DevSys &S, const DevTables *dT, const Sweep2Params &P)
{
    const int KM = (S.M + 31) / 32, pk = S.pot.kind;
    const bool fuse = P.fuse != 0;
    sweep_fn k = pk == PIMC_POT_ZERO ? pick_sweep<PIMC_POT_ZERO>(KM, fuse) : (pk == PIMC_POT_HARMONIC ? pick_sweep<PIMC_POT_HARMONIC>(KM, fuse) : pick_sweep<PIMC_POT_LATTICE>(KM, fuse));
    static sweep_fn configured[32]; static int nconf = 0;
    bool seen = false; for (int i = 0; i < nconf; ++i) seen |= configured[i] == k;
    if (!seen) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (nconf < 32) configured[nconf++] = k;
    }
    k<<<grid, SWEEP_THREADS, smem, st>>>(S, dT, P);
    return cudaGetLastError();
}
cudaError_t pimc_launch_swap_iter(int grid, cudaStream_t st, const DevSys &S, const DevTables *dT, const Sweep2Params &P)
{
    const size_t smem = swap_smem_bytes(S.N, S.M);
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k_swap_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    k_swap_iter<<<grid, 32, smem, st>>>(S, dT, P);
    return cudaGetLastError();
}
cudaError_t pimc_launch_measure(int grid, cudaStream_t st, const DevSys &S, const DevTables *dT, const MeasParams &P, const unsigned char *mdone)
{
    const int KM = (S.M + 31) / 32, pk = S.pot.kind;
    meas_fn k = pk == PIMC_POT_ZERO ? pick_meas<PIMC_POT_ZERO>(KM) : (pk == PIMC_POT_HARMONIC ? pick_meas<PIMC_POT_HARMONIC>(KM) : pick_meas<PIMC_POT_LATTICE>(KM));
    // Energy pass fed by TMA (even M, register-resident variants): 2 mbarriers + 2 stages of dim rows per warp
    MeasParams Q = P;
    Q.tma = (P.nen > 0 && KM <= 8 && (S.M % 2) == 0 && getenv("PIMC_NO_TMA") == nullptr) ? 1 : 0;
    const size_t smem = Q.tma ? meas_smem_bytes(256 / 32, S.dim, S.M) : 0;
    if (smem > 48 * 1024) {
        static meas_fn configured[16]; static int nconf = 0;
        bool seen = false; for (int i = 0; i < nconf; ++i) seen |= configured[i] == k;
        if (!seen) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); if (e != cudaSuccess) return e; if (nconf < 16) configured[nconf++] = k; }
    }
    k<<<grid, 256, smem, st>>>(S, dT, Q, mdone);
    return cudaGetLastError();
}
