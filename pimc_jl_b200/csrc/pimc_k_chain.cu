// pimc_k_chain.cu -- chain-major persistent kernel of the SWEEP schedule (independent worldlines); see pimc_chain.cuh.
#include <cstdlib>
#include "pimc_chain.cuh"

typedef void (*chain_fn)(const DevSys, const DevTables *, const ChainParams);
template <int POT> static chain_fn pick_chain(int KM) { return KM <= 1 ? k_chain<POT, 1> : KM <= 2 ? k_chain<POT, 2> : KM <= 4 ? k_chain<POT, 4> : k_chain<POT, 8>; }

// grid: every CTA slot of the device once (occupancy of this instantiation at this shared-memory size) -- or fewer when there are fewer chains
cudaError_t pimc_launch_chain(size_t smem, cudaStream_t st, const DevSys &S, const DevTables *dT, const ChainParams &Q, int *grid_out)
{
    const int KM = (S.M + 31) / 32, pk = S.pot.kind;
    chain_fn k = pk == PIMC_POT_ZERO ? pick_chain<PIMC_POT_ZERO>(KM) : (pk == PIMC_POT_HARMONIC ? pick_chain<PIMC_POT_HARMONIC>(KM) : pick_chain<PIMC_POT_LATTICE>(KM));
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0, dev = 0, sms = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SWEEP_THREADS, smem)) != cudaSuccess) return e;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (const char *env = getenv("PIMC_CHAIN_CTAS_PER_SM")) { const int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }
    int grid = per_sm * sms; if (grid > S.C) grid = S.C; if (grid < 1) grid = 1;
    if (grid_out) *grid_out = grid;
    k<<<grid, SWEEP_THREADS, smem, st>>>(S, dT, Q);
    return cudaGetLastError();
}
