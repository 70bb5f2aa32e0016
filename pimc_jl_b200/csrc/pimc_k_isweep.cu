// pimc_k_isweep.cu -- optimistic-parallel SWEEP kernels for interacting worldlines (hard core through the cell list); see pimc_isweep.cuh.
#include "pimc_isweep.cuh"

cudaError_t pimc_launch_isweep(int grid, cudaStream_t st, const DevSys &S, const ISweepParams &P, bool has_rs, bool has_com)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_isweep_reshape, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (has_rs) k_isweep_reshape<<<grid, ISW_THREADS, isw_rs_smem_bytes(S.N, S.M), st>>>(S, P);
    if (has_com) {
        const size_t smem = isw_com_smem_bytes(S.N);
        if ((S.M + 31) / 32 <= 4) k_isweep_com<4><<<grid, ISW_THREADS, smem, st>>>(S, P);
        else k_isweep_com<8><<<grid, ISW_THREADS, smem, st>>>(S, P);
    }
    return cudaGetLastError();
}
