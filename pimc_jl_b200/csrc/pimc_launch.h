// pimc_launch.h -- parameter blocks shared by the kernel translation units and the host side (pimc_b200.cu), and the
// launchers each kernel TU exports.  The library is built from several TUs compiled in parallel (no relocatable device code:
// every device function lives in a header and each TU carries its own copy).
#pragma once
#include "pimc_device.cuh"

#ifndef SWEEP_THREADS
#define SWEEP_THREADS 256   // threads per CTA of k_sweep / k_chain (A/B knob: 128 = twice the CTAs per SM, scripts/build_variant_all.sh)
#endif

struct SweepParams {
    unsigned long long iter;
    const DevSys *Sg;   // device copy of the system descriptor (for out-of-line slow paths)
    int nupd; int upd_id[PIMC_MAXU]; double w[PIMC_MAXU];
    int kind[PIMC_MAXU]; double vmax[PIMC_MAXU];   // copies of the update descriptors' constants (no global load on the prologue path)
    unsigned long long *stats;
    pimc_roundkeys rk;  // Philox round keys of the seed (constant-bank operands)
    int com_stage_off;  // byte offset of the COM half's TMA staging area in dynamic shared memory (0: register path with plain loads)
};

// measurement_Z_sector (measurement.jl:1-17) for every chain at one cadence hit.  ord = 0-based ordinal of this measurement within
// the current run; Energy object e stores it at index en_k0[e] + ord (its OWN count, like the reference's findfirst(ismissing, ...),
// measurement.jl:119-120).
struct MeasParams { int tma; int nen; int en_id[PIMC_MAXE]; long long en_k0[PIMC_MAXE]; int nde; int de_id[PIMC_MAXD]; long long ord; };

// The update descriptors travel by value in the kernel parameters (constant bank): no dependent global loads of T->upd[...]
// on the prologue or the bookkeeping tail of a CTA (measured: +24 % on the centre-of-mass half, 2x at N = 1024).
// fuse: the sweep launch of a measurement iteration also evaluates the Energy functor for the chains whose picked update streamed
// every worldline anyway (centre-of-mass sweep of a chain without exchange cycles); mdone[c] = 1 tells k_measure to skip that chain.
struct Sweep2Params { SweepParams sp; UpdDev upd[PIMC_MAXU]; int cap; int fuse; MeasParams mp; unsigned char *mdone; double *fscr; int swap_in_sweep; };

// TMA ring of the Energy pass (d_energy_block_tma): per warp MEAS_STAGES mbarriers and MEAS_STAGES stages of one worldline (dim rows of M doubles)
#ifndef MEAS_STAGES
#define MEAS_STAGES 2       // (3 measured: no change, 100.8 vs 101 us per event on C2 -- the pass is not short of bytes in flight)
#endif
#ifndef MEAS_MINBLOCKS
#define MEAS_MINBLOCKS 4     // resident CTAs per SM k_measure is compiled for (64 registers; 3 = 80 registers, no spills: A/B knob, measured no change)
#endif
#ifndef COM_PREFETCH_NEXT
#define COM_PREFETCH_NEXT 0   // centre-of-mass sweep: permutation entry of the next proposal fetched one step ahead (A/B knob; measured 4 % SLOWER on C2: the entry is an L1 hit, the extra live register is not free at the 64-register cap)
#endif
__host__ __device__ inline size_t meas_smem_bytes(int nwarps, int dim, int M) { return (size_t)nwarps * MEAS_STAGES * 8 + (size_t)nwarps * MEAS_STAGES * dim * M * sizeof(double); }
__host__ __device__ inline size_t pcom_smem_bytes(int N) { return (size_t)53 * N + 64; }
__host__ __device__ inline size_t swap_smem_bytes(int N, int M) { return ((size_t)N + 10 * (size_t)(M + 1)) * sizeof(double) + 16; }
#define FA_ARR 6
__host__ __device__ inline size_t faithful_scratch_doubles(int N, int M) { return (size_t)((N + 1) & ~1) + 2 * FA_ARR * (size_t)(M + 1); }

// ---- optimistic-parallel sweep of interacting worldlines (pimc_isweep.cuh) ----
#define ISW_THREADS 128
#define ISW_RARR 5
struct ISweepParams {
    SweepParams sp;
    UpdDev upd[PIMC_MAXU];
    int *nw_head;               // [C][ncell][M] lists over the PROPOSED positions (prop), all -1 between launches
    int *nw_next;               // [C][N][M]
    unsigned long long *prof;   // [16] phase cycle counters / replay counts (PIMC_PROF), or null
};

// shared-memory layout of both kernels: flag[N] | stat[N] | mlen[N] | prev[N] | lead[N] | per-warp rows
__host__ __device__ inline size_t isw_fixed_bytes(int N) { return (((size_t)N + 15) & ~(size_t)15) + (size_t)N * (4 + 2 + 2 + 4) + 64; }
__host__ __device__ inline size_t isw_rs_smem_bytes(int N, int M) { return isw_fixed_bytes(N) + (size_t)(ISW_THREADS / 32) * ISW_RARR * (M + 2) * sizeof(double); }

__host__ __device__ inline size_t isw_com_smem_bytes(int N) { return isw_fixed_bytes(N) + 40 * sizeof(double); }

#ifndef PIMC_CELLS_THREADS
#define PIMC_CELLS_THREADS 128
#endif

// ---- launchers (host functions defined next to their kernels) ----
cudaError_t pimc_launch_run(bool cells, int grid, int threads, size_t smem, cudaStream_t st, const DevSys &S, const DevTables *dT, const RunParams &P);
cudaError_t pimc_launch_sweep(int grid, size_t smem, cudaStream_t st, const DevSys &S, const DevTables *dT, const Sweep2Params &P);
cudaError_t pimc_launch_swap_iter(int grid, cudaStream_t st, const DevSys &S, const DevTables *dT, const Sweep2Params &P);
cudaError_t pimc_launch_measure(int grid, cudaStream_t st, const DevSys &S, const DevTables *dT, const MeasParams &P, const unsigned char *mdone);
size_t pimc_paircorr_smem(const DevSys &S, const PcDev &G, int *TS, int *smem_hist);
// chain-major persistent kernel (pimc_chain.cuh)
struct ChainParams {
    Sweep2Params sw;            // update descriptors, weights, round keys (sw.sp.iter = first iteration of the call)
    MeasParams mp;              // estimator objects; mp.ord = index of the first measurement event of this call
    long long n;                // iterations per chain
    long long Nctr0; int Ncycle; int measure;   // cadence counter at the start of the call (measurement.jl:1-17)
    int *queue;                 // next chain to take
};
cudaError_t pimc_launch_chain(size_t smem, cudaStream_t st, const DevSys &S, const DevTables *dT, const ChainParams &Q, int *grid_out);
cudaError_t pimc_launch_paircorr(int grid, cudaStream_t st, const DevSys &S, const PcDev &G);
cudaError_t pimc_launch_winding(int grid, cudaStream_t st, const DevSys &S, const WiDev &W, long long k);
size_t pimc_structure_smem(const DevSys &S, int *TS);
cudaError_t pimc_launch_structure(int grid, cudaStream_t st, const DevSys &S, const SkDev &K);
cudaError_t pimc_launch_isweep(int grid, cudaStream_t st, const DevSys &S, const ISweepParams &P, bool has_rs, bool has_com);
cudaError_t pimc_launch_iswap(int grid, cudaStream_t st, const DevSys &S, const DevTables *dT, const Sweep2Params &P);
