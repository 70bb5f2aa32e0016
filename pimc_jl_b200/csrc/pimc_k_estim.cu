// pimc_k_estim.cu -- the estimators the reference lists as TODO (src/measurement.jl:125-127): radial distribution g(r) and winding number
// (superfluid fraction).  Definitions: include/pimc_b200.h (the CPU checker restates them); parity bar: the histogram integer for
// integer, the winding sums to 1e-9 (and their nearest integers equal).
#include "pimc_launch.h"

__device__ __forceinline__ double e_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- g(r): equal-time pair distances of one chain per CTA.  Tiles of TS consecutive slices are staged in shared memory (every particle's
// TS-slice segment is one or two full 32-byte sectors of its row), all pairs i < j of a tile are binned into a shared-memory histogram
// (32-bit counts: at most N^2 M / 2 per launch), which is added to the 64-bit global counters once per chain.  HBM traffic: the positions
// once (16 B per bead); the pair arithmetic (N/2 distances per bead) is the bound -- fp64 pipe.
__global__ void __launch_bounds__(256) k_paircorr(const __grid_constant__ DevSys S, const __grid_constant__ PcDev G, int TS, int smem_hist)
{
    extern __shared__ double sm[];
    const int c = blockIdx.x, tid = threadIdx.x, N = S.N, M = S.M, dim = S.dim;
    const int TP = TS + 1;                                   // padded row: consecutive particles hit different banks
    double *xs = sm, *ys = sm + (size_t)N * TP;
    unsigned *hist = (unsigned *)(sm + (size_t)dim * N * TP);
    const int nb = (int)G.nbins;
    if (smem_hist) for (int i = tid; i < nb; i += blockDim.x) hist[i] = 0u;
    const double *rc = S.r + (size_t)c * N * dim * M;
    const double twoL = 2 * S.L, bin = G.bin;
    const double inv = 1.0 / bin;
    for (int m0 = 0; m0 < M; m0 += TS) {
        const int ts = M - m0 < TS ? M - m0 : TS;
        __syncthreads();
        for (int idx = tid; idx < N * TS; idx += blockDim.x) {
            const int n = idx / TS, s = idx - n * TS;
            if (s < ts) {
                xs[n * TP + s] = rc[(size_t)(n * dim) * M + m0 + s];
                if (dim > 1) ys[n * TP + s] = rc[(size_t)(n * dim + 1) * M + m0 + s];
            }
        }
        __syncthreads();
        // rows i and N-2-i are folded onto each other so that every pass over j has N-1 partners in total (load balance)
        for (int ii = 0; ii < (N - 1 + 1) / 2; ++ii) {
            const int ia = ii, ib = N - 2 - ii;              // row ia pairs with j > ia (N-1-ia partners), row ib with j > ib (ia+1 partners)
            const int na = N - 1 - ia, nbp = ib > ia ? N - 1 - ib : 0;
            for (int t = tid; t < na + nbp; t += blockDim.x) {
                const int i = t < na ? ia : ib, j = t < na ? ia + 1 + t : ib + 1 + (t - na);
                for (int s = 0; s < ts; ++s) {
                    double dx = fabs(xs[i * TP + s] - xs[j * TP + s]); { const double alt = twoL - dx; dx = alt < dx ? alt : dx; }   // distance(), propagator.jl:6-9
                    double d2 = dx * dx;
                    if (dim > 1) { double dy = fabs(ys[i * TP + s] - ys[j * TP + s]); const double alt = twoL - dy; dy = alt < dy ? alt : dy; d2 = d2 + dy * dy; }
                    const double d = sqrt(d2);
                    // floor(d / bin) through the reciprocal, re-evaluated with the IEEE division next to an integer (identical integers)
                    const double q = d * inv; double f = floor(q);
                    const double fr = q - f, thr = (q + 1.0) * 4e-15;
                    if (!(fr >= thr && 1.0 - fr >= thr)) f = floor(d / bin);
                    if (f < (double)nb) {
                        if (smem_hist) atomicAdd(&hist[(int)f], 1u);
                        else atomicAdd(&G.hist[(int)f], 1ull);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (smem_hist) for (int i = tid; i < nb; i += blockDim.x) if (hist[i]) atomicAdd(&G.hist[i], (unsigned long long)hist[i]);
}

// ---- winding number: W_k = (1 / 2L) sum over the links of teleport(r_next[k] - r[k], L); one CTA per chain, one warp per worldline,
// lanes stride the slices (coalesced).  Streams the positions once: HBM-bound (16 B per bead).
__global__ void __launch_bounds__(256) k_winding(const __grid_constant__ DevSys S, const __grid_constant__ WiDev Wd, long long k)
{
    __shared__ double red[2][8];
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, N = S.N, M = S.M, dim = S.dim;
    const double *rc = S.r + (size_t)c * N * dim * M;
    const int *nextc = S.next + (size_t)c * N;
    const double L = S.L, twoL = 2 * S.L;
    double wx = 0.0, wy = 0.0;
    for (int n = warp; n < N; n += nw) {
        const double *rx = rc + (size_t)(n * dim) * M, *ry = rx + M;
        const double *qx = rc + (size_t)(nextc[n] * dim) * M, *qy = qx + M;
        for (int j = lane; j < M; j += 32) {
            const double bx = j == M - 1 ? qx[0] : rx[j + 1];
            wx += d_teleport(bx - rx[j], L);
            if (dim > 1) { const double by = j == M - 1 ? qy[0] : ry[j + 1]; wy += d_teleport(by - ry[j], L); }
        }
    }
    wx = e_warp_sum(wx); wy = e_warp_sum(wy);
    if (lane == 0) { red[0][warp] = wx; red[1][warp] = wy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        wx = 0.0; wy = 0.0;
        for (int i = 0; i < nw; ++i) { wx += red[0][i]; wy += red[1][i]; }
        if (k < Wd.cap) {
            Wd.W[((size_t)k * dim + 0) * S.C + c] = wx / twoL;
            if (dim > 1) Wd.W[((size_t)k * dim + 1) * S.C + c] = wy / twoL;
        }
    }
}

// ---- static structure factor S(k) = <|rho_k|^2> / N, rho_k(m) = sum_n exp(i k . r_n[m]) on the wave vectors of the periodic box
// k = (pi / L)(a, b), a = 0..kmax, |b| <= kmax (half plane: a > 0, or a = 0 and b > 0; the other half is the complex conjugate; 1-D: b = 0).
// One CTA per chain; tiles of TS slices staged in shared memory like k_paircorr (positions read once: 16 B per bead); warp w owns a = w, its
// lanes stride the particles of one slice after the other.  Per bead: exp(i a th_x) by one sincospi, the table exp(i b th_y), b = 0..kmax, by
// one sincospi and a complex-product recurrence; rho(a, b) and rho(a, -b) share the four real sums A = cx cy, B = sx sy, C = cx sy, D = sx cy:
// rho(a, b) = (A - B, C + D), rho(a, -b) = (A + B, D - C).  The sums over the slices are added to the chain's own accumulators by one lane
// (no atomics: bit-reproducible).  Bound: fp64 pipe (~200 flops per bead and value of a).
__global__ void __launch_bounds__(256) k_structure(const __grid_constant__ DevSys S, const __grid_constant__ SkDev K, int TS)
{
    extern __shared__ double sm[];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, a = tid >> 5, N = S.N, M = S.M, dim = S.dim, kmax = K.kmax;
    const int TP = TS + 1;
    double *xs = sm, *ys = sm + (size_t)N * TP;
    const double *rc = S.r + (size_t)c * N * dim * M;
    const double invL = 1.0 / S.L;
    double accp[PIMC_SK_KMAX + 1], accm[PIMC_SK_KMAX + 1];   // sums over the slices of |rho(a, b)|^2 and |rho(a, -b)|^2
#pragma unroll
    for (int b = 0; b <= PIMC_SK_KMAX; ++b) { accp[b] = 0.0; accm[b] = 0.0; }
    for (int m0 = 0; m0 < M; m0 += TS) {
        const int ts = M - m0 < TS ? M - m0 : TS;
        __syncthreads();
        for (int idx = tid; idx < N * TS; idx += blockDim.x) {
            const int n = idx / TS, s = idx - n * TS;
            if (s < ts) {
                xs[n * TP + s] = rc[(size_t)(n * dim) * M + m0 + s];
                if (dim > 1) ys[n * TP + s] = rc[(size_t)(n * dim + 1) * M + m0 + s];
            }
        }
        __syncthreads();
        if (a > kmax) continue;
        for (int s = 0; s < ts; ++s) {
            double A[PIMC_SK_KMAX + 1], B[PIMC_SK_KMAX + 1], Cc[PIMC_SK_KMAX + 1], D[PIMC_SK_KMAX + 1];
#pragma unroll
            for (int b = 0; b <= PIMC_SK_KMAX; ++b) { A[b] = 0.0; B[b] = 0.0; Cc[b] = 0.0; D[b] = 0.0; }
            for (int n = lane; n < N; n += 32) {
                double sx, cx, s1, c1;
                sincospi((double)a * (xs[n * TP + s] * invL), &sx, &cx);
                if (dim > 1) sincospi(ys[n * TP + s] * invL, &s1, &c1); else { s1 = 0.0; c1 = 1.0; }
                double cy = 1.0, sy = 0.0;
#pragma unroll
                for (int b = 0; b <= PIMC_SK_KMAX; ++b) {
                    if (b <= kmax) {
                        A[b] += cx * cy; B[b] += sx * sy; Cc[b] += cx * sy; D[b] += sx * cy;
                        const double cn = cy * c1 - sy * s1, sn = sy * c1 + cy * s1;
                        cy = cn; sy = sn;
                    }
                }
            }
#pragma unroll
            for (int b = 0; b <= PIMC_SK_KMAX; ++b) {
                if (b <= kmax) {
                    const double Ar = e_warp_sum(A[b]), Br = e_warp_sum(B[b]), Cr = e_warp_sum(Cc[b]), Dr = e_warp_sum(D[b]);
                    const double pr = Ar - Br, pi_ = Cr + Dr, mr = Ar + Br, mi = Dr - Cr;
                    accp[b] += pr * pr + pi_ * pi_;
                    accm[b] += mr * mr + mi * mi;
                }
            }
        }
    }
    if (a <= kmax && lane == 0) {
        const int nb = 2 * kmax + 1;
        double *out = K.S + ((size_t)c * (kmax + 1) + a) * nb;
#pragma unroll
        for (int b = 0; b <= PIMC_SK_KMAX; ++b) {
            if (b > kmax) continue;
            const bool plus = dim == 1 ? (b == 0 && a > 0) : (a > 0 || b > 0);     // the half plane of independent wave vectors
            const bool minus = dim > 1 && a > 0 && b > 0;
            if (plus) out[kmax + b] += accp[b];
            if (minus) out[kmax - b] += accm[b];
        }
    }
}

size_t pimc_paircorr_smem(const DevSys &S, const PcDev &G, int *TS, int *smem_hist)
{
    *smem_hist = G.nbins <= 8192 ? 1 : 0;
    const size_t hb = *smem_hist ? (size_t)G.nbins * sizeof(unsigned) : 0;
    int ts = 8;
    while (ts > 1 && (size_t)S.dim * S.N * (ts + 1) * sizeof(double) + hb > 160 * 1024) ts >>= 1;
    *TS = ts;
    return (size_t)S.dim * S.N * (ts + 1) * sizeof(double) + hb + 16;
}
cudaError_t pimc_launch_paircorr(int grid, cudaStream_t st, const DevSys &S, const PcDev &G)
{
    int TS, sh; const size_t smem = pimc_paircorr_smem(S, G, &TS, &sh);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) { cudaError_t e = cudaFuncSetAttribute(k_paircorr, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e != cudaSuccess) return e; configured = true; }
    k_paircorr<<<grid, 256, smem, st>>>(S, G, TS, sh);
    return cudaGetLastError();
}
cudaError_t pimc_launch_winding(int grid, cudaStream_t st, const DevSys &S, const WiDev &W, long long k)
{
    k_winding<<<grid, 256, 0, st>>>(S, W, k);
    return cudaGetLastError();
}
size_t pimc_structure_smem(const DevSys &S, int *TS)
{
    int ts = 8;
    while (ts > 1 && (size_t)S.dim * S.N * (ts + 1) * sizeof(double) > 160 * 1024) ts >>= 1;
    *TS = ts;
    return (size_t)S.dim * S.N * (ts + 1) * sizeof(double) + 16;
}
cudaError_t pimc_launch_structure(int grid, cudaStream_t st, const DevSys &S, const SkDev &K)
{
    int TS; const size_t smem = pimc_structure_smem(S, &TS);
    if (smem > 200 * 1024 || K.kmax < 1 || K.kmax > PIMC_SK_KMAX) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) { cudaError_t e = cudaFuncSetAttribute(k_structure, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e != cudaSuccess) return e; configured = true; }
    k_structure<<<grid, 256, smem, st>>>(S, K, TS);
    return cudaGetLastError();
}
