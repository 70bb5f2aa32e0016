// micro-benchmark: throughput of fp64 rounding / conversion instructions (FRND.F64, F2I.F64, I2F.F64) against DADD on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, int iters, double seed)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 0.37 + i * 1.13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = a[i] + 1.000001;                                  // DADD
            if (OP == 1) a[i] = floor(a[i]) + 0.37;                               // FRND.F64.FLOOR + DADD
            if (OP == 2) a[i] = (double)((int)a[i] & 1023) + 0.37;                // F2I.F64 + LOP + I2F.F64 + DADD
            if (OP == 3) { double r = (a[i] + 6755399441055744.0) - 6755399441055744.0; a[i] = (r > a[i] ? r - 1.0 : r) + 0.37; }   // magic floor
            if (OP == 4) a[i] = sqrt(a[i]) + 1.37;                                // MUFU.RSQ64H + Newton
        }
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, double ops_per_iter)
{
    double *d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    k<OP><<<148 * 8, 256>>>(d, 64, 1.5);
    cudaEventRecord(e0); k<OP><<<148 * 8, 256>>>(d, iters, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * 8 * 256 * iters * 8;
    printf("%-28s %8.3f ms  %7.2f lane-iterations/clk/SM (1.965 GHz)\n", name, ms, lane_ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() { run<0>("DADD", 1); run<1>("floor (FRND.F64) + DADD", 1); run<2>("F2I.F64 + I2F.F64 + DADD", 1); run<3>("magic floor (4 fp64 ops)", 1); run<4>("sqrt + DADD", 1); return 0; }
