// Microbenchmark: do non-FP64 instructions take issue opportunities away from the FP64 pipe?
// Per SM one CTA of 256 threads (2 warps per scheduler) runs N iterations of 128 independent-chain DFMA plus
// (mode 1) 64 IADD3-class integer ops, (mode 2) 128 integer ops, (mode 3) 128 FP32 FFMA, (mode 4) 128 MOV-like selects.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, long long *cycles, int n, int seed)
{
    const int tid = threadIdx.x;
    double f0 = tid * 1e-3, f1 = 1.0, f2 = 2.0, f3 = 3.0, f4 = 0.5, f5 = 0.25, f6 = 0.125, f7 = 4.0;
    int i0 = tid + seed, i1 = tid * 3 + seed, i2 = tid * 5 + seed, i3 = tid * 7 + seed;
    float g0 = tid * 0.5f, g1 = 1.5f, g2 = 2.5f, g3 = 3.5f;
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            f0 = fma(f0, 1.0000001, 1e-9); f1 = fma(f1, 1.0000001, 1e-9); f2 = fma(f2, 1.0000001, 1e-9); f3 = fma(f3, 1.0000001, 1e-9);
            f4 = fma(f4, 1.0000001, 1e-9); f5 = fma(f5, 1.0000001, 1e-9); f6 = fma(f6, 1.0000001, 1e-9); f7 = fma(f7, 1.0000001, 1e-9);
            if (MODE == 1 || MODE == 2) {
                i0 = i0 * 3 + i1; i1 = i1 ^ (i2 + 7); i2 = i2 * 5 + i3; i3 = i3 ^ (i0 + 11);
                if (MODE == 2) { i0 = i0 + (i1 >> 3); i1 = i1 * 9 + i2; i2 = i2 ^ (i3 << 2); i3 = i3 + i0 * 13; }
            }
            if (MODE == 3) {
                g0 = fmaf(g0, 1.0001f, 1e-5f); g1 = fmaf(g1, 1.0001f, 1e-5f); g2 = fmaf(g2, 1.0001f, 1e-5f); g3 = fmaf(g3, 1.0001f, 1e-5f);
                g0 = fmaf(g0, 0.9999f, 1e-5f); g1 = fmaf(g1, 0.9999f, 1e-5f); g2 = fmaf(g2, 0.9999f, 1e-5f); g3 = fmaf(g3, 0.9999f, 1e-5f);
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + tid] = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + i0 + i1 + i2 + i3 + g0 + g1 + g2 + g3;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 256 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int n = 20000;
    const char *names[] = {"128 DFMA", "128 DFMA + 64 INT", "128 DFMA + 128 INT", "128 DFMA + 128 FFMA"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
            case 0: k<0><<<148, 256>>>(out, cyc, n, rep); break;
            case 1: k<1><<<148, 256>>>(out, cyc, n, rep); break;
            case 2: k<2><<<148, 256>>>(out, cyc, n, rep); break;
            case 3: k<3><<<148, 256>>>(out, cyc, n, rep); break;
            }
            cudaDeviceSynchronize();
        }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("{\"mode\": \"%s\", \"cycles_per_iteration_2_warps_per_scheduler\": %.1f}\n", names[mode], (double)h[0] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
