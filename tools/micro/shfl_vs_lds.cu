// Microbenchmark: do warp shuffles compete with shared-memory loads for the same SM data path?
// Per SM: one CTA of 256 threads (8 warps) runs N iterations of (a) 16 LDS.128, (b) 64 SHFL.32, (c) both, (d) DFMA only,
// (e) DFMA + LDS, (f) DFMA + SHFL.  Reports cycles per iteration.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, long long *cycles, int n)
{
    __shared__ double2 buf[256 + 8];
    const int tid = threadIdx.x;
    buf[tid + 4] = make_double2(tid, -tid);
    if (tid < 4) { buf[tid] = make_double2(0, 0); buf[260 + tid] = make_double2(0, 0); }
    __syncthreads();
    double2 acc = make_double2(out[tid], out[tid + 256]);
    double f0 = tid * 1e-3, f1 = 1.0, f2 = 2.0, f3 = 3.0, f4 = 0.5, f5 = 0.25, f6 = 0.125, f7 = 4.0;
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 4) {
            const unsigned base = (unsigned)__cvta_generic_to_shared(buf + tid + 4);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double2 a, b, c, d;
                asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2 + -32];" : "=d"(a.x), "=d"(a.y) : "r"(base));
                asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2 + -16];" : "=d"(b.x), "=d"(b.y) : "r"(base));
                asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2 + 16];" : "=d"(c.x), "=d"(c.y) : "r"(base));
                asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2 + 32];" : "=d"(d.x), "=d"(d.y) : "r"(base));
                acc.x += a.x + b.x + c.x + d.x;
                acc.y += a.y + b.y + c.y + d.y;
            }
        }
        if (MODE == 1 || MODE == 2 || MODE == 5) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double2 v = acc;
                double ax = __shfl_up_sync(0xffffffffu, v.x, 2), ay = __shfl_up_sync(0xffffffffu, v.y, 2);
                double bx = __shfl_up_sync(0xffffffffu, v.x, 1), by = __shfl_up_sync(0xffffffffu, v.y, 1);
                double cx = __shfl_down_sync(0xffffffffu, v.x, 1), cy = __shfl_down_sync(0xffffffffu, v.y, 1);
                double dx = __shfl_down_sync(0xffffffffu, v.x, 2), dy = __shfl_down_sync(0xffffffffu, v.y, 2);
                acc.x += ax + bx + cx + dx;
                acc.y += ay + by + cy + dy;
            }
        }
        if (MODE >= 3) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                f0 = fma(f0, 1.0000001, 1e-9); f1 = fma(f1, 1.0000001, 1e-9); f2 = fma(f2, 1.0000001, 1e-9); f3 = fma(f3, 1.0000001, 1e-9);
                f4 = fma(f4, 1.0000001, 1e-9); f5 = fma(f5, 1.0000001, 1e-9); f6 = fma(f6, 1.0000001, 1e-9); f7 = fma(f7, 1.0000001, 1e-9);
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + tid] = acc.x + acc.y + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 256 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int n = 20000;
    const char *names[] = {"16 LDS.128 (+8 DADD)", "64 SHFL.32 (+8 DADD)", "LDS + SHFL", "128 DFMA", "128 DFMA + 16 LDS.128", "128 DFMA + 64 SHFL.32"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
            case 0: k<0><<<148, 256>>>(out, cyc, n); break;
            case 1: k<1><<<148, 256>>>(out, cyc, n); break;
            case 2: k<2><<<148, 256>>>(out, cyc, n); break;
            case 3: k<3><<<148, 256>>>(out, cyc, n); break;
            case 4: k<4><<<148, 256>>>(out, cyc, n); break;
            case 5: k<5><<<148, 256>>>(out, cyc, n); break;
            }
            cudaDeviceSynchronize();
        }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("{\"mode\": \"%s\", \"cycles_per_iteration_8_warps\": %.1f}\n", names[mode], (double)h[0] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
