// diag_acc.cuh -- accumulator of the scalar diagnostics (shared by diagnostics.cu and by the RK kernels that fuse the
// reduction into their first stage, where v = H(u) of the state entering the step is available for free).
//
// What the reference computes on the host after copying the field back:
//   * chemical potential  mu = i <u, H u> / <u, u>   (chemical_potential_1d/_2d, nls.f90:921-971)
//   * damping integral    sum (n - 1) |u|^2 dA        (Solution.getDampingIntegral, nls/model.py:350-365)
//   * particle number     sum |u|^2 dA,  peak density max |u|^2,  peak reservoir max n   (nls/model.py:367-380)
// Per member: 8 doubles {Re M, Im M, Re E, Im E, damping, particles, max |u|^2, max reservoir},
// M = sum w conj(u) u, E = sum w conj(u) v, mu = i E / M.
#pragma once

#include "internal.h"
#include "device_math.cuh"

#include <cuda_runtime.h>

namespace nlsb {

constexpr int kDiagSums = 6;

struct DiagAcc {
    double s[kDiagSums];     // M_re, M_im, E_re, E_im, damping, particles
    double m[2];             // max |u|^2, max reservoir
};

__host__ __device__ __forceinline__ DiagAcc diag_zero()
{
    DiagAcc a;
#pragma unroll
    for (int i = 0; i < kDiagSums; ++i) a.s[i] = 0.0;
    a.m[0] = a.m[1] = 0.0;
    return a;
}

// One node's contribution: u the field, v = H(u), w the weight of the dot products, wd the area element.
__host__ __device__ __forceinline__ void diag_accumulate(DiagAcc &a, const RhsCoeffs &c, double cp, double2 u, double2 v, double w,
                                                         double wd)
{
    const double ur = u.x * w, ui = u.y * w, vr = v.x * w, vi = v.y * w;     // conj(u) * (x * w), as reduce.cu
    a.s[0] += u.x * ur + u.y * ui;
    a.s[1] += u.x * ui - u.y * ur;
    a.s[2] += u.x * vr + u.y * vi;
    a.s[3] += u.x * vi - u.y * vr;
    // |u|^2 and the reservoir exactly as the right-hand side forms them (device_math.cuh::rhs_abm): inside a
    // time-stepping kernel the compiler then reuses that step's values instead of dividing a second time (the IEEE
    // divide is ~20 FP64 instructions with a branch; div_fast is within 1 ulp of it on the ABI's coefficient domain)
    const double usq = fma(u.x, u.x, u.y * u.y);
    const double res = div_fast(cp, fma(c.c14, usq, c.c13));                   // getReservoir, nls/model.py:376-380
    a.s[4] += (res - 1.0) * usq * wd;
    a.s[5] += usq * wd;
    a.m[0] = fmax(a.m[0], usq);
    a.m[1] = fmax(a.m[1], res);
}

#if defined(__CUDACC__)
__device__ __forceinline__ DiagAcc diag_warp_reduce(DiagAcc a)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < kDiagSums; ++i) a.s[i] += __shfl_down_sync(0xffffffffu, a.s[i], off);
#pragma unroll
        for (int i = 0; i < 2; ++i) a.m[i] = fmax(a.m[i], __shfl_down_sync(0xffffffffu, a.m[i], off));
    }
    return a;
}

// Fixed-order reduction over the CTA (warp shuffles, then one warp over the per-warp partials staged in `part`,
// which holds at least blockDim.x / 32 entries); the result is valid in thread 0.  Contains a __syncthreads.
__device__ __forceinline__ DiagAcc diag_block_reduce(DiagAcc a, DiagAcc *part)
{
    a = diag_warp_reduce(a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) part[warp] = a;
    __syncthreads();
    DiagAcc t = diag_zero();
    if (warp == 0) {
        if (lane < nwarps) t = part[lane];
        t = diag_warp_reduce(t);
    }
    return t;
}
#endif

}  // namespace nlsb
