// reduce.cu -- deterministic device reductions for the diagnostics of the hot path.
//
// weighted_dots computes, in one pass over u0 and v,
//     M = sum conj(u0) * u0 * w      E = sum conj(u0) * v * w
// which is all chemical_potential_1d/_2d need (nls.f90:945-947, :967-970: mu = i*E / M, with
// w = r = (i-1)*dx in the radial case and w = 1 on the square).  Warp-shuffle tree inside a warp,
// shared memory across the warps of a CTA, one partial per CTA, and a single-CTA second pass: the
// summation tree is fixed by (npts, grid), so results are run-to-run reproducible.

#include "kernels.h"

namespace nlsb {

namespace {

constexpr int kThreads = 256;

struct Dots {
    double m_re, m_im, e_re, e_im;
};

__device__ __forceinline__ Dots warp_sum(Dots d)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        d.m_re += __shfl_down_sync(0xffffffffu, d.m_re, off);
        d.m_im += __shfl_down_sync(0xffffffffu, d.m_im, off);
        d.e_re += __shfl_down_sync(0xffffffffu, d.e_re, off);
        d.e_im += __shfl_down_sync(0xffffffffu, d.e_im, off);
    }
    return d;
}

__device__ __forceinline__ Dots block_sum(Dots d)
{
    __shared__ Dots part[kThreads / 32];
    d = warp_sum(d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = d;
    __syncthreads();
    Dots t = {0.0, 0.0, 0.0, 0.0};
    if (warp == 0) {
        if (lane < kThreads / 32) t = part[lane];
        t = warp_sum(t);
    }
    return t;   // valid in thread 0
}

// radial_dx > 0: weight of node i is i*radial_dx (1D); radial_dx == 0: weight 1 (2D)
__global__ void __launch_bounds__(kThreads)
weighted_dots_kernel(size_t npts, double radial_dx, const double2 *__restrict__ u0, const double2 *__restrict__ v,
                     Dots *__restrict__ partial)
{
    Dots d = {0.0, 0.0, 0.0, 0.0};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < npts; t += stride) {
        const double2 a = u0[t], b = v[t];
        const double w = radial_dx > 0.0 ? ((double)(t + 1) - 1.0) * radial_dx : 1.0;
        // conj(a) * (x * w)
        const double ar = a.x * w, ai = a.y * w, br = b.x * w, bi = b.y * w;
        d.m_re += a.x * ar + a.y * ai;
        d.m_im += a.x * ai - a.y * ar;
        d.e_re += a.x * br + a.y * bi;
        d.e_im += a.x * bi - a.y * br;
    }
    d = block_sum(d);
    if (threadIdx.x == 0) partial[blockIdx.x] = d;
}

__global__ void __launch_bounds__(kThreads)
finish_dots_kernel(int nparts, const Dots *__restrict__ partial, double *__restrict__ out4)
{
    Dots d = {0.0, 0.0, 0.0, 0.0};
    for (int t = threadIdx.x; t < nparts; t += blockDim.x) {
        const Dots p = partial[t];
        d.m_re += p.m_re;
        d.m_im += p.m_im;
        d.e_re += p.e_re;
        d.e_im += p.e_im;
    }
    d = block_sum(d);
    if (threadIdx.x == 0) {
        out4[0] = d.m_re;
        out4[1] = d.m_im;
        out4[2] = d.e_re;
        out4[3] = d.e_im;
    }
}

}  // namespace

size_t weighted_dots_scratch_bytes() { return sizeof(Dots) * 1024 + 4 * sizeof(double); }

// out4 (device, 4 doubles) = {Re M, Im M, Re E', Im E'} with E' = sum conj(u0) v w; scratch from
// weighted_dots_scratch_bytes().
int launch_weighted_dots(size_t npts, double radial_dx, const double2 *u0, const double2 *v, void *scratch,
                         double *out4, cudaStream_t stream)
{
    size_t blocks = (npts + kThreads - 1) / kThreads;
    if (blocks > 1024) blocks = 1024;
    if (blocks == 0) blocks = 1;
    Dots *partial = static_cast<Dots *>(scratch);
    weighted_dots_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(npts, radial_dx, u0, v, partial);
    finish_dots_kernel<<<1, kThreads, 0, stream>>>((int)blocks, partial, out4);
    count_launches(2);
    return (int)cudaGetLastError();
}

}  // namespace nlsb
