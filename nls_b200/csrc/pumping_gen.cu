// pumping_gen.cu -- pumping profiles of an ensemble generated ON THE DEVICE: C3 / C5-sized
// ensembles otherwise build and upload one full profile per member (0.5 - 2 GB) in numpy on every solve()
// (nls/solver.py:32, nls/model.py:220-232).
//
// Families (the leaf profiles the reference's examples and the BASELINE configs use, nls/pumping.py:113-178):
//   kind 0  GaussianPumping[1D|2D]   P = power exp(-((x-x0)^2 + (y-y0)^2) / (2 var^2))          (1D: y = 0.0)
//   kind 1  GaussianRingPumping1D    P = 1.0 (G(x; +R) + G(x; -R)),  G(x; c) = power exp(-((x-c)^2 + 0^2)/(2 var^2))
//           GaussianRingPumping2D    the same on radii = sqrt((x-x0)^2 + (y-y0)^2)
// params: [batch][5] = {power, x0, y0, variation, radius}.
//
// Bit-exactness policy: the grid (numpy.linspace: arange(n) * step + start, last point = stop) and every
// multiply / add / divide / sqrt of the expression tree are performed in the reference's order with
// round-to-nearest and no contraction, so they are bit-identical to numpy's; the only deviation is exp(), where
// CUDA's and glibc's double-precision exp may differ in the last place.  tests/test_gpu_engine.py holds the
// generated profiles to 4 ulp of the host classes (nls_b200/pumping.py, themselves bit-exact to the reference).

#include "kernels.h"

namespace nlsb {

namespace {

__device__ __forceinline__ double linspace_at(int i, int n, double start, double stop, double step)
{
    return i == n - 1 && n > 1 ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
}

__device__ __forceinline__ double gauss(double power, double dx, double dy, double variation)
{
    // power * exp(-(dx**2 + dy**2) / (2.0 * variation**2))        (nls/pumping.py:126)
    const double num = -__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double den = __dmul_rn(2.0, __dmul_rn(variation, variation));
    return __dmul_rn(power, exp(__ddiv_rn(num, den)));
}

__device__ __forceinline__ double ring(double power, double radius, double variation, double x)
{
    // OpSumPumping of two GaussianPumping1D(power, +-radius, 0.0, variation) called with y = 0.0 (nls/pumping.py:54, :139, :156-159)
    const double lhs = gauss(power, __dsub_rn(x, radius), __dsub_rn(0.0, 0.0), variation);
    const double rhs = gauss(power, __dsub_rn(x, -radius), __dsub_rn(0.0, 0.0), variation);
    return __dmul_rn(1.0, __dadd_rn(lhs, rhs));
}

__global__ void pumping_2d_kernel(int kind, int n, double start, double stop, double step,
                                  const double *__restrict__ params, double *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;     // out[i][j]: x = grid[j], y = grid[i]
    if (j >= n) return;
    const double *p = params + 5 * (size_t)blockIdx.z;
    const double x = linspace_at(j, n, start, stop, step), y = linspace_at(i, n, start, stop, step);
    double v;
    if (kind == 0) {
        v = gauss(p[0], __dsub_rn(x, p[1]), __dsub_rn(y, p[2]), p[3]);
    } else {
        const double ex = __dsub_rn(x, p[1]), ey = __dsub_rn(y, p[2]);
        v = ring(p[0], p[4], p[3], __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey))));   // nls/pumping.py:173-175
    }
    out[((size_t)blockIdx.z * n + i) * n + j] = v;
}

__global__ void pumping_1d_kernel(int kind, int n, double start, double stop, double step,
                                  const double *__restrict__ params, double *__restrict__ out)
{
    const int j = blockIdx.y * blockDim.x + threadIdx.x;      // members on the x axis: ensembles exceed 65535
    if (j >= n) return;
    const double *p = params + 5 * (size_t)blockIdx.x;
    const double x = linspace_at(j, n, start, stop, step);
    out[(size_t)blockIdx.x * n + j] = kind == 0 ? gauss(p[0], __dsub_rn(x, p[1]), __dsub_rn(0.0, p[2]), p[3])
                                                : ring(p[0], p[4], p[3], x);
}

}  // namespace

// dim 1: x = linspace(0, n dx, n) (nls/model.py:222-224); dim 2: x = y = linspace(-n dx / 2, n dx / 2, n) (:228-230)
int launch_pumping_profiles(int dim, int kind, int batch, int n, double dx, const double *params_dev, double *out,
                            cudaStream_t stream)
{
    if (n > 65535 || (dim == 2 && batch > 65535))
        return fail(NLSB_ESIZE, "pumping profiles: n (and the batch of 2D profiles) must not exceed 65535");
    const double right = dim == 1 ? n * dx : n * dx / 2;
    const double start = dim == 1 ? 0.0 : -right, stop = right;
    const double step = n > 1 ? (stop - start) / (n - 1) : 0.0;
    if (dim == 1) {
        const dim3 grid(batch, (n + 127) / 128);
        pumping_1d_kernel<<<grid, 128, 0, stream>>>(kind, n, start, stop, step, params_dev, out);
    } else {
        const dim3 grid((n + 127) / 128, n, batch);
        pumping_2d_kernel<<<grid, 128, 0, stream>>>(kind, n, start, stop, step, params_dev, out);
    }
    count_launches(1);
    return (int)cudaGetLastError();
}

}  // namespace nlsb
