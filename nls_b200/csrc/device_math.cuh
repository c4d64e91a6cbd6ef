// device_math.cuh -- pointwise right-hand side shared by every kernel.
#pragma once

#include "internal.h"

#include <cuda_runtime.h>

namespace nlsb {

// Pointwise part of v = H(u) plus the Laplacian terms (nls.f90:637-647 / :857-867):
//     n   = c12*P / (c13 + c14*|u|^2)                       reservoir, closed form (nls.f90:580)
//     a   = c3*n - c4 ,  b = c5*|u|^2 + c6*n
//     v   = (a*u_re + b*u_im - L u_im) + i (a*u_im - b*u_re + L u_re)
// `cp` is the product c12*P (the reference's own left-to-right association, formed once per solve).
// The divide is div_fast below (never more than 1 ulp from the reference's correctly rounded divide, equal to it
// for all but ~1 operand pair in 10^7).
// a / b without the branchy IEEE slow path: hardware reciprocal seed (MUFU.RCP64H: the high word of 1/b, relative
// error <= 2^-23), ONE Newton step (error^2 <= 2^-40), then one residual correction of the quotient
// (q = a x; r = a - b q exactly, by FMA; q + r x), which squares the error once more -- 5 dependent FP64
// operations and no control flow, so the scheduler can interleave the chains of the nodes a thread works on.
// The corrected quotient is a/b (1 + d) with |d| ~ (seed error)^4 ~ 2^-76 before its final rounding: it is the
// correctly rounded quotient unless a/b lies within 2^-76 of a rounding boundary.  Measured on B200 against
// __ddiv_rn (tests/test_gpu_engine.py::test_fast_divide_against_the_ieee_divide_on_its_domain, 2^24 operand pairs of
// the reservoir's domain incl. denominators one ulp around powers of two): maximum difference 1 ulp, 1 pair of
// 16.8 million differs.  Two more FP64 operations per divide (a second Newton step) would remove those cases at
// +5 % FP64 work per step; the 1e-10 parity bar does not need them.
// Precondition: b is a finite, normal number (the reservoir denominator c13 + c14 |psi|^2 is >= c13 = 1 for every
// model the host layer builds); a zero, infinite or NaN denominator yields NaN where IEEE division would yield
// +-inf / 0.
__host__ __device__ __forceinline__ double div_fast(double a, double b)
{
    double x;
#if defined(__CUDA_ARCH__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
#else
    x = 1.0 / b;   // host emulation of the kernels (tests/emu): same refinement, different seed
#endif
    const double e = fma(-b, x, 1.0);
    x = fma(x, e, x);
    const double q = a * x;
    const double r = fma(-b, q, a);
    return fma(r, x, q);
}

// The two halves of rhs_point: the coefficients a, b depend on the node's own value only (they can be formed
// before the node's neighbours are known), the second half folds in the Laplacian.
__host__ __device__ __forceinline__ void rhs_ab(const RhsCoeffs &c, double cp, double2 u, double &a, double &b)
{
    const double usq = fma(u.x, u.x, u.y * u.y);
    const double res = div_fast(cp, fma(c.c14, usq, c.c13));
    a = fma(c.c3, res, -c.c4);
    b = fma(c.c5, usq, c.c6 * res);
}

__host__ __device__ __forceinline__ double2 rhs_apply(double a, double b, double2 u, double lap_re, double lap_im)
{
    double2 v;
    v.x = fma(a, u.x, fma(b, u.y, -lap_im));
    v.y = fma(a, u.y, fma(-b, u.x, lap_re));
    return v;
}

__host__ __device__ __forceinline__ double2 rhs_point(const RhsCoeffs &c, double cp, double2 u, double lap_re, double lap_im)
{
    double a, b;
    rhs_ab(c, cp, u, a, b);
    return rhs_apply(a, b, u, lap_re, lap_im);
}

// The time-stepping kernels fold the CENTRE tap w0 of the Laplacian into b:
//     v = (a u_re + (b - w0) u_im - L' u_im) + i (a u_im - (b - w0) u_re + L' u_re),   L' = L without its centre tap,
// with bm = b - w0 = c5 |u|^2 + (c6 n - w0) formed by the two FMAs b took before -- two FP64 instructions fewer per
// node and stage (the centre products of the real and the imaginary part), 8 of 147 per RK step in 2D, 8 of 115 in 1D.
// These kernels are FP64-issue bound, so the instruction count is the time.  The result differs from
// the unmerged form (rhs_point, kept for hamiltonian / hamiltonian_2d and the stand-alone diagnostics) in the last
// place only: w0 u is rounded inside the FMA chain either way, at the same magnitude.
__host__ __device__ __forceinline__ void rhs_abm(const RhsCoeffs &c, double cp, double2 u, double w0, double &a, double &bm)
{
    const double usq = fma(u.x, u.x, u.y * u.y);
    const double res = div_fast(cp, fma(c.c14, usq, c.c13));
    a = fma(c.c3, res, -c.c4);
    bm = fma(c.c5, usq, fma(c.c6, res, -w0));
}

// `lap_re` / `lap_im`: the Laplacian WITHOUT its centre tap; w0: the centre tap.
__host__ __device__ __forceinline__ double2 rhs_point_c(const RhsCoeffs &c, double cp, double2 u, double w0, double lap_re,
                                                        double lap_im)
{
    double a, bm;
    rhs_abm(c, cp, u, w0, a, bm);
    return rhs_apply(a, bm, u, lap_re, lap_im);
}

__host__ __device__ __forceinline__ RhsCoeffs load_rhs_coeffs(const double *__restrict__ coeffs23)
{
    RhsCoeffs c;
    c.c3 = coeffs23[2];
    c.c4 = coeffs23[3];
    c.c5 = coeffs23[4];
    c.c6 = coeffs23[5];
    c.c12 = coeffs23[11];
    c.c13 = coeffs23[12];
    c.c14 = coeffs23[13];
    return c;
}

}  // namespace nlsb
