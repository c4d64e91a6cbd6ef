// kernels_1d.cu -- radial (1D) kernels for sm_100a.
//
// rk4_1d_resident: the whole time loop of one radial system runs inside ONE CTA.  Each thread owns
// PPT consecutive grid nodes and keeps psi, the RK accumulator, c12*P and its rows of the operator
// table in registers for all `iters` steps; per RK stage the only data that moves are the K edge
// nodes each thread publishes to (and reads from) shared memory, laid out [slot][thread] so that
// consecutive lanes touch consecutive 16-byte words (conflict-free LDS.128/STS.128).  Two
// shared-memory buffers alternate between stages, so one __syncthreads per stage is enough.  HBM is
// touched once at the start (psi, P, taps) and once at the end (psi): zero bytes per step.
// A batch of independent systems (the ensemble mode) maps one system to one CTA (blockIdx.x).
//
// Reference semantics: runge_kutta (nls.f90:705-734) calling hamiltonian (:621-650) with the band
// matvec rgbmv (:530-541); stage arguments u + k*dt/2 (k3: *dt), update
// u + (k1 + 2 k2 + 2 k3 + k4)*dt/6.

#include "device_math.cuh"
#include "diag_acc.cuh"
#include "kernels.h"

#include <cstdlib>

namespace nlsb {

namespace {

constexpr int kResidentThreads = 256;   // 255 registers per thread available: taps stay in registers

// MINBLOCKS = 2 (<= 128 registers, two CTAs per SM) feeds the FP64 pipe better when an ensemble fills
// the GPU; a lone system runs faster with all 255 registers (MINBLOCKS = 1, no spills).
//
// DIAG: the first stage of the LAST step holds k1 = H(psi) of the state entering that step; its scalar diagnostics
// (diag_acc.cuh: chemical-potential sums with weight r = i dx, damping integral and particle number with the area
// element 2 pi r dx on the reference's linspace grid, peak density / reservoir) are reduced there and written to
// out8[member] -- the convergence loop of tools/check.py without a second pass over the field.
template <int M, int PPT, bool TAPS_IN_REGS, int MINBLOCKS, bool DIAG>
__global__ void __launch_bounds__(kResidentThreads, MINBLOCKS)
rk4_1d_resident(int n, int iters, double dt, const double *__restrict__ taps, const double *__restrict__ pumping,
                const double *__restrict__ coeffs, double2 *__restrict__ psi, double dx, double *__restrict__ out8)
{
    constexpr int K = (M - 1) / 2;
    static_assert(PPT >= K, "a thread must own at least K nodes so halos come from adjacent threads only");
    extern __shared__ double2 halo[];   // [2 buffers][2K slots][T + 2]

    const int T = blockDim.x, tid = threadIdx.x;
    const int pitch = T + 2;
    const size_t member = blockIdx.x;
    const int i0 = tid * PPT;

    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);
    const double *P = pumping + member * n;
    double2 *u_g = psi + member * n;

    double tap[TAPS_IN_REGS ? PPT : 1][M];
    double cp[PPT];
    double2 u[PPT], y[PPT], acc[PPT];
    bool live[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int i = i0 + p;
        live[p] = i < n;
        u[p] = live[p] ? u_g[i] : make_double2(0.0, 0.0);
        cp[p] = live[p] ? c.c12 * P[i] : 0.0;
        if (TAPS_IN_REGS) {
#pragma unroll
            for (int t = 0; t < M; ++t) tap[p][t] = live[p] ? taps[(size_t)i * M + t] : 0.0;
        }
        y[p] = u[p];
        acc[p] = make_double2(0.0, 0.0);
    }

    // zero the pads (thread -1 and thread T) of both buffers once
    for (int q = tid; q < 2 * 2 * K; q += T) {
        halo[q * pitch] = make_double2(0.0, 0.0);
        halo[q * pitch + T + 1] = make_double2(0.0, 0.0);
    }

    const double half_dt = dt / 2, dt6 = dt / 6;
    DiagAcc dacc = diag_zero();
    const double ring = n > 1 ? (n * dx) / (n - 1) : 0.0;              // spacing of linspace(0, n dx, n)

    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            double2 *buf = halo + (s & 1) * (2 * K) * pitch;
            // publish my first K nodes (slots 0..K-1) and my last K nodes (slots K..2K-1)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                buf[j * pitch + tid + 1] = y[j];
                buf[(K + j) * pitch + tid + 1] = y[PPT - K + j];
            }
            __syncthreads();
            double2 w[PPT + 2 * K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                w[j] = buf[(K + j) * pitch + tid];            // left neighbour's last K nodes
                w[K + PPT + j] = buf[j * pitch + tid + 2];    // right neighbour's first K nodes
            }
#pragma unroll
            for (int p = 0; p < PPT; ++p) w[K + p] = y[p];

#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                // the operator row WITHOUT its centre tap: rhs_point_c folds that one into the pointwise part
                double lr = 0.0, li = 0.0, tap0 = 0.0;
#pragma unroll
                for (int t = 0; t < M; ++t) {
                    // large systems re-read their operator rows through L1 instead of pinning registers
                    const double a = TAPS_IN_REGS ? tap[p][t] : (live[p] ? __ldg(taps + (size_t)(i0 + p) * M + t) : 0.0);
                    if (t == K) {
                        tap0 = a;
                        continue;
                    }
                    lr = fma(a, w[p + t].x, lr);
                    li = fma(a, w[p + t].y, li);
                }
                // Nodes beyond the end of the system need no mask: their psi, c12*P and operator rows are zero, so their k is
                // exactly zero (rhs_point of zeros) and they stay zero -- the right edge sees the truncated band matrix.
                // (On B200 every non-FP64 instruction costs this FP64-bound loop an issue cycle: tools/micro/fp64_issue.cu.)
                const double2 k = rhs_point_c(c, cp[p], w[K + p], tap0, lr, li);
                if (DIAG && s == 0 && it == iters - 1 && live[p]) {
                    const int i = i0 + p;
                    diag_accumulate(dacc, c, cp[p], w[K + p], k, ((double)(i + 1) - 1.0) * dx,      // nls.f90:940-947
                                    6.283185307179586 * (i * ring) * dx);                           // nls/model.py:357-361
                }
                if (s == 0) {
                    acc[p] = k;
                    y[p].x = fma(k.x, half_dt, u[p].x);
                    y[p].y = fma(k.y, half_dt, u[p].y);
                } else if (s == 1) {
                    acc[p].x = fma(2.0, k.x, acc[p].x);
                    acc[p].y = fma(2.0, k.y, acc[p].y);
                    y[p].x = fma(k.x, half_dt, u[p].x);
                    y[p].y = fma(k.y, half_dt, u[p].y);
                } else if (s == 2) {
                    acc[p].x = fma(2.0, k.x, acc[p].x);
                    acc[p].y = fma(2.0, k.y, acc[p].y);
                    y[p].x = fma(k.x, dt, u[p].x);
                    y[p].y = fma(k.y, dt, u[p].y);
                } else {
                    u[p].x = fma(acc[p].x + k.x, dt6, u[p].x);
                    u[p].y = fma(acc[p].y + k.y, dt6, u[p].y);
                    y[p] = u[p];
                }
            }
        }
    }

#pragma unroll
    for (int p = 0; p < PPT; ++p)
        if (live[p]) u_g[i0 + p] = u[p];
    if (DIAG) {
        __syncthreads();                      // the halo buffers are free: stage them for the block reduction
        const DiagAcc total = diag_block_reduce(dacc, reinterpret_cast<DiagAcc *>(halo));
        if (tid == 0) {
            double *o = out8 + member * 8;
#pragma unroll
            for (int i = 0; i < kDiagSums; ++i) o[i] = total.s[i];
            o[6] = total.m[0];
            o[7] = total.m[1];
        }
    }
}

template <int M>
__global__ void hamiltonian_1d_kernel(int n, const double *__restrict__ taps, const double *__restrict__ pumping,
                                      const double *__restrict__ coeffs, const double2 *__restrict__ u,
                                      double2 *__restrict__ v)
{
    constexpr int K = (M - 1) / 2;
    const int i = blockIdx.y * blockDim.x + threadIdx.x;    // members on grid.x (no 65535 limit), node blocks on grid.y
    const size_t member = blockIdx.x;
    if (i >= n) return;
    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);
    const double2 *um = u + member * n;
    double lr = 0.0, li = 0.0;
#pragma unroll
    for (int t = 0; t < M; ++t) {
        const int j = i + t - K;
        if (j >= 0 && j < n) {
            const double a = taps[(size_t)i * M + t];
            const double2 x = um[j];
            lr = fma(a, x.x, lr);
            li = fma(a, x.y, li);
        }
    }
    v[member * n + i] = rhs_point(c, c.c12 * pumping[member * n + i], um[i], lr, li);
}

// Fallback for systems too large to live in one CTA (n > kMaxResident1D): one launch per RK stage,
// stage inputs and the accumulator go through global memory (same dataflow as the 2D stage kernel).
template <int M>
__global__ void stage_1d_kernel(int n, int mode, const double *__restrict__ taps, const double *__restrict__ pumping,
                                const double *__restrict__ coeffs, const double2 *__restrict__ ysrc,
                                const double2 *__restrict__ ubase, double2 *__restrict__ acc,
                                double2 *__restrict__ ydst, double cy, double dt6)
{
    constexpr int K = (M - 1) / 2;
    const int i = blockIdx.y * blockDim.x + threadIdx.x;    // members on grid.x (no 65535 limit), node blocks on grid.y
    const size_t member = blockIdx.x;
    if (i >= n) return;
    const size_t g = member * n + i;
    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);
    const double2 *ym = ysrc + member * n;
    double lr = 0.0, li = 0.0;
#pragma unroll
    for (int t = 0; t < M; ++t) {
        const int j = i + t - K;
        if (j >= 0 && j < n) {
            const double a = taps[(size_t)i * M + t];
            const double2 x = ym[j];
            lr = fma(a, x.x, lr);
            li = fma(a, x.y, li);
        }
    }
    const double2 k = rhs_point(c, c.c12 * pumping[g], ym[i], lr, li);
    const double2 u = ubase[g];
    if (mode == kStageFirst) {
        acc[g] = k;
        ydst[g] = make_double2(fma(k.x, cy, u.x), fma(k.y, cy, u.y));
    } else if (mode == kStageMid) {
        const double2 a = acc[g];
        acc[g] = make_double2(fma(2.0, k.x, a.x), fma(2.0, k.y, a.y));
        ydst[g] = make_double2(fma(k.x, cy, u.x), fma(k.y, cy, u.y));
    } else {
        const double2 a = acc[g];
        ydst[g] = make_double2(fma(a.x + k.x, dt6, u.x), fma(a.y + k.y, dt6, u.y));
    }
}

template <int M>
__global__ void band_matvec_1d_kernel(int n, const double *__restrict__ taps, const double *__restrict__ x,
                                      double *__restrict__ u, double sign)
{
    constexpr int K = (M - 1) / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < M; ++t) {
        const int j = i + t - K;
        if (j >= 0 && j < n) s = fma(taps[(size_t)i * M + t], x[j], s);
    }
    u[i] = fma(sign, s, u[i]);
}

template <int M, int PPT, bool TAPS_IN_REGS, int MINBLOCKS, bool DIAG>
int launch_resident_d(int batch, int n, int iters, double dt, const double *taps, const double *pumping,
                      const double *coeffs, double2 *psi, double dx, double *out8, cudaStream_t stream)
{
    constexpr int K = (M - 1) / 2;
    int threads = (n + PPT - 1) / PPT;
    threads = (threads + 31) / 32 * 32;
    size_t smem = sizeof(double2) * 2 * 2 * K * (threads + 2);
    if (DIAG && smem < sizeof(DiagAcc) * (kResidentThreads / 32)) smem = sizeof(DiagAcc) * (kResidentThreads / 32);
    cudaError_t e = cudaFuncSetAttribute(rk4_1d_resident<M, PPT, TAPS_IN_REGS, MINBLOCKS, DIAG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    rk4_1d_resident<M, PPT, TAPS_IN_REGS, MINBLOCKS, DIAG><<<batch, threads, smem, stream>>>(n, iters, dt, taps, pumping,
                                                                                            coeffs, psi, dx, out8);
    count_launches(1);
    return (int)cudaGetLastError();
}

template <int M, int PPT, bool TAPS_IN_REGS, int MINBLOCKS>
int launch_resident(int batch, int n, int iters, double dt, const double *taps, const double *pumping,
                    const double *coeffs, double2 *psi, double dx, double *out8, cudaStream_t stream)
{
    if (out8 && iters > 0)
        return launch_resident_d<M, PPT, TAPS_IN_REGS, MINBLOCKS, true>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, out8, stream);
    return launch_resident_d<M, PPT, TAPS_IN_REGS, MINBLOCKS, false>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, out8, stream);
}

template <int M>
int launch_resident_m(int batch, int n, int iters, double dt, const double *taps, const double *pumping,
                      const double *coeffs, double2 *psi, double dx, double *out8, cudaStream_t stream)
{
    static const int force_one = [] {
        const char *e = std::getenv("NLSB_1D_ONE_CTA");        // tuning knob: 1 = always the one-CTA-per-SM flavour
        return e ? std::atoi(e) : 0;
    }();
    if (n <= 4 * kResidentThreads) {
        if (M <= 5 && batch >= 2 * 148 && !force_one)   // enough members for two CTAs on every SM
            return launch_resident<M, 4, true, (M <= 5) ? 2 : 1>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, out8, stream);
        return launch_resident<M, 4, true, 1>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, out8, stream);
    }
    return launch_resident<M, 8, false, 1>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, out8, stream);
}

}  // namespace

int launch_rk4_1d(int batch, int n, int order, int iters, double dt, const double *taps, const double *pumping,
                  const double *coeffs, double2 *psi, cudaStream_t stream, double dx, double *diag_out8)
{
    if (n > kMaxResident1D)
        return fail(NLSB_ESIZE, "resident 1D kernel handles n <= %d (n = %d): use launch_rk4_1d_staged", kMaxResident1D, n);
    switch (order) {
    case 3: return launch_resident_m<3>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, diag_out8, stream);
    case 5: return launch_resident_m<5>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, diag_out8, stream);
    case 7: return launch_resident_m<7>(batch, n, iters, dt, taps, pumping, coeffs, psi, dx, diag_out8, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

int launch_rk4_1d_staged(int batch, int n, int order, int iters, double dt, const double *taps,
                         const double *pumping, const double *coeffs, double2 *psi, double2 *work,
                         cudaStream_t stream)
{
    if (order != 3 && order != 5 && order != 7) return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    const size_t np = (size_t)batch * n;
    double2 *ya = work, *yb = work + np, *acc = work + 2 * np;
    const dim3 block(128), grid(batch, (n + 127) / 128);
    auto stage = [&](int mode, const double2 *ysrc, double2 *ydst, double cy) {
        switch (order) {
        case 3: stage_1d_kernel<3><<<grid, block, 0, stream>>>(n, mode, taps, pumping, coeffs, ysrc, psi, acc, ydst, cy, dt / 6); break;
        case 5: stage_1d_kernel<5><<<grid, block, 0, stream>>>(n, mode, taps, pumping, coeffs, ysrc, psi, acc, ydst, cy, dt / 6); break;
        default: stage_1d_kernel<7><<<grid, block, 0, stream>>>(n, mode, taps, pumping, coeffs, ysrc, psi, acc, ydst, cy, dt / 6); break;
        }
    };
    for (int it = 0; it < iters; ++it) {
        stage(kStageFirst, psi, ya, dt / 2);
        stage(kStageMid, ya, yb, dt / 2);
        stage(kStageMid, yb, ya, dt);
        stage(kStageLast, ya, psi, 0.0);
    }
    count_launches(4ull * (unsigned long long)iters);
    return (int)cudaGetLastError();
}

int launch_hamiltonian_1d(int batch, int n, int order, const double *taps, const double *pumping,
                          const double *coeffs, const double2 *u, double2 *v, cudaStream_t stream)
{
    const dim3 block(128), grid(batch, (n + 127) / 128);
    switch (order) {
    case 3: hamiltonian_1d_kernel<3><<<grid, block, 0, stream>>>(n, taps, pumping, coeffs, u, v); break;
    case 5: hamiltonian_1d_kernel<5><<<grid, block, 0, stream>>>(n, taps, pumping, coeffs, u, v); break;
    case 7: hamiltonian_1d_kernel<7><<<grid, block, 0, stream>>>(n, taps, pumping, coeffs, u, v); break;
    default: return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    }
    count_launches(1);
    return (int)cudaGetLastError();
}

int launch_band_matvec_1d(int n, int order, const double *taps, const double *x, double *u, double sign,
                          cudaStream_t stream)
{
    const dim3 block(128), grid((n + 127) / 128);
    switch (order) {
    case 1: band_matvec_1d_kernel<1><<<grid, block, 0, stream>>>(n, taps, x, u, sign); break;
    case 3: band_matvec_1d_kernel<3><<<grid, block, 0, stream>>>(n, taps, x, u, sign); break;
    case 5: band_matvec_1d_kernel<5><<<grid, block, 0, stream>>>(n, taps, x, u, sign); break;
    case 7: band_matvec_1d_kernel<7><<<grid, block, 0, stream>>>(n, taps, x, u, sign); break;
    default: return fail(NLSB_EORDER, "band width must be 1, 3, 5 or 7 (got %d)", order);
    }
    count_launches(1);
    return (int)cudaGetLastError();
}

}  // namespace nlsb
