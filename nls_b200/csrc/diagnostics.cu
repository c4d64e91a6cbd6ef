// diagnostics.cu -- the scalar diagnostics of a device-resident state in ONE pass per member.
//
// What the reference computes on the host after copying the field back:
//   * chemical potential  mu = i <u, H u> / <u, u>   (chemical_potential_1d/_2d, nls.f90:921-971; weight r = i dx in
//     the radial case, 1 on the square; the stencil order is the model's here, the reference hard-wires 5)
//   * damping integral    sum (n - 1) |u|^2 dA        (Solution.getDampingIntegral, nls/model.py:350-365; dA = dx^2 on
//     the square, 2 pi r dx with r = linspace(0, n dx, n) in the radial case)
//   * particle number     sum |u|^2 dA,  peak density max |u|^2,  peak reservoir max n   (getDensity / getReservoir,
//     nls/model.py:367-380)
// The kernels evaluate v = H(u) on the fly (nothing is materialised), reduce with warp shuffles and a fixed
// shared-memory tree, write one partial per CTA and finish per member in a second launch: the summation tree
// depends only on (n, grid), so the numbers are run-to-run reproducible.
//
// Output per member: 8 doubles {Re M, Im M, Re E, Im E, damping, particles, max |u|^2, max reservoir} with
// M = sum w conj(u) u, E = sum w conj(u) v; mu = i E / M.

#include "device_math.cuh"
#include "diag_acc.cuh"
#include "kernels.h"

namespace nlsb {

namespace {

constexpr int kThreads = 256;
constexpr int kSums = kDiagSums;
using Acc = DiagAcc;

__device__ __forceinline__ Acc acc_zero() { return diag_zero(); }

__device__ __forceinline__ Acc block_reduce(Acc a)
{
    __shared__ Acc part[kThreads / 32];
    return diag_block_reduce(a, part);   // valid in thread 0
}

__device__ __forceinline__ void accumulate(Acc &a, const RhsCoeffs &c, double cp, double2 u, double2 v, double w, double wd)
{
    diag_accumulate(a, c, cp, u, v, w, wd);
}

// A CTA's reduced sums: a partial for the finishing launch, or -- when the member has a single CTA -- the result.
__device__ __forceinline__ void emit(const Acc &a, size_t member, Acc *partial, double *out8)
{
    if (gridDim.y > 1) {
        partial[member * gridDim.y + blockIdx.y] = a;
        return;
    }
    double *o = out8 + member * 8;
#pragma unroll
    for (int i = 0; i < kSums; ++i) o[i] = a.s[i];
    o[6] = a.m[0];
    o[7] = a.m[1];
}

template <int K>
struct WeightsK {
    double wx[2 * K + 1];
    double wy[2 * K + 1];
};

// A CTA walks over 64 x 16-node tiles (thread = one column, four rows of the tile): the y neighbours of a tile
// stay in L1 instead of being fetched from L2 once per row.
constexpr int kTileW = 64, kTileH = 16;

template <int K>
__global__ void __launch_bounds__(kThreads)
diagnostics_2d_kernel(int rows, int cols, double area, WeightsK<K> w, const double *__restrict__ pumping,
                      const double *__restrict__ coeffs, const double2 *__restrict__ u, Acc *__restrict__ partial,
                      double *__restrict__ out8)
{
    const size_t member = blockIdx.x;
    const size_t plane = (size_t)rows * cols;
    const double2 *src = u + member * plane;
    const double *P = pumping + member * plane;
    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);
    Acc a = acc_zero();
    const int tiles_x = (cols + kTileW - 1) / kTileW, tiles_y = (rows + kTileH - 1) / kTileH;
    const int tx = threadIdx.x % kTileW, ty = threadIdx.x / kTileW;
    for (int tile = blockIdx.y; tile < tiles_x * tiles_y; tile += gridDim.y) {
        const int x = (tile % tiles_x) * kTileW + tx;
        if (x >= cols) continue;
        for (int y = (tile / tiles_x) * kTileH + ty; y < rows && y < (tile / tiles_x + 1) * kTileH; y += kThreads / kTileW) {
            const size_t t = (size_t)y * cols + x;
            const double2 centre = src[t];
            double lr = w.wx[K] * centre.x, li = w.wx[K] * centre.y;       // same tap order as stage_2d_kernel
#pragma unroll
            for (int s = 1; s <= K; ++s) {
                if (x - s >= 0) {
                    const double2 v = src[t - s];
                    lr = fma(w.wx[K - s], v.x, lr);
                    li = fma(w.wx[K - s], v.y, li);
                }
                if (x + s < cols) {
                    const double2 v = src[t + s];
                    lr = fma(w.wx[K + s], v.x, lr);
                    li = fma(w.wx[K + s], v.y, li);
                }
                if (y - s >= 0) {
                    const double2 v = src[t - (size_t)s * cols];
                    lr = fma(w.wy[K - s], v.x, lr);
                    li = fma(w.wy[K - s], v.y, li);
                }
                if (y + s < rows) {
                    const double2 v = src[t + (size_t)s * cols];
                    lr = fma(w.wy[K + s], v.x, lr);
                    li = fma(w.wy[K + s], v.y, li);
                }
            }
            const double cp = c.c12 * P[t];
            accumulate(a, c, cp, centre, rhs_point(c, cp, centre, lr, li), 1.0, area);
        }
    }
    a = block_reduce(a);
    if (threadIdx.x == 0) emit(a, member, partial, out8);
}

template <int M>
__global__ void __launch_bounds__(kThreads)
diagnostics_1d_kernel(int n, double dx, const double *__restrict__ taps, const double *__restrict__ pumping,
                      const double *__restrict__ coeffs, const double2 *__restrict__ u, Acc *__restrict__ partial,
                      double *__restrict__ out8)
{
    constexpr int K = (M - 1) / 2;
    const size_t member = blockIdx.x;          // members on the x axis of the grid: ensembles exceed 65535
    const double2 *um = u + member * n;
    const double *P = pumping + member * n;
    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);
    const double ring = n > 1 ? (n * dx) / (n - 1) : 0.0;              // spacing of linspace(0, n dx, n)
    Acc a = acc_zero();
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        double lr = 0.0, li = 0.0;
#pragma unroll
        for (int t = 0; t < M; ++t) {
            const int j = i + t - K;
            if (j >= 0 && j < n) {
                const double tap = taps[(size_t)i * M + t];
                lr = fma(tap, um[j].x, lr);
                li = fma(tap, um[j].y, li);
            }
        }
        const double cp = c.c12 * P[i];
        const double w = ((double)(i + 1) - 1.0) * dx;                  // chemical_potential_1d, nls.f90:940-947
        const double wd = 6.283185307179586 * (i * ring) * dx;          // getDampingIntegral, nls/model.py:357-361
        accumulate(a, c, cp, um[i], rhs_point(c, cp, um[i], lr, li), w, wd);
    }
    a = block_reduce(a);
    if (threadIdx.x == 0) emit(a, member, partial, out8);
}

// Large ensembles of short systems (C3: 65 536 x 1000): ONE WARP per member -- nodes strided over the lanes (coalesced
// 512-byte loads), a shuffle reduction and no barrier, eight members per CTA -- instead of a CTA per member whose 256
// threads see four nodes each and then pay a block reduction.
template <int M>
__global__ void __launch_bounds__(kThreads)
diagnostics_1d_warp_kernel(int batch, int n, double dx, const double *__restrict__ taps, const double *__restrict__ pumping,
                           const double *__restrict__ coeffs, const double2 *__restrict__ u, double *__restrict__ out8)
{
    constexpr int K = (M - 1) / 2;
    const int lane = threadIdx.x & 31;
    const size_t member = (size_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (member >= (size_t)batch) return;                   // whole warps leave: no barrier below
    const double2 *um = u + member * n;
    const double *P = pumping + member * n;
    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);
    const double ring = n > 1 ? (n * dx) / (n - 1) : 0.0;
    Acc a = acc_zero();
    for (int i = lane; i < n; i += 32) {
        double lr = 0.0, li = 0.0;
#pragma unroll
        for (int t = 0; t < M; ++t) {
            const int j = i + t - K;
            if (j >= 0 && j < n) {
                const double tap = taps[(size_t)i * M + t];
                lr = fma(tap, um[j].x, lr);
                li = fma(tap, um[j].y, li);
            }
        }
        const double cp = c.c12 * P[i];
        const double w = ((double)(i + 1) - 1.0) * dx;
        const double wd = 6.283185307179586 * (i * ring) * dx;
        accumulate(a, c, cp, um[i], rhs_point(c, cp, um[i], lr, li), w, wd);
    }
    a = diag_warp_reduce(a);
    if (lane == 0) {
        double *o = out8 + member * 8;
#pragma unroll
        for (int i = 0; i < kSums; ++i) o[i] = a.s[i];
        o[6] = a.m[0];
        o[7] = a.m[1];
    }
}

__global__ void __launch_bounds__(kThreads)
finish_diagnostics_kernel(int nparts, const Acc *__restrict__ partial, double *__restrict__ out8)
{
    const size_t member = blockIdx.x;
    Acc a = acc_zero();
    for (int t = threadIdx.x; t < nparts; t += blockDim.x) {
        const Acc p = partial[member * nparts + t];
#pragma unroll
        for (int i = 0; i < kSums; ++i) a.s[i] += p.s[i];
        a.m[0] = fmax(a.m[0], p.m[0]);
        a.m[1] = fmax(a.m[1], p.m[1]);
    }
    a = block_reduce(a);
    if (threadIdx.x == 0) {
        double *o = out8 + member * 8;
#pragma unroll
        for (int i = 0; i < kSums; ++i) o[i] = a.s[i];
        o[6] = a.m[0];
        o[7] = a.m[1];
    }
}

int parts_for(size_t units, int batch)      // units: CTA-sized pieces of work of one member
{
    size_t blocks = units;
    // enough CTAs to fill the GPU across the batch (592 = 4 per SM), at most 592 partials per member; a large
    // ensemble gets one CTA per member (one block reduction per member instead of one per 256 nodes)
    size_t cap = batch >= 592 ? 1 : 592 / (size_t)batch + 1;
    if (cap > 592) cap = 592;
    if (blocks > cap) blocks = cap;
    return blocks ? (int)blocks : 1;
}

template <int K>
WeightsK<K> pack(const CrossWeights &w)
{
    WeightsK<K> p;
    for (int t = 0; t < 2 * K + 1; ++t) {
        p.wx[t] = w.wx[t];
        p.wy[t] = w.wy[t];
    }
    return p;
}

}  // namespace

// Sum `parts` partials per member in a fixed order (second pass of every diagnostics reduction).
int launch_finish_diagnostics(int batch, int parts, const void *partial, double *out8, cudaStream_t stream)
{
    finish_diagnostics_kernel<<<(unsigned)batch, kThreads, 0, stream>>>(parts, static_cast<const Acc *>(partial), out8);
    count_launches(1);
    return (int)cudaGetLastError();
}

size_t diagnostics_scratch_bytes(int batch) { return sizeof(Acc) * 592 * (size_t)(batch > 0 ? batch : 1); }

int launch_diagnostics_2d(int batch, int rows, int cols, int order, double dx, const CrossWeights &w,
                          const double *pumping, const double *coeffs, const double2 *u, void *scratch, double *out8,
                          cudaStream_t stream)
{
    const int parts = parts_for((size_t)((cols + kTileW - 1) / kTileW) * ((rows + kTileH - 1) / kTileH), batch);
    const dim3 grid((unsigned)batch, (unsigned)parts);
    Acc *partial = static_cast<Acc *>(scratch);
    switch (order) {
    case 3: diagnostics_2d_kernel<1><<<grid, kThreads, 0, stream>>>(rows, cols, dx * dx, pack<1>(w), pumping, coeffs, u, partial, out8); break;
    case 5: diagnostics_2d_kernel<2><<<grid, kThreads, 0, stream>>>(rows, cols, dx * dx, pack<2>(w), pumping, coeffs, u, partial, out8); break;
    case 7: diagnostics_2d_kernel<3><<<grid, kThreads, 0, stream>>>(rows, cols, dx * dx, pack<3>(w), pumping, coeffs, u, partial, out8); break;
    default: return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    }
    if (parts > 1) finish_diagnostics_kernel<<<(unsigned)batch, kThreads, 0, stream>>>(parts, partial, out8);
    count_launches(parts > 1 ? 2 : 1);
    return (int)cudaGetLastError();
}

int launch_diagnostics_1d(int batch, int n, int order, double dx, const double *taps, const double *pumping,
                          const double *coeffs, const double2 *u, void *scratch, double *out8, cudaStream_t stream)
{
    const int parts = parts_for(((size_t)n + kThreads - 1) / kThreads, batch);
    if (parts == 1 && batch >= 8 * 592 && n <= 4096) {      // a warp per member keeps every SM busy and needs no barrier
        const unsigned ctas = (unsigned)((batch + kThreads / 32 - 1) / (kThreads / 32));
        switch (order) {
        case 3: diagnostics_1d_warp_kernel<3><<<ctas, kThreads, 0, stream>>>(batch, n, dx, taps, pumping, coeffs, u, out8); break;
        case 5: diagnostics_1d_warp_kernel<5><<<ctas, kThreads, 0, stream>>>(batch, n, dx, taps, pumping, coeffs, u, out8); break;
        case 7: diagnostics_1d_warp_kernel<7><<<ctas, kThreads, 0, stream>>>(batch, n, dx, taps, pumping, coeffs, u, out8); break;
        default: return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
        }
        count_launches(1);
        return (int)cudaGetLastError();
    }
    const dim3 grid((unsigned)batch, (unsigned)parts);
    Acc *partial = static_cast<Acc *>(scratch);
    switch (order) {
    case 3: diagnostics_1d_kernel<3><<<grid, kThreads, 0, stream>>>(n, dx, taps, pumping, coeffs, u, partial, out8); break;
    case 5: diagnostics_1d_kernel<5><<<grid, kThreads, 0, stream>>>(n, dx, taps, pumping, coeffs, u, partial, out8); break;
    case 7: diagnostics_1d_kernel<7><<<grid, kThreads, 0, stream>>>(n, dx, taps, pumping, coeffs, u, partial, out8); break;
    default: return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    }
    if (parts > 1) finish_diagnostics_kernel<<<(unsigned)batch, kThreads, 0, stream>>>(parts, partial, out8);
    count_launches(parts > 1 ? 2 : 1);
    return (int)cudaGetLastError();
}

}  // namespace nlsb
