// kernels_2d.cu -- per-stage 2D kernels for sm_100a (one launch per RK stage).
//
// stage_2d_kernel evaluates k = H(y) on a rows x cols grid (cols contiguous) with the separable
// cross stencil of make_laplacian_2d (nls.f90:297-385; truncated = zero outside the square) and
// folds the RK4 stage algebra of runge_kutta_2d (nls.f90:892-899) into the same pass:
//     first  : acc = k,         y' = u + cy*k
//     middle : acc += 2k,       y' = u + cy*k
//     last   : u  = u + (acc + k)*dt/6
//     rhs    : v  = k                                  (hamiltonian_2d, nls.f90:841-870)
// Every thread owns one complex128 node: a warp reads 512 contiguous bytes per row, the x-taps hit
// the same L1 lines, the y-taps are L1/L2 hits from the neighbouring rows of the CTA tile.
// This is the simple 4-pass formulation (expected DRAM traffic 56+88+88+72 B per node-step, see
// DESIGN.md); the fused-step kernel in fused_2d.cu is the fast path for whole time loops.

#include "device_math.cuh"
#include "kernels.h"

namespace nlsb {

namespace {

template <int K>
struct WeightsK {
    double wx[2 * K + 1];
    double wy[2 * K + 1];
};

template <int K, int MODE>
__global__ void __launch_bounds__(256)
stage_2d_kernel(int rows, int cols, WeightsK<K> w, const double2 *__restrict__ ysrc,
                const double2 *__restrict__ ubase, const double *__restrict__ pumping,
                const double *__restrict__ coeffs, double2 *__restrict__ acc, double2 *__restrict__ ydst,
                double cy, double dt6)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const size_t member = blockIdx.z;
    const size_t plane = (size_t)rows * cols;
    const size_t g = member * plane + (size_t)y * cols + x;
    const double2 *src = ysrc + member * plane;
    const RhsCoeffs c = load_rhs_coeffs(coeffs + member * 23);

    const double2 centre = src[(size_t)y * cols + x];
    double lr = w.wx[K] * centre.x, li = w.wx[K] * centre.y;
#pragma unroll
    for (int s = 1; s <= K; ++s) {
        if (x - s >= 0) {
            const double2 v = src[(size_t)y * cols + (x - s)];
            lr = fma(w.wx[K - s], v.x, lr);
            li = fma(w.wx[K - s], v.y, li);
        }
        if (x + s < cols) {
            const double2 v = src[(size_t)y * cols + (x + s)];
            lr = fma(w.wx[K + s], v.x, lr);
            li = fma(w.wx[K + s], v.y, li);
        }
        if (y - s >= 0) {
            const double2 v = src[(size_t)(y - s) * cols + x];
            lr = fma(w.wy[K - s], v.x, lr);
            li = fma(w.wy[K - s], v.y, li);
        }
        if (y + s < rows) {
            const double2 v = src[(size_t)(y + s) * cols + x];
            lr = fma(w.wy[K + s], v.x, lr);
            li = fma(w.wy[K + s], v.y, li);
        }
    }
    const double2 k = rhs_point(c, c.c12 * pumping[g], centre, lr, li);

    if (MODE == kStageRhs) {
        ydst[g] = k;
    } else if (MODE == kStageFirst) {
        const double2 u = centre;   // ysrc == ubase in the first stage
        acc[g] = k;
        ydst[g] = make_double2(fma(k.x, cy, u.x), fma(k.y, cy, u.y));
    } else if (MODE == kStageMid) {
        const double2 u = ubase[g], a = acc[g];
        acc[g] = make_double2(fma(2.0, k.x, a.x), fma(2.0, k.y, a.y));
        ydst[g] = make_double2(fma(k.x, cy, u.x), fma(k.y, cy, u.y));
    } else {
        const double2 u = ubase[g], a = acc[g];
        ydst[g] = make_double2(fma(a.x + k.x, dt6, u.x), fma(a.y + k.y, dt6, u.y));
    }
}

template <int K>
__global__ void __launch_bounds__(256)
cross_matvec_2d_kernel(int rows, int cols, WeightsK<K> w, const double *__restrict__ xin, double *__restrict__ yio,
                       double sign)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    double s = w.wx[K] * xin[(size_t)y * cols + x];
#pragma unroll
    for (int d = 1; d <= K; ++d) {
        if (x - d >= 0) s = fma(w.wx[K - d], xin[(size_t)y * cols + (x - d)], s);
        if (x + d < cols) s = fma(w.wx[K + d], xin[(size_t)y * cols + (x + d)], s);
        if (y - d >= 0) s = fma(w.wy[K - d], xin[(size_t)(y - d) * cols + x], s);
        if (y + d < rows) s = fma(w.wy[K + d], xin[(size_t)(y + d) * cols + x], s);
    }
    const size_t g = (size_t)y * cols + x;
    yio[g] = fma(sign, s, yio[g]);
}

// ---- general block-band operator ---------------------------------------------------------------------------------
// rbbmv (nls.f90:408-527) accepts ANY blocks memory of the make_laplacian_2d layout: m - 1 off-diagonal blocks that
// are diagonals of n entries (weight of line i + s at position p: blk_b[p], s = b - k) and a middle block that is an
// m x n band along the line (A(p, q) = band(k + p - q, q), BLAS GB storage).  The fast kernels take the constant-weight
// cross stencil only (blocks_to_weights); an operator whose entries vary along the line -- variable coefficients,
// anisotropy, a user's own boundary rows -- runs through the kernels below: same right-hand side and stage algebra,
// the weights read from the blocks memory on the device.  One node per thread, no tiling: a compatibility path.
//
// Orientation: `line` is the index rbbmv cuts the flattened vector by, `pos` the index inside a line; element
// (line, pos) sits at line * n + pos.  The (n, n) arrays of hamiltonian_2d / runge_kutta_2d arrive at the C ABI in the
// reference's (Fortran) memory order, whose lines are the slow index as well (nls_b200/native.py hands them over so).
struct GeneralOp {
    const double *blocks;    // device copy of the blocks memory, n (2m - 1) doubles
    int m, n;
};

__device__ __forceinline__ const double *general_block(const GeneralOp &op, int b)
{
    const int k = (op.m - 1) / 2;
    const size_t n = (size_t)op.n;
    return op.blocks + (b <= k ? (size_t)b * n : (size_t)k * n + (size_t)op.m * n + (size_t)(b - k - 1) * n);
}

template <typename T, typename Load>
__device__ __forceinline__ void general_apply(const GeneralOp &op, int line, int pos, Load load, T &acc)
{
    const int k = (op.m - 1) / 2, n = op.n;
    for (int b = 0; b < op.m; ++b) {
        if (b == k) {
            const double *band = general_block(op, k);
            const int lo = pos - k > 0 ? pos - k : 0, hi = pos + k < n - 1 ? pos + k : n - 1;
            for (int q = lo; q <= hi; ++q) acc.add(band[(size_t)(k + pos - q) + (size_t)op.m * q], load(line, q));
        } else {
            const int src = line + (b - k);
            if (src >= 0 && src < n) acc.add(general_block(op, b)[pos], load(src, pos));
        }
    }
}

struct RealAcc {
    double s = 0.0;
    __device__ __forceinline__ void add(double w, double v) { s = fma(w, v, s); }
};
struct ComplexAcc {
    double re = 0.0, im = 0.0;
    __device__ __forceinline__ void add(double w, double2 v)
    {
        re = fma(w, v.x, re);
        im = fma(w, v.y, im);
    }
};

__global__ void __launch_bounds__(256)
general_matvec_2d_kernel(GeneralOp op, const double *__restrict__ xin, double *__restrict__ yio, double sign)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, bidx = blockIdx.y * blockDim.y + threadIdx.y;
    if (a >= op.n || bidx >= op.n) return;
    const size_t g = (size_t)bidx * op.n + a;                    // memory order: a (= pos) is the contiguous index
    const int line = bidx, pos = a;
    RealAcc acc;
    general_apply(op, line, pos, [&](int l, int p) { return xin[(size_t)l * op.n + p]; }, acc);
    yio[g] = fma(sign, acc.s, yio[g]);
}

template <int MODE>
__global__ void __launch_bounds__(256)
stage_2d_general_kernel(GeneralOp op, const double2 *__restrict__ ysrc, const double2 *__restrict__ ubase,
                        const double *__restrict__ pumping, const double *__restrict__ coeffs, double2 *__restrict__ acc,
                        double2 *__restrict__ ydst, double cy, double dt6)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, bidx = blockIdx.y * blockDim.y + threadIdx.y;
    if (a >= op.n || bidx >= op.n) return;
    const size_t g = (size_t)bidx * op.n + a;
    const int line = bidx, pos = a;
    const RhsCoeffs c = load_rhs_coeffs(coeffs);
    ComplexAcc lap;
    general_apply(op, line, pos, [&](int l, int p) { return ysrc[(size_t)l * op.n + p]; }, lap);
    const double2 centre = ysrc[g];
    const double2 k = rhs_point(c, c.c12 * pumping[g], centre, lap.re, lap.im);
    if (MODE == kStageRhs) {
        ydst[g] = k;
    } else if (MODE == kStageFirst) {
        acc[g] = k;
        ydst[g] = make_double2(fma(k.x, cy, centre.x), fma(k.y, cy, centre.y));
    } else if (MODE == kStageMid) {
        const double2 u = ubase[g], s = acc[g];
        acc[g] = make_double2(fma(2.0, k.x, s.x), fma(2.0, k.y, s.y));
        ydst[g] = make_double2(fma(k.x, cy, u.x), fma(k.y, cy, u.y));
    } else {
        const double2 u = ubase[g], s = acc[g];
        ydst[g] = make_double2(fma(s.x + k.x, dt6, u.x), fma(s.y + k.y, dt6, u.y));
    }
}

__global__ void reservoir_kernel(size_t npts, RhsCoeffs c, const double *__restrict__ pumping,
                                 const double *__restrict__ u_sqr, double *__restrict__ r)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < npts; t += stride)
        // nls.f90:580 / :838, each operation rounded separately as in the reference (no contraction)
        r[t] = __ddiv_rn(__dmul_rn(c.c12, pumping[t]), __dadd_rn(c.c13, __dmul_rn(c.c14, u_sqr[t])));
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

template <int K>
int launch_stage_k(StageMode mode, const CrossWeights &w, const Stage2DArgs &a, cudaStream_t stream)
{
    const dim3 block(64, 4);
    const dim3 grid((a.cols + block.x - 1) / block.x, (a.rows + block.y - 1) / block.y, a.batch);
    const WeightsK<K> p = pack<K>(w);
#define NLSB_LAUNCH_STAGE(MODE)                                                                         \
    stage_2d_kernel<K, MODE><<<grid, block, 0, stream>>>(a.rows, a.cols, p, a.ysrc, a.ubase, a.pumping, \
                                                         a.coeffs, a.acc, a.ydst, a.cy, a.dt6)
    switch (mode) {
    case kStageRhs: NLSB_LAUNCH_STAGE(kStageRhs); break;
    case kStageFirst: NLSB_LAUNCH_STAGE(kStageFirst); break;
    case kStageMid: NLSB_LAUNCH_STAGE(kStageMid); break;
    case kStageLast: NLSB_LAUNCH_STAGE(kStageLast); break;
    }
#undef NLSB_LAUNCH_STAGE
    count_launches(1);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_stage_2d(int order, StageMode mode, const CrossWeights &w, const Stage2DArgs &a, cudaStream_t stream)
{
    if (a.batch > 65535) return fail(NLSB_ESIZE, "batch = %d exceeds the grid z-limit 65535", a.batch);
    switch (order) {
    case 3: return launch_stage_k<1>(mode, w, a, stream);
    case 5: return launch_stage_k<2>(mode, w, a, stream);
    case 7: return launch_stage_k<3>(mode, w, a, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

int launch_general_matvec_2d(int n, int m, const double *blocks_dev, const double *x, double *y, double sign,
                             cudaStream_t stream)
{
    const dim3 block(64, 4), grid((n + block.x - 1) / block.x, (n + block.y - 1) / block.y);
    general_matvec_2d_kernel<<<grid, block, 0, stream>>>(GeneralOp{blocks_dev, m, n}, x, y, sign);
    count_launches(1);
    return (int)cudaGetLastError();
}

int launch_stage_2d_general(int n, int m, const double *blocks_dev, StageMode mode, const Stage2DArgs &a, cudaStream_t stream)
{
    if (a.batch != 1 || a.rows != n || a.cols != n) return fail(NLSB_EINVAL, "general operator: one n x n grid");
    const dim3 block(64, 4), grid((n + block.x - 1) / block.x, (n + block.y - 1) / block.y);
    const GeneralOp op{blocks_dev, m, n};
#define NLSB_LAUNCH_GENERAL(MODE) \
    stage_2d_general_kernel<MODE><<<grid, block, 0, stream>>>(op, a.ysrc, a.ubase, a.pumping, a.coeffs, a.acc, a.ydst, a.cy, a.dt6)
    switch (mode) {
    case kStageRhs: NLSB_LAUNCH_GENERAL(kStageRhs); break;
    case kStageFirst: NLSB_LAUNCH_GENERAL(kStageFirst); break;
    case kStageMid: NLSB_LAUNCH_GENERAL(kStageMid); break;
    case kStageLast: NLSB_LAUNCH_GENERAL(kStageLast); break;
    }
#undef NLSB_LAUNCH_GENERAL
    count_launches(1);
    return (int)cudaGetLastError();
}

int launch_cross_matvec_2d(int rows, int cols, int order, const CrossWeights &w, const double *x, double *y,
                           double sign, cudaStream_t stream)
{
    const dim3 block(64, 4);
    const dim3 grid((cols + block.x - 1) / block.x, (rows + block.y - 1) / block.y);
    switch (order) {
    case 3: cross_matvec_2d_kernel<1><<<grid, block, 0, stream>>>(rows, cols, pack<1>(w), x, y, sign); break;
    case 5: cross_matvec_2d_kernel<2><<<grid, block, 0, stream>>>(rows, cols, pack<2>(w), x, y, sign); break;
    case 7: cross_matvec_2d_kernel<3><<<grid, block, 0, stream>>>(rows, cols, pack<3>(w), x, y, sign); break;
    default: return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    }
    count_launches(1);
    return (int)cudaGetLastError();
}

// Self-test of the divide every fused kernel uses: quotients by div_fast (device_math.cuh) and by the correctly
// rounded IEEE divide, element by element, for the parity tests.
__global__ void divide_check_kernel(size_t n, const double *__restrict__ a, const double *__restrict__ b,
                                    double *__restrict__ fast, double *__restrict__ exact)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fast[i] = div_fast(a[i], b[i]);
        exact[i] = __ddiv_rn(a[i], b[i]);
    }
}

int launch_divide_check(size_t n, const double *a, const double *b, double *fast, double *exact, cudaStream_t stream)
{
    if (n == 0) return 0;
    divide_check_kernel<<<148 * 8, 256, 0, stream>>>(n, a, b, fast, exact);
    count_launches(1);
    return (int)cudaGetLastError();
}

int launch_reservoir(size_t npts, RhsCoeffs c, const double *pumping, const double *u_sqr, double *r,
                     cudaStream_t stream)
{
    if (npts == 0) return 0;
    const int block = 256;
    size_t blocks = (npts + block - 1) / block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    reservoir_kernel<<<(unsigned)blocks, block, 0, stream>>>(npts, c, pumping, u_sqr, r);
    count_launches(1);
    return (int)cudaGetLastError();
}

}  // namespace nlsb
