// stream_2d.cu -- streaming fused RK4 step of the 2D solver for sm_100a: one launch = one RK4 step
// (runge_kutta_2d, nls.f90:892-899), the field read ONCE and written ONCE (40 B per node-step).
//
// A CTA owns a strip of columns and marches down a chunk of rows with the four RK stages skewed by K rows
// (stream_2d_core.cuh).  psi rows arrive through a ring of TMA batches (cp.async.bulk.tensor over the
// interleaved complex128 array viewed as a (2, cols, rows, batch) tensor of doubles; rows / columns outside
// the array arrive as zeros = the reference's truncated band matrix) and so do the rows of the pumping (a second
// tensor map; its row stride must be a multiple of 16 bytes, so grids with an odd number of columns take the tile
// kernel); the new psi is stored with one coalesced 16-byte store per thread and row.
// One __syncthreads per iteration: a stage ring row is written in iteration `it` and read in it + K.
//
// Compared with the tile kernel (fused_2d.cu): y neighbours never touch shared memory, redundant work falls
// from 1.43 to about 1.12, no planar working copy, no split / join passes.  Needs grids large enough to fill
// the GPU with (strip, chunk) CTAs; small grids stay on the tile kernel (api.cu chooses).

#include "kernels.h"
#include "stream_2d_core.cuh"

#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace nlsb {

namespace {

using namespace stream2d;

struct StreamArgs {
    int rows, cols;          // extent of the local arrays
    int grow0, grows;        // global row of local row 0, global number of rows
    int out_row0, out_row1;  // local rows [out_row0, out_row1) are written
    int strips, chunk_rows;
    double2 *out;            // [batch][rows][cols]
    const double *coeffs;    // [batch][23] -- unused when UNIFORM
    RhsCoeffs cu;
    double dt;
};

template <int K>
struct StreamWeights {
    double wx[2 * K + 1];
    double wy[2 * K + 1];
};

struct TensorMap {
    alignas(64) unsigned char bytes[128];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <typename C, bool UNIFORM>
__global__ void __launch_bounds__(C::T, C::CTAS_PER_SM)
rk4_stream_kernel(const __grid_constant__ StreamArgs a, const __grid_constant__ StreamWeights<C::K> wa,
                  const __grid_constant__ TensorMap map, const __grid_constant__ TensorMap map_p)
{
    constexpr int K = C::K, U = C::U, RB = C::RB, NB = C::NB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *yr = reinterpret_cast<double2 *>(smem_raw);
    double2 *ring = reinterpret_cast<double2 *>(smem_raw + C::RING_OFFSET);
    double *pring = reinterpret_cast<double *>(smem_raw + C::PRING_OFFSET);
    const uint32_t bar0 = smem_u32(smem_raw + C::BAR_OFFSET);

    const int tid = threadIdx.x;
    const int strip = blockIdx.x % a.strips, chunk = blockIdx.x / a.strips;
    const size_t member = blockIdx.y;
    const size_t plane = (size_t)a.rows * a.cols;
    const Chunk g = make_chunk<C>(strip, chunk, a.chunk_rows, a.out_row0, a.out_row1);
    const Lane<C> L = make_lane<C>(g, tid, ring, yr, a.out + member * plane, a.rows, a.cols, a.grow0, a.grows, a.dt);
    RhsCoeffs cl;
    if (!UNIFORM) cl = load_rhs_coeffs(a.coeffs + member * 23);
    const RhsCoeffs &c = UNIFORM ? a.cu : cl;

    constexpr uint32_t kBatchBytes = sizeof(double2) * RB * C::T, kBatchBytesP = sizeof(double) * RB * C::T;
    auto issue = [&](int b) {     // thread 0: request TMA batch b (rows base + b RB ...) of psi and of the pumping
        const uint32_t bar = bar0 + 8 * (b % NB);
        mbar_expect_tx(bar, kBatchBytes + kBatchBytesP);
        tma_load_4d(smem_u32(ring) + (b % NB) * kBatchBytes, &map, bar, 0, g.c0 - C::HALO, g.base + b * RB, (int)member);
        tma_load_3d(smem_u32(pring) + (b % NB) * kBatchBytesP, &map_p, bar, g.c0 - C::HALO, g.base + b * RB, (int)member);
    };

    if (tid == 0) {
        for (int b = 0; b < NB; ++b) mbar_init(bar0 + 8 * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int b = 0; b < NB && b < g.nbatches; ++b) issue(b);
    }
    for (int i = tid; i < 3 * C::YS * C::YP; i += C::T) yr[i] = make_double2(0.0, 0.0);
    __syncthreads();

    State<C> s;
    mbar_wait(bar0, 0);
    march_begin<C>(s, L, g);

    // U = 2 RB: the batch boundaries fall on fixed phases of the unrolled body
    const double2 *rh = ring + tid, *ro = ring + U * C::T + tid;
    const double *ph_ = pring + tid, *po_ = pring + U * C::T + tid;
    for (int itb = 0; itb < g.niter; itb += U) {
#pragma unroll
        for (int ph = 0; ph < U; ++ph) {
            const int it = itb + ph;
            if ((ph + 2 * K) % RB == 0) {          // psi(j + K) is the first row of a new batch
                const int b = (it + 2 * K) / RB;
                mbar_wait(bar0 + 8 * (b % NB), (b / NB) & 1);
            }
            march_iter<C>(s, L, g, c, wa.wx, wa.wy, it, ph, rh, ro, ph_, po_);
            __syncthreads();
            if ((ph + K + 1) % RB == 0 && tid == 0) {   // every row of batch (it + K + 1) / RB - 1 has been consumed
                const int nb = (it + K + 1) / RB - 1 + NB;
                // WAR across proxies (generic reads, then the TMA write) is ordered by the barrier above
                if (nb < g.nbatches) issue(nb);
            }
        }
        const double2 *t = rh; rh = ro; ro = t;
        const double *tp = ph_; ph_ = po_; po_ = tp;
    }
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// complex = true: interleaved complex128 [batch][rows][cols] as a (2, cols, rows, batch) tensor of doubles;
// complex = false: float64 [batch][rows][cols] as (cols, rows, batch) (cols must be even: 16-byte row stride)
int encode_field_map(TensorMap *out, const void *base, bool complex, int batch, int rows, int cols, int box_cols, int box_rows)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return (int)e;
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(NLSB_EINVAL, "cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    CUtensorMap map;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (complex) {
        const cuuint64_t dims[4] = {2, (cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
        const cuuint64_t strides[3] = {sizeof(double2), (cuuint64_t)cols * sizeof(double2),
                                       (cuuint64_t)cols * rows * sizeof(double2)};
        const cuuint32_t box[4] = {2, (cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
        r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)cols * sizeof(double), (cuuint64_t)cols * rows * sizeof(double)};
        const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
        r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(NLSB_EINVAL, "cuTensorMapEncodeTiled (stream kernel) failed with CUresult %d", (int)r);
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    std::memcpy(out->bytes, &map, 128);
    return 0;
}

// The time loop ping-pongs between two buffers: keep the last few encoded maps instead of re-encoding per step.
struct MapKey {
    const void *base;
    bool complex;
    int batch, rows, cols, box_cols, box_rows;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && complex == o.complex && batch == o.batch && rows == o.rows && cols == o.cols && box_cols == o.box_cols &&
               box_rows == o.box_rows;
    }
};

int cached_map(const MapKey &key, TensorMap *out)
{
    constexpr int N = 8;
    static thread_local MapKey keys[N] = {};
    static thread_local TensorMap maps[N];
    static thread_local int next = 0;
    for (int i = 0; i < N; ++i)
        if (keys[i].base && keys[i] == key) {
            *out = maps[i];
            return 0;
        }
    int rc = encode_field_map(out, key.base, key.complex, key.batch, key.rows, key.cols, key.box_cols, key.box_rows);
    if (rc) return rc;
    keys[next] = key;
    maps[next] = *out;
    next = (next + 1) % N;
    return 0;
}

template <typename C, bool UNIFORM>
int configure_stream()
{
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(rk4_stream_kernel<C, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(rk4_stream_kernel<C, UNIFORM>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    return 0;
}

int sm_count()
{
    static int sms[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) {
        int n = 0;
        sms[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return sms[dev];
}

// Rows per chunk.  One CTA per SM runs at a time, every CTA of m U iterations spends 6K of them filling and
// draining the stage pipeline (plus a fixed prologue worth about 8): pick the m that minimises
// waves x (iterations per CTA), i.e. long chunks, but not so long that the last wave is mostly empty.
template <typename C>
int chunk_rows_for(int out_rows, int strips, int batch)
{
    static const int forced = [] {
        const char *e = std::getenv("NLSB_STREAM_ITERS");     // tuning knob: iterations per CTA
        return e ? std::atoi(e) : 0;
    }();
    if (forced >= 6 * C::K + C::U) return C::chunk_rows(forced);
    const long long sms = (long long)sm_count() * C::CTAS_PER_SM;      // CTAs resident at once
    int best_h = C::chunk_rows(140);
    double best_cost = 0.0;
    for (int m = (6 * C::K) / C::U + 2; m <= 96; ++m) {
        const int h = m * C::U - 6 * C::K;
        const long long ctas = (long long)strips * ((out_rows + h - 1) / h) * batch;
        const long long waves = (ctas + sms - 1) / sms;
        const int last = out_rows - (out_rows - 1) / h * h;            // rows of the last chunk
        const int iters = out_rows > h ? m * C::U : (last + 6 * C::K + C::U - 1) / C::U * C::U;
        const double cost = (double)waves * (iters + 8);
        if (best_cost == 0.0 || cost <= best_cost) {
            best_cost = cost;
            best_h = h;
        }
        if (out_rows <= h) break;
    }
    return best_h;
}

template <typename C, bool UNIFORM>
int launch_stream_cfg(const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    int rc = configure_stream<C, UNIFORM>();
    if (rc) return rc;
    const int out_rows = s.out_row1 - s.out_row0;
    if (out_rows <= 0 || s.cols <= 0 || s.batch <= 0) return 0;
    StreamArgs a{};
    a.rows = s.rows; a.cols = s.cols; a.grow0 = s.grow0; a.grows = s.grows;
    a.out_row0 = s.out_row0; a.out_row1 = s.out_row1;
    a.strips = (s.cols + C::W - 1) / C::W;
    a.chunk_rows = chunk_rows_for<C>(out_rows, a.strips, s.batch);
    a.out = s.out; a.coeffs = s.coeffs;
    if (UNIFORM) a.cu = *s.uniform;
    a.dt = s.dt;
    StreamWeights<C::K> wa;
    for (int i = 0; i < C::NW; ++i) { wa.wx[i] = w.wx[i]; wa.wy[i] = w.wy[i]; }
    TensorMap map, map_p;
    rc = cached_map(MapKey{s.in, true, s.batch, s.rows, s.cols, C::T, C::RB}, &map);
    if (rc) return rc;
    rc = cached_map(MapKey{s.pumping, false, s.batch, s.rows, s.cols, C::T, C::RB}, &map_p);
    if (rc) return rc;
    const int chunks = (out_rows + a.chunk_rows - 1) / a.chunk_rows;
    const dim3 grid((unsigned)(a.strips * chunks), (unsigned)s.batch);
    rk4_stream_kernel<C, UNIFORM><<<grid, C::T, C::SMEM, stream>>>(a, wa, map, map_p);
    count_launches(1);
    return (int)cudaGetLastError();
}

template <int K>
struct StreamShape {
    using Wide = Cfg<K, (K == 3) ? 192 : 256>;      // order 7: the rings of 256 columns do not fit in shared memory
    using Narrow = Cfg<K, 128>;                     // two CTAs per SM: their barriers and load bursts are independent
};

// Two 128-thread CTAs per SM overlap each other's load bursts and barriers (+3 % per unit of work, measured) but
// sweep T / W = 128 / 112 columns per useful column instead of 256 / 240: taken where that costs nothing
// (1024-wide ensemble members: 10 x 128 = 5 x 256 columns).
template <int K>
bool narrow_shape(int cols)
{
    using C = typename StreamShape<K>::Wide;
    using N = typename StreamShape<K>::Narrow;
    static const int force = [] {
        const char *e = std::getenv("NLSB_STREAM_T");         // tuning knob: 128 or 256
        return e ? std::atoi(e) : 0;
    }();
    const long long narrow = (long long)((cols + N::W - 1) / N::W) * N::T, wide = (long long)((cols + C::W - 1) / C::W) * C::T;
    return force == 128 || (force == 0 && K != 3 && narrow <= wide);
}

template <int K>
int launch_stream_k(const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    using C = typename StreamShape<K>::Wide;
    using N = typename StreamShape<K>::Narrow;
    if (narrow_shape<K>(s.cols))
        return s.uniform ? launch_stream_cfg<N, true>(s, w, stream) : launch_stream_cfg<N, false>(s, w, stream);
    return s.uniform ? launch_stream_cfg<C, true>(s, w, stream) : launch_stream_cfg<C, false>(s, w, stream);
}

template <int K>
void plan_k(int batch, int out_rows, int cols, int *threads, int *strips, int *chunk_rows)
{
    using C = typename StreamShape<K>::Wide;
    using N = typename StreamShape<K>::Narrow;
    if (narrow_shape<K>(cols)) {
        *threads = N::T; *strips = (cols + N::W - 1) / N::W; *chunk_rows = chunk_rows_for<N>(out_rows, *strips, batch);
    } else {
        *threads = C::T; *strips = (cols + C::W - 1) / C::W; *chunk_rows = chunk_rows_for<C>(out_rows, *strips, batch);
    }
}

}  // namespace

// The launch geometry launch_rk4_step_stream_2d would use (host arithmetic only; needs no device).
int stream_2d_plan(int order, int batch, int out_rows, int cols, int *threads, int *strips, int *chunk_rows)
{
    switch (order) {
    case 3: plan_k<1>(batch, out_rows, cols, threads, strips, chunk_rows); return 0;
    case 5: plan_k<2>(batch, out_rows, cols, threads, strips, chunk_rows); return 0;
    case 7: plan_k<3>(batch, out_rows, cols, threads, strips, chunk_rows); return 0;
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

int launch_rk4_step_stream_2d(int order, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    if (s.batch > 65535) return fail(NLSB_ESIZE, "batch = %d exceeds the grid y-limit 65535", s.batch);
    if ((reinterpret_cast<uintptr_t>(s.in) & 15) != 0) return fail(NLSB_EINVAL, "psi must be 16-byte aligned");
    // the pumping's tensor map needs a 16-byte row stride and base: other grids take the tile kernel (same bits)
    if ((s.cols & 1) || (reinterpret_cast<uintptr_t>(s.pumping) & 15) != 0) return launch_rk4_step_fused_2d(order, 0, s, w, stream);
    switch (order) {
    case 3: return launch_stream_k<1>(s, w, stream);
    case 5: return launch_stream_k<2>(s, w, stream);
    case 7: return launch_stream_k<3>(s, w, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

}  // namespace nlsb
