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
#include "peer_flags.cuh"

#include <cuda.h>
#include <atomic>
#include <type_traits>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace nlsb {

namespace {

using namespace stream2d;

// Tuning knobs (nlsb_set_stream_tuning; initial values from NLSB_STREAM_SYNC / NLSB_STREAM_T / NLSB_STREAM_ITERS):
// -1 / 0 = the library's own choice.
int env_int(const char *name, int fallback)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : fallback;
}
std::atomic<int> g_tune_sync{env_int("NLSB_STREAM_SYNC", -1)};
std::atomic<int> g_tune_width{env_int("NLSB_STREAM_T", 0)};
std::atomic<int> g_tune_iters{env_int("NLSB_STREAM_ITERS", 0)};

struct StreamArgs {
    int rows, cols;          // extent of the local arrays
    int grow0, grows;        // global row of local row 0, global number of rows
    int out_row0, out_row1;  // local rows [out_row0, out_row1) are written
    int strips, chunk_rows;
    double2 *out;            // [batch][rows][cols]
    const double *coeffs;    // [batch][23] -- unused when UNIFORM
    RhsCoeffs cu;
    double dt;
    DiagAcc *diag;           // DIAG: [batch][CTAs per member] partial sums of the entering state's diagnostics
    double area;             // DIAG: area element dx^2
    PeerStep peer;           // PEER: the halo exchange performed by this launch (kernels.h)
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

// How often the threads of a CTA meet.  A stage-ring row written in iteration `it` is read in it + K and overwritten
// in it + 2K + 1, so ONE __syncthreads per min(K, 2) iterations is enough: the barrier after every second iteration
// still separates each write from its read K >= 2 iterations later and each read from the overwrite K + 1 later.
//   kSyncEvery  a barrier per iteration (round 1)
//   kSyncPair   a barrier per two iterations when K >= 2
// Measured on B200 (profiles/r2_stream_sync_sweep.jsonl, order 5): the pair flavour is 1-3 % faster with two
// 128-thread CTAs per SM and 1 % slower with one 256-thread CTA.  Two further flavours were measured and removed:
// split-phase mbarriers (warps arrive after iteration it, wait for the arrivals of it - K: 6-19 % SLOWER -- the
// TRYWAIT round trip per iteration costs more than the drift buys) and one barrier between stage 1 and stages
// 2-4 (9-14 % slower: the stage 2-4 loads can then no longer be hoisted above the stage-1 arithmetic).
enum StreamSync { kSyncEvery = 0, kSyncPair = 1 };

// PEER (the last step of a slab's cycle, multi-GPU): this launch IS the halo exchange.
//   start   CTA 0 publishes READY(e) in both neighbours' flag blocks (my earlier steps, which read my halo rows, are
//           done: stream order); CTAs whose rows include boundary rows wait for the neighbours' READY(e);
//   march   stage 4 stores the boundary rows a second time, straight into the neighbours' halo rows (peer-mapped);
//   end     every CTA fences its stores at system scope and takes a ticket; the last one publishes DATA(e) in the
//           neighbours' flag blocks, waits for theirs and advances the epoch: when the launch completes, this rank's
//           halo rows hold the neighbours' new boundary rows.  Waits time out (peer_flags.cuh).
template <typename C, bool UNIFORM, int SYNC, bool DIAG, bool PEER>
__global__ void __launch_bounds__(C::T, C::CTAS_PER_SM)
rk4_stream_kernel(const __grid_constant__ StreamArgs a, const __grid_constant__ StreamWeights<C::K> wa,
                  const __grid_constant__ TensorMap map, const __grid_constant__ TensorMap map_p)
{
    constexpr int K = C::K, U = C::U, RB = C::RB, NB = C::NB;
    constexpr int PERIOD = (SYNC == kSyncPair && K >= 2) ? 2 : 1;     // iterations per __syncthreads (U is even)
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
    Lane<C> L = make_lane<C>(g, tid, ring, yr, a.out + member * plane, a.rows, a.cols, a.grow0, a.grows, a.dt);
    unsigned long long epoch = 0;
    if (PEER) {
        L.up_delta = a.peer.up_delta; L.dn_delta = a.peer.dn_delta;
        L.up0 = a.peer.up_row0; L.nup = a.peer.flags_up ? a.peer.nrows : 0;
        L.dn0 = a.peer.dn_row0; L.ndn = a.peer.flags_down ? a.peer.nrows : 0;
        if (tid == 0) {
            epoch = a.peer.state[0] + 1;       // the same for every CTA: state[0] changes after the last ticket
            if (blockIdx.x == 0 && blockIdx.y == 0) {
                if (a.peer.flags_up) st_release_sys(a.peer.flags_up + kReadyFromDown, epoch);
                if (a.peer.flags_down) st_release_sys(a.peer.flags_down + kReadyFromUp, epoch);
            }
        }
    }
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
    if (PEER && tid == 0) {
        // rows this CTA writes into a neighbour's halo: that neighbour must have finished reading them
        bool ok = true;
        if (L.nup > 0 && g.r0 < L.up0 + L.nup && g.r1 > L.up0)
            ok = wait_epoch(a.peer.flags_mine + kReadyFromUp, epoch, a.peer.timeout_cycles) && ok;
        if (L.ndn > 0 && g.r0 < L.dn0 + L.ndn && g.r1 > L.dn0)
            ok = wait_epoch(a.peer.flags_mine + kReadyFromDown, epoch, a.peer.timeout_cycles) && ok;
        if (!ok) atomicAdd(a.peer.state + 2, 1ull);
    }
    __syncthreads();
    // warps whose frame columns all lie to the right of the domain have nothing to compute (active_threads): they
    // leave here, after their share of the set-up (the later barriers wait for the non-exited threads of the CTA
    // only).  The diagnostics flavour keeps them for its CTA-wide reduction.
    if (!DIAG && tid >= active_threads<C>(g, a.cols)) return;

    State<C> s;
    DiagAcc dacc = diag_zero();
    mbar_wait(bar0, 0);
    march_begin<C>(s, L, g);

    // U = 2 RB: the batch boundaries fall on fixed phases of the unrolled body
    const double2 *rh = ring + tid, *ro = ring + U * C::T + tid;
    const double *ph_ = pring + tid, *po_ = pring + U * C::T + tid;
    // One unrolled block of U iterations, with or without the domain masks (see march_iter): a block needs them when
    // one of the rows j - 3K .. j of its iterations lies outside the domain -- the first / last blocks of the chunks
    // at the top and bottom of the domain (columns outside the domain are handled by predicated ring stores in both
    // bodies, so the edge strips run the mask-free body too; measured: no difference on 8192^2, one rule less).
    // Both bodies are compiled; the choice is uniform over the CTA.
    auto block = [&](auto masked_tag, int itb) {
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
        for (int ph = 0; ph < U; ++ph) {
            const int it = itb + ph;
            if ((ph + 2 * K) % RB == 0) {          // psi(j + K) is the first row of a new batch
                const int b = (it + 2 * K) / RB;
                mbar_wait(bar0 + 8 * (b % NB), (b / NB) & 1);
            }
            march_iter<C, DIAG, MASKED, PEER>(s, L, g, c, wa.wx, wa.wy, it, ph, rh, ro, ph_, po_, &dacc, a.area);
            if ((ph + 1) % PERIOD == 0) {
                __syncthreads();
                // thread 0 re-requests the batches whose last reader was one of the iterations this barrier closes:
                // iteration x is the last reader of batch (x + K + 1) / RB - 1 when (x + K + 1) % RB == 0.
                // WAR across proxies (generic reads, then the TMA write) is ordered by the barrier above
#pragma unroll
                for (int back = PERIOD - 1; back >= 0; --back) {
                    if ((ph - back + K + 1) % RB == 0 && tid == 0) {
                        const int nb = (it - back + K + 1) / RB - 1 + NB;
                        if (nb < g.nbatches) issue(nb);
                    }
                }
            }
        }
    };
    const int row_lo = L.dlo, row_hi = L.dlo + L.dspan;
    for (int itb = 0; itb < g.niter; itb += U) {
        const int j0 = g.jstart + itb;                                             // stage-1 row of the block's first iteration
        if (j0 - C::SKEW < row_lo || j0 + U > row_hi)
            block(std::true_type{}, itb);
        else
            block(std::false_type{}, itb);
        const double2 *t = rh; rh = ro; ro = t;
        const double *tp = ph_; ph_ = po_; po_ = tp;
    }
    if (DIAG) {
        // one partial per CTA, reduced in a fixed order (the stage rings are free now: the march is over)
        __syncthreads();
        const DiagAcc total = diag_block_reduce(dacc, reinterpret_cast<DiagAcc *>(smem_raw));
        if (tid == 0) a.diag[member * gridDim.x + blockIdx.x] = total;
    }
    if (PEER) {
        __threadfence_system();               // this thread's peer stores are visible system-wide ...
        __syncthreads();                      // ... before the CTA's ticket is taken
        if (tid == 0) {
            const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y;
            if (atomicAdd(a.peer.state + 1, 1ull) == total - 1) {
                __threadfence_system();
                if (a.peer.flags_up) st_release_sys(a.peer.flags_up + kDataFromDown, epoch);
                if (a.peer.flags_down) st_release_sys(a.peer.flags_down + kDataFromUp, epoch);
                bool ok = true;
                if (a.peer.flags_up) ok = wait_epoch(a.peer.flags_mine + kDataFromUp, epoch, a.peer.timeout_cycles) && ok;
                if (a.peer.flags_down) ok = wait_epoch(a.peer.flags_mine + kDataFromDown, epoch, a.peer.timeout_cycles) && ok;
                if (!ok) atomicAdd(a.peer.state + 2, 1ull);
                a.peer.state[1] = 0;
                __threadfence();
                a.peer.state[0] = epoch;
            }
        }
    }
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// complex = true: interleaved complex128 [batch][rows][cols] as a (2, cols, rows, batch) tensor of doubles;
// complex = false: float64 [batch][rows][pitch] as (cols, rows, batch) (pitch must be even: 16-byte row stride)
int encode_field_map(TensorMap *out, const void *base, bool complex, int batch, int rows, int cols, int pitch, int box_cols,
                     int box_rows)
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
        const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(double), (cuuint64_t)pitch * rows * sizeof(double)};
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
    int batch, rows, cols, pitch, box_cols, box_rows;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && complex == o.complex && batch == o.batch && rows == o.rows && cols == o.cols && pitch == o.pitch &&
               box_cols == o.box_cols && box_rows == o.box_rows;
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
    int rc = encode_field_map(out, key.base, key.complex, key.batch, key.rows, key.cols, key.pitch, key.box_cols, key.box_rows);
    if (rc) return rc;
    keys[next] = key;
    maps[next] = *out;
    next = (next + 1) % N;
    return 0;
}

template <typename C, bool UNIFORM, int SYNC, bool DIAG, bool PEER>
int configure_stream()
{
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(rk4_stream_kernel<C, UNIFORM, SYNC, DIAG, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(rk4_stream_kernel<C, UNIFORM, SYNC, DIAG, PEER>, cudaFuncAttributePreferredSharedMemoryCarveout,
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

// Launch geometry: strip width T (threads per CTA), strips, rows per chunk.
// * A strip of T threads produces W = T - 8K columns, so `cols` columns sweep ceil(cols / W) * T frame columns.
// * Every CTA of m U iterations spends 6K of them filling and draining the stage pipeline (plus a fixed prologue
//   worth about 8), and CTAs run in waves of (SMs x CTAs per SM).
// The planner minimises   waves x (iterations per CTA + 8) x (time of one iteration of a full SM)   over the
// compiled widths and the chunk lengths.  The time of an iteration of a full SM is T x CTAs-per-SM x shape_cost(T):
// two 128-thread CTAs per SM overlap each other's load bursts and barriers and need less time per thread than one
// 256-thread CTA -- 8.5 % less in 30-step bursts at 1965 MHz, 3-5 % less in second-long runs, where this kernel sits
// at the 1000 W power cap (1.75-1.80 GHz) and every redundantly swept column costs energy (measured on B200, order
// 5, 8192^2 and 32..64 x 1024^2: profiles/r2_stream_sync_sweep.jsonl, r2_stream_sustained.jsonl; the sustained figure
// is the one used).  160 / 192 / 224-thread strips were measured too and dropped: a CTA whose warps do not divide
// evenly over the four schedulers runs at the pace of the fullest one -- 224 threads take as long per iteration as 256.
struct StreamPlan {
    int threads, strips, chunk_rows, sync;
    double cost;
};

inline double shape_cost(int K, int T)
{
    if (K == 2 && T == 128) return 0.96;
    if (K == 1 && T == 128) return 1.03;
    return 1.0;
}

template <typename C>
StreamPlan plan_shape(int out_rows, int cols, int batch)
{
    StreamPlan p{C::T, (cols + C::W - 1) / C::W, C::chunk_rows(140), kSyncEvery, 0.0};
    p.sync = (C::T <= 128 && C::K >= 2) ? kSyncPair : kSyncEvery;
    // the last strip keeps only the warps whose columns lie inside the domain (active_threads): an SM's time goes with
    // the warps it runs, so the swept columns per row of strips are (strips - 1) T + the last strip's active threads
    const int last_need = cols - (p.strips - 1) * C::W + C::HALO;
    const int last_active = last_need < C::T ? (last_need + 31) / 32 * 32 : C::T;
    const double swept = ((double)(p.strips - 1) * C::T + last_active) / ((double)p.strips * C::T);
    const double per_iter = (double)C::T * C::CTAS_PER_SM * shape_cost(C::K, C::T) * swept;
    const long long slots = (long long)sm_count() * C::CTAS_PER_SM;      // CTAs resident at once
    const int forced = g_tune_iters.load();                              // tuning knob: iterations per CTA
    const int m_lo = forced >= 6 * C::K + C::U ? (forced + C::U - 1) / C::U : (6 * C::K) / C::U + 2;
    // up to ONE chunk per strip: on 8192^2 one wave of 74 x 4 CTAs marching 2048 rows each beats three waves of 683-row
    // chunks by 1-1.5 % (measured) -- the 6K fill / drain iterations and the set-up are paid once instead of three times
    const int m_whole = (out_rows + 6 * C::K + C::U - 1) / C::U + 1;
    const int m_hi = forced >= 6 * C::K + C::U ? m_lo : (m_whole > 96 ? m_whole : 96);
    for (int m = m_lo; m <= m_hi; ++m) {
        const int h = m * C::U - 6 * C::K;
        const long long ctas = (long long)p.strips * ((out_rows + h - 1) / h) * batch;
        const long long waves = (ctas + slots - 1) / slots;
        const int last = out_rows - (out_rows - 1) / h * h;            // rows of the last chunk
        const int iters = out_rows > h ? m * C::U : (last + 6 * C::K + C::U - 1) / C::U * C::U;
        const double cost = (double)waves * (iters + 8) * per_iter;
        if (p.cost == 0.0 || cost <= p.cost) {
            p.cost = cost;
            p.chunk_rows = h;
        }
        if (out_rows <= h) break;
    }
    return p;
}

template <typename C, bool UNIFORM, int SYNC, bool DIAG, bool PEER = false>
int launch_stream_sync(const Fused2DStep &s, const CrossWeights &w, const StreamPlan &p, cudaStream_t stream)
{
    int rc = configure_stream<C, UNIFORM, SYNC, DIAG, PEER>();
    if (rc) return rc;
    const int out_rows = s.out_row1 - s.out_row0;
    StreamArgs a{};
    a.rows = s.rows; a.cols = s.cols; a.grow0 = s.grow0; a.grows = s.grows;
    a.out_row0 = s.out_row0; a.out_row1 = s.out_row1;
    a.strips = p.strips;
    a.chunk_rows = p.chunk_rows;
    a.out = s.out; a.coeffs = s.coeffs;
    if (UNIFORM) a.cu = *s.uniform;
    a.dt = s.dt;
    a.diag = static_cast<DiagAcc *>(s.diag_partial);
    a.area = s.diag_area;
    if (PEER) a.peer = *s.peer;
    StreamWeights<C::K> wa;
    for (int i = 0; i < C::NW; ++i) { wa.wx[i] = w.wx[i]; wa.wy[i] = w.wy[i]; }
    TensorMap map, map_p;
    rc = cached_map(MapKey{s.in, true, s.batch, s.rows, s.cols, s.cols, C::T, C::RB}, &map);
    if (rc) return rc;
    rc = cached_map(MapKey{s.pumping, false, s.batch, s.rows, s.cols, s.p_pitch ? s.p_pitch : s.cols, C::T, C::RB}, &map_p);
    if (rc) return rc;
    const int chunks = (out_rows + a.chunk_rows - 1) / a.chunk_rows;
    const dim3 grid((unsigned)(a.strips * chunks), (unsigned)s.batch);
    rk4_stream_kernel<C, UNIFORM, SYNC, DIAG, PEER><<<grid, C::T, C::SMEM, stream>>>(a, wa, map, map_p);
    count_launches(1);
    return (int)cudaGetLastError();
}

template <typename C>
int launch_stream_cfg(const Fused2DStep &s, const CrossWeights &w, const StreamPlan &p, cudaStream_t stream)
{
    if (s.peer) {           // the exchange-carrying step of a slab's cycle: shared coefficients
        if (!s.uniform || s.diag_partial || s.batch != 1) return fail(NLSB_EINVAL, "the exchange-carrying step takes one grid with shared coefficients");
        // same barrier cadence as the plain steps of the cycle
        if (p.sync == kSyncPair) return launch_stream_sync<C, true, kSyncPair, false, true>(s, w, p, stream);
        return launch_stream_sync<C, true, kSyncEvery, false, true>(s, w, p, stream);
    }
    if (s.diag_partial)     // the diagnostics-carrying step of a chunk (one launch in many): a single flavour
        return s.uniform ? launch_stream_sync<C, true, kSyncEvery, true>(s, w, p, stream)
                         : launch_stream_sync<C, false, kSyncEvery, true>(s, w, p, stream);
    if (p.sync == kSyncPair)
        return s.uniform ? launch_stream_sync<C, true, kSyncPair, false>(s, w, p, stream)
                         : launch_stream_sync<C, false, kSyncPair, false>(s, w, p, stream);
    return s.uniform ? launch_stream_sync<C, true, kSyncEvery, false>(s, w, p, stream)
                     : launch_stream_sync<C, false, kSyncEvery, false>(s, w, p, stream);
}

// Strip widths the march is compiled for (order 7 has one: rings of 256 columns exceed 227 KB of shared memory and
// two 128-column CTAs do not fit one SM either).
template <int K> struct StreamWidths;
template <> struct StreamWidths<1> { static constexpr int n = 2; static constexpr int t[2] = {128, 256}; };
template <> struct StreamWidths<2> { static constexpr int n = 2; static constexpr int t[2] = {128, 256}; };
template <> struct StreamWidths<3> { static constexpr int n = 1; static constexpr int t[1] = {192}; };

// Best plan over the compiled widths (or the forced one); ties go to the wider strip.
template <int K, int I = 0>
StreamPlan plan_stream(int batch, int out_rows, int cols)
{
    using SW = StreamWidths<K>;
    StreamPlan p = plan_shape<Cfg<K, SW::t[I]>>(out_rows, cols, batch);
    const int force_t = g_tune_width.load(), force_sync = g_tune_sync.load();
    if (force_sync == kSyncEvery || force_sync == kSyncPair) p.sync = force_sync;
    if constexpr (I + 1 < SW::n) {
        const StreamPlan q = plan_stream<K, I + 1>(batch, out_rows, cols);
        if (force_t == p.threads) return p;
        if (force_t == q.threads || q.cost <= p.cost) return q;
    }
    return p;
}

template <int K, int I = 0>
int launch_stream_k(const Fused2DStep &s, const CrossWeights &w, const StreamPlan &p, cudaStream_t stream)
{
    using SW = StreamWidths<K>;
    if constexpr (I < SW::n) {
        if (p.threads == SW::t[I]) return launch_stream_cfg<Cfg<K, SW::t[I]>>(s, w, p, stream);
        return launch_stream_k<K, I + 1>(s, w, p, stream);
    } else {
        return fail(NLSB_EINVAL, "strip width %d is not compiled for order %d", p.threads, 2 * K + 1);
    }
}

template <int K>
int launch_stream_order(const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    const int out_rows = s.out_row1 - s.out_row0;
    if (out_rows <= 0 || s.cols <= 0 || s.batch <= 0) return 0;
    return launch_stream_k<K>(s, w, plan_stream<K>(s.batch, out_rows, s.cols), stream);
}

}  // namespace

void stream_2d_get_tuning(int *t3)
{
    t3[0] = g_tune_sync.load();
    t3[1] = g_tune_width.load();
    t3[2] = g_tune_iters.load();
}

void stream_2d_set_tuning(int sync, int width, int iters)
{
    g_tune_sync.store(sync);
    g_tune_width.store(width);
    g_tune_iters.store(iters);
}

// The launch geometry launch_rk4_step_stream_2d would use (host arithmetic only; needs no device).
int stream_2d_plan(int order, int batch, int out_rows, int cols, int *threads, int *strips, int *chunk_rows)
{
    StreamPlan p;
    switch (order) {
    case 3: p = plan_stream<1>(batch, out_rows, cols); break;
    case 5: p = plan_stream<2>(batch, out_rows, cols); break;
    case 7: p = plan_stream<3>(batch, out_rows, cols); break;
    default: return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    }
    *threads = p.threads; *strips = p.strips; *chunk_rows = p.chunk_rows;
    return 0;
}

// Whether launch_rk4_step_stream_2d runs the strip-marching kernel for this step (else it hands over to the tile kernel).
bool stream_2d_takes(const Fused2DStep &s)
{
    const int pitch = s.p_pitch ? s.p_pitch : s.cols;
    return (pitch & 1) == 0 && (reinterpret_cast<uintptr_t>(s.pumping) & 15) == 0;
}

// CTAs per member of the launch = partial sums per member a diagnostics-carrying step writes.
int stream_2d_diag_parts(int order, int batch, int out_rows, int cols)
{
    int threads = 0, strips = 0, chunk_rows = 0;
    if (stream_2d_plan(order, batch, out_rows, cols, &threads, &strips, &chunk_rows) || chunk_rows <= 0) return 0;
    return strips * ((out_rows + chunk_rows - 1) / chunk_rows);
}

int launch_rk4_step_stream_2d(int order, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    if (s.batch > 65535) return fail(NLSB_ESIZE, "batch = %d exceeds the grid y-limit 65535", s.batch);
    if ((reinterpret_cast<uintptr_t>(s.in) & 15) != 0) return fail(NLSB_EINVAL, "psi must be 16-byte aligned");
    // the pumping's tensor map needs a 16-byte row stride and base: callers that cannot provide them (one-off slab
    // steps on grids with an odd number of columns) get the tile kernel -- same bits
    if (!stream_2d_takes(s)) {
        if (s.peer) return fail(NLSB_EINVAL, "the exchange-carrying step needs the strip-marching kernel (even column count, aligned pumping)");
        if (s.p_pitch && s.p_pitch != s.cols) return fail(NLSB_EINVAL, "a pitched pumping copy must have an even pitch and a 16-byte aligned base");
        return launch_rk4_step_fused_2d(order, 0, s, w, stream);
    }
    switch (order) {
    case 3: return launch_stream_order<1>(s, w, stream);
    case 5: return launch_stream_order<2>(s, w, stream);
    case 7: return launch_stream_order<3>(s, w, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

}  // namespace nlsb
