// api.cu -- the C ABI of libnls_b200.so (see include/nls_b200.h for the contract and for the
// reference routine each entry point stands in for).

#include "kernels.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

namespace nlsb {

// ---- error channel ------------------------------------------------------------------------------
static thread_local char g_error[512] = "";

// Kernel launches issued by this library since load (launches replayed from a CUDA graph included).
static std::atomic<unsigned long long> g_launches{0};
void count_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

static int cuda_fail(cudaError_t e, const char *what)
{
    return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

#define NLSB_CUDA(expr)                                          \
    do {                                                         \
        cudaError_t e_ = (expr);                                 \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr);      \
    } while (0)

#define NLSB_TRY(expr)                 \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != 0) {                \
            if (rc_ > 0) cuda_fail((cudaError_t)rc_, #expr); \
            return rc_;                \
        }                              \
    } while (0)

int launch_weighted_dots(size_t npts, double radial_dx, const double2 *u0, const double2 *v, void *scratch,
                         double *out4, cudaStream_t stream);
size_t weighted_dots_scratch_bytes();

namespace {

bool valid_order(int m) { return m == 3 || m == 5 || m == 7; }

// The fused kernels divide by c13 + c14 |psi|^2 with a reciprocal-seeded divide whose precondition is a finite,
// normal, positive denominator (device_math.cuh): every model the reference's model.py builds has c13 = 1 and
// c14 = R phi0^2 / gamma_R > 0 (model.py:157-159).  Coefficient sets outside that domain (where the reference's
// own divide would produce +-inf / NaN fields anyway) are refused instead of being silently mis-divided.
int check_coeffs(const double *c23)
{
    for (int i : {2, 3, 4, 5, 11, 12, 13})
        if (!std::isfinite(c23[i])) return fail(NLSB_EINVAL, "coeffs[%d] is not finite", i);
    if (!(c23[12] >= 1e-290) || !(c23[13] >= 0.0))
        return fail(NLSB_EINVAL, "coeffs[12] (= %g) must be positive and coeffs[13] (= %g) non-negative: the reservoir "
                                 "denominator c13 + c14 |psi|^2 must stay positive", c23[12], c23[13]);
    return 0;
}

int check_order_size(int n, int order)
{
    if (!valid_order(order)) return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    if (n < order) return fail(NLSB_ESIZE, "n = %d is smaller than the stencil width %d", n, order);
    return 0;
}

// Internal stream of the host-buffer entry points: one per (thread, device), created lazily.
int internal_stream(cudaStream_t *out)
{
    static thread_local cudaStream_t streams[64] = {};
    int dev = 0;
    NLSB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(NLSB_EINVAL, "device ordinal %d out of range", dev);
    if (!streams[dev]) NLSB_CUDA(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
    *out = streams[dev];
    return 0;
}

// The library's PRIVATE stream-ordered memory pool (one per device): scratch of the host-buffer entry points stays
// cached here between calls (the default pool would hand it back to the driver at every synchronisation, and its
// attributes belong to the embedding application).  nlsb_trim_memory() returns the cached memory to the driver.
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};

int private_pool(cudaMemPool_t *out)
{
    int dev = 0;
    NLSB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(NLSB_EINVAL, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        NLSB_CUDA(cudaMemPoolCreate(&g_pools[dev], &props));
        unsigned long long keep = ~0ull;      // keep freed blocks until nlsb_trim_memory
        NLSB_CUDA(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = g_pools[dev];
    return 0;
}

// Stream-ordered scratch allocations released when the scope ends (also on error paths).
class Arena {
public:
    explicit Arena(cudaStream_t s) : stream_(s) {}
    ~Arena()
    {
        for (void *p : ptrs_) cudaFreeAsync(p, stream_);
    }
    template <typename T>
    int alloc(T **out, size_t count)
    {
        cudaMemPool_t pool;
        NLSB_TRY(private_pool(&pool));
        void *p = nullptr;
        cudaError_t e = cudaMallocFromPoolAsync(&p, sizeof(T) * (count ? count : 1), pool, stream_);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocFromPoolAsync");
        ptrs_.push_back(p);
        *out = static_cast<T *>(p);
        return 0;
    }
    template <typename T>
    int upload(T **out, const T *host, size_t count)
    {
        NLSB_TRY(alloc(out, count));
        NLSB_CUDA(cudaMemcpyAsync(*out, host, sizeof(T) * count, cudaMemcpyHostToDevice, stream_));
        return 0;
    }

private:
    cudaStream_t stream_;
    std::vector<void *> ptrs_;
};

// ---- 2D time loop: 4 stage launches per RK step, replayed from a CUDA graph ----------------------
void enqueue_step_2d(int batch, int rows, int cols, int order, double dt, const CrossWeights &w,
                     const double *pumping, const double *coeffs, double2 *psi, double2 *work, cudaStream_t stream,
                     int *rc)
{
    const size_t np = (size_t)batch * rows * cols;
    double2 *ya = work, *yb = work + np, *acc = work + 2 * np;
    Stage2DArgs a{batch, rows, cols, nullptr, psi, pumping, coeffs, acc, nullptr, 0.0, 1.0, dt / 6};
    auto stage = [&](StageMode mode, const double2 *src, double2 *dst, double cy) {
        a.ysrc = src;
        a.ydst = dst;
        a.cy = cy;
        int r = launch_stage_2d(order, mode, w, a, stream);
        if (r && !*rc) *rc = r;
    };
    stage(kStageFirst, psi, ya, dt / 2);
    stage(kStageMid, ya, yb, dt / 2);
    stage(kStageMid, yb, ya, dt);
    stage(kStageLast, ya, psi, 0.0);
}

// 0 = automatic (see below), 1 = per-stage kernels, 2 / 3 = fused 32x32 / 32x64 tiles filled with plain loads from
// interleaved psi, 4 / 5 = fused 32x32 / 32x64 tiles filled by TMA from the planar working copy, one launch
// per step, 6 / 7 = as 4 / 5 but the whole time loop in one persistent, neighbour-synchronised launch when
// every tile is resident at once (experimental: measured slower than per-step launches),
// 8 = streaming strip-marching kernel (stream_2d.cu) on interleaved psi, one launch per step,
// 9 = resident kernel (resident_2d.cu): the whole time loop in one cooperative launch, field in registers,
//     for grids whose patches are all resident at once (falls back to 4 otherwise).
// Automatic: the streaming kernel for launches of at least 2^20 nodes (measured on B200, order 5: 1.0x the tile
// kernel at 1024^2, 1.5x at 2048^2, 1.56x at 8192^2; 4096^2: 1.75x at order 3, 1.84x at order 7), the TMA tile
// kernel below that -- with 32x64 tiles when those give every SM at most one CTA while 32x32 tiles would not
// (512^2: 7.86 vs 8.21 us per step), else 32x32.  All produce the same bits.
static std::atomic<int> g_path_2d{0};

static bool stream_preferred(int order, int batch, int out_rows, int cols)
{
    (void)order;
    (void)cols;
    return (long long)batch * out_rows * cols >= (1ll << 20);
}

static int launch_interleaved_step(int order, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    const int path = g_path_2d.load();
    if (path == 8 || (path == 0 && stream_preferred(order, s.batch, s.out_row1 - s.out_row0, s.cols)))
        return launch_rk4_step_stream_2d(order, s, w, stream);
    return launch_rk4_step_fused_2d(order, path == 3 ? 1 : 0, s, w, stream);
}

// Coefficients shared by every member of the batch (host copy), or null: set by the entry points that
// know them so that the fused kernel can read them from its constant bank.
static thread_local const RhsCoeffs *g_uniform_coeffs = nullptr;
struct UniformCoeffsScope {
    RhsCoeffs value;
    explicit UniformCoeffsScope(const double *coeffs23_host)
    {
        if (coeffs23_host) {
            value = rhs_coeffs_from(coeffs23_host);
            g_uniform_coeffs = &value;
        }
    }
    ~UniformCoeffsScope() { g_uniform_coeffs = nullptr; }
};

// A request to reduce the scalar diagnostics of the state entering the LAST step of a time loop inside that step's
// launch (strip-marching kernel): partial sums per CTA land in `partial`.
struct DiagRequest {
    void *partial;
    double area;
};

// Fused path: one launch per RK step, psi ping-pongs between `psi` and the first plane of `work`.
void enqueue_fused_steps_2d(int batch, int rows, int cols, int order, double dt, const CrossWeights &w,
                            const double *pumping, const double *coeffs, double2 *psi, double2 *work, int first_step,
                            int nsteps, cudaStream_t stream, int *rc, const DiagRequest *diag_on_last = nullptr, int p_pitch = 0)
{
    Fused2DStep s{batch, rows, cols, 0, rows, 0, rows, nullptr, nullptr, pumping, coeffs, dt, g_uniform_coeffs};
    s.p_pitch = p_pitch;
    for (int i = 0; i < nsteps; ++i) {
        const bool even = ((first_step + i) & 1) == 0;
        s.in = even ? psi : work;
        s.out = even ? work : psi;
        if (diag_on_last && i == nsteps - 1) {
            s.diag_partial = diag_on_last->partial;
            s.diag_area = diag_on_last->area;
        }
        int r = launch_interleaved_step(order, s, w, stream);
        if (r && !*rc) *rc = r;
    }
}

// Instantiated graphs of 32-step chunks of the fused time loop, keyed by everything the captured launches depend
// on (buffers, geometry, coefficients passed by value, kernel choice): repeated advance() / solve calls on the same
// buffers replay the cached executable instead of capturing and instantiating again (milliseconds per call).
struct LoopKey {
    const void *psi, *work, *pumping, *coeffs;
    int batch, rows, cols, order, path, device, has_uniform, tune[3];
    double dt;
    RhsCoeffs uniform;
    CrossWeights w;
};
struct LoopGraph {
    LoopKey key;
    cudaGraphExec_t exec;
    unsigned long long stamp;
};
std::mutex g_graph_mutex;
LoopGraph g_graphs[8] = {};
unsigned long long g_graph_stamp = 0;

LoopKey make_loop_key(const void *psi, const void *work, const void *pumping, const void *coeffs, int batch, int rows,
                      int cols, int order, int path, double dt, const CrossWeights &w, int *rc)
{
    LoopKey key;
    std::memset(&key, 0, sizeof(key));
    key.psi = psi; key.work = work; key.pumping = pumping; key.coeffs = coeffs;
    key.batch = batch; key.rows = rows; key.cols = cols; key.order = order; key.path = path;
    cudaError_t e = cudaGetDevice(&key.device);
    *rc = e == cudaSuccess ? 0 : cuda_fail(e, "cudaGetDevice");
    key.has_uniform = g_uniform_coeffs ? 1 : 0;
    if (g_uniform_coeffs) key.uniform = *g_uniform_coeffs;
    stream_2d_get_tuning(key.tune);
    key.dt = dt;
    key.w = w;
    return key;
}

// The cached executable graph of `key`, recorded by `record(stream, &rc)` (a chunk of step launches) on first use.
// Call with g_graph_mutex held; the returned slot stays valid while the mutex is held.
template <class Record>
int cached_loop_graph(const LoopKey &key, int launches_recorded, Record record, cudaGraphExec_t *exec_out)
{
    LoopGraph *slot = nullptr, *victim = &g_graphs[0];
    for (LoopGraph &g : g_graphs) {
        if (g.exec && std::memcmp(&g.key, &key, sizeof(key)) == 0) slot = &g;
        if (!g.exec || (victim->exec && g.stamp < victim->stamp)) victim = &g;
    }
    if (!slot) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaStream_t rec;
        NLSB_TRY(internal_stream(&rec));
        NLSB_CUDA(cudaStreamBeginCapture(rec, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        record(rec, &rc);
        cudaError_t e = cudaStreamEndCapture(rec, &graph);
        count_launches(0ull - (unsigned long long)launches_recorded);   // recorded, not run
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
        if (victim->exec) cudaGraphExecDestroy(victim->exec);   // released once its in-flight launches complete
        victim->key = key;
        victim->exec = exec;
        slot = victim;
    }
    slot->stamp = ++g_graph_stamp;
    *exec_out = slot->exec;
    return 0;
}

int enqueue_rk4_2d_fused(int batch, int rows, int cols, int order, int iters, double dt, const CrossWeights &w,
                         const double *pumping, const double *coeffs, double2 *psi, double2 *work, cudaStream_t stream,
                         const DiagRequest *diag = nullptr)
{
    int rc = 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NLSB_CUDA(cudaStreamIsCapturing(stream, &cap));
    const int chunk = 32;   // even: a replayed chunk starts and ends in `psi`
    int done = 0;
    const int replayable = diag ? iters - 1 : iters;      // the diagnostics-carrying last step is launched directly
    // The strip-marching kernel fetches the pumping rows with TMA (16-byte row stride and base): a grid with an odd
    // number of columns gets a copy with an even pitch, made once per time loop in the unused part of the work space
    // (the loop ping-pongs between psi and the first psi-sized plane of `work`); 8 + 8 bytes per node, once.
    int p_pitch = 0;
    {
        const int path = g_path_2d.load();
        const bool marching = path == 8 || (path == 0 && stream_preferred(order, batch, rows, cols));
        if (marching && iters > 0 && ((cols & 1) || (reinterpret_cast<uintptr_t>(pumping) & 15) != 0)) {
            p_pitch = (cols + 1) & ~1;
            double *padded = reinterpret_cast<double *>(work + (size_t)batch * rows * cols);
            NLSB_CUDA(cudaMemcpy2DAsync(padded, sizeof(double) * p_pitch, pumping, sizeof(double) * cols, sizeof(double) * cols,
                                        (size_t)batch * rows, cudaMemcpyDeviceToDevice, stream));
            pumping = padded;
        }
    }
    if (cap == cudaStreamCaptureStatusNone && replayable >= 2 * chunk) {
        int krc = 0;
        const LoopKey key = make_loop_key(psi, work, pumping, coeffs, batch, rows, cols, order, g_path_2d.load(), dt, w, &krc);
        if (krc) return krc;
        std::lock_guard<std::mutex> lock(g_graph_mutex);
        cudaGraphExec_t exec = nullptr;
        NLSB_TRY(cached_loop_graph(key, chunk, [&](cudaStream_t rec, int *r) {
            enqueue_fused_steps_2d(batch, rows, cols, order, dt, w, pumping, coeffs, psi, work, 0, chunk, rec, r, nullptr, p_pitch);
        }, &exec));
        for (; done + chunk <= replayable; done += chunk) {
            cudaError_t e = cudaGraphLaunch(exec, stream);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
            count_launches((unsigned long long)chunk);
        }
    }
    enqueue_fused_steps_2d(batch, rows, cols, order, dt, w, pumping, coeffs, psi, work, done, iters - done, stream, &rc, diag, p_pitch);
    if (rc) return rc;
    if (iters & 1)
        NLSB_CUDA(cudaMemcpyAsync(psi, work, sizeof(double2) * (size_t)batch * rows * cols, cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int enqueue_rk4_2d_staged(int batch, int rows, int cols, int order, int iters, double dt, const CrossWeights &w,
                          const double *pumping, const double *coeffs, double2 *psi, double2 *work, cudaStream_t stream);

// Planar path: psi is split into re/im planes once, every step is one TMA-fed fused launch that
// ping-pongs between two planar buffers, and the result is interleaved back at the end.
int enqueue_rk4_2d_planar(int batch, int rows, int cols, int order, int iters, double dt, const CrossWeights &w,
                          const double *pumping, const double *coeffs, double2 *psi, double2 *work, int variant,
                          bool allow_persistent, cudaStream_t stream)
{
    Fused2DPlanar p;
    p.batch = batch; p.rows = rows; p.cols = cols; p.pitch = planar_pitch(cols);
    p.grow0 = 0; p.grows = rows; p.out_row0 = 0; p.out_row1 = rows;
    p.psi_a = reinterpret_cast<double *>(work);
    p.psi_b = p.psi_a + planar_psi_doubles(batch, rows, cols);
    p.cp = p.psi_b + planar_psi_doubles(batch, rows, cols);
    p.coeffs = coeffs; p.uniform = g_uniform_coeffs; p.dt = dt;
    PlanarMaps maps;
    NLSB_TRY(make_planar_maps(order, variant, p, &maps));
    NLSB_TRY(launch_split_planar(p, psi, pumping, stream));

    // small grids: every tile resident at once -> the whole time loop in one cooperative launch
    bool fits = false;
    long long tiles = 0;
    if (allow_persistent && iters >= 2) NLSB_TRY(persistent_2d_fits(order, variant, p, &fits, &tiles));
    if (fits) {
        Arena mem(stream);
        int *flags;
        NLSB_TRY(mem.alloc(&flags, (size_t)tiles));
        NLSB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)tiles, stream));
        NLSB_TRY(launch_rk4_persistent_2d_planar(order, variant, p, maps, iters, flags, w, stream));
        NLSB_TRY(launch_join_planar(p, (iters & 1) != 0, psi, stream));
        return 0;
    }

    auto steps = [&](int first, int count, cudaStream_t s, int *rc) {
        for (int i = 0; i < count; ++i) {
            // the first step follows the split kernel, which writes c12*P: no early start for that one
            int r = launch_rk4_step_fused_2d_planar(order, variant, p, maps, ((first + i) & 1) == 0, first + i > 0, w, s);
            if (r && !*rc) *rc = r;
        }
    };
    int rc = 0, done = 0;
    const int chunk = 32;   // even: a replayed chunk starts and ends in buffer A
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NLSB_CUDA(cudaStreamIsCapturing(stream, &cap));
    if (cap == cudaStreamCaptureStatusNone && iters >= 2 * chunk) {
        // the chunk's launches depend on the planar working copy inside `work` (tensor maps by value), the geometry and
        // the coefficients: repeated calls on the same buffers replay the cached executable (capture + instantiation
        // cost about a millisecond per call -- 3 % of a 5000-step solve of the 512^2 grid)
        int krc = 0;
        const LoopKey key = make_loop_key(psi, work, pumping, coeffs, batch, rows, cols, order, 100 + variant, dt, w, &krc);
        if (krc) return krc;
        std::lock_guard<std::mutex> lock(g_graph_mutex);
        cudaGraphExec_t exec = nullptr;
        NLSB_TRY(cached_loop_graph(key, chunk, [&](cudaStream_t rec, int *r) { steps(0, chunk, rec, r); }, &exec));
        for (; done + chunk <= iters; done += chunk) {
            cudaError_t e = cudaGraphLaunch(exec, stream);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
            count_launches((unsigned long long)chunk);
        }
    }
    steps(done, iters - done, stream, &rc);
    if (rc) return rc;
    NLSB_TRY(launch_join_planar(p, (iters & 1) != 0, psi, stream));
    return 0;
}

// Resident path: returns 0 and sets *done when the grid fits and the launch was enqueued.
int try_rk4_2d_resident(int batch, int rows, int cols, int order, int iters, double dt, const CrossWeights &w,
                        const double *pumping, const double *coeffs, double2 *psi, cudaStream_t stream, bool *done)
{
    *done = false;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NLSB_CUDA(cudaStreamIsCapturing(stream, &cap));
    if (cap != cudaStreamCaptureStatusNone || iters < 1) return 0;
    Resident2D r{batch, rows, cols, iters, psi, pumping, coeffs, g_uniform_coeffs, nullptr, 0u, dt};
    bool fits = false;
    size_t bytes = 0;
    NLSB_TRY(resident_2d_query(order, r, &fits, &bytes));
    if (!fits) return 0;
    Arena mem(stream);
    char *mail;
    NLSB_TRY(mem.alloc(&mail, bytes));
    NLSB_CUDA(cudaMemsetAsync(mail, 0, bytes, stream));
    r.mailbox = mail;
    // sequence numbers are 32-bit and start from 0 in a fresh mailbox: split very long horizons
    const int most = 1 << 28;
    for (int left = iters; left > 0;) {
        r.steps = left < most ? left : most;
        NLSB_TRY(launch_rk4_resident_2d(order, r, w, stream));
        left -= r.steps;
        if (left > 0) NLSB_CUDA(cudaMemsetAsync(mail, 0, bytes, stream));
    }
    *done = true;
    return 0;
}

int enqueue_rk4_2d(int batch, int rows, int cols, int order, int iters, double dt, const CrossWeights &w,
                   const double *pumping, const double *coeffs, double2 *psi, double2 *work, cudaStream_t stream)
{
    const int path = g_path_2d.load();
    if (path == 9) {
        bool done = false;
        NLSB_TRY(try_rk4_2d_resident(batch, rows, cols, order, iters, dt, w, pumping, coeffs, psi, stream, &done));
        if (done) return 0;
    }
    if (path == 1)
        return enqueue_rk4_2d_staged(batch, rows, cols, order, iters, dt, w, pumping, coeffs, psi, work, stream);
    if (path == 2 || path == 3 || path == 8 || batch > 32767 || rows > 65535 ||
        (path == 0 && stream_preferred(order, batch, rows, cols)))
        return enqueue_rk4_2d_fused(batch, rows, cols, order, iters, dt, w, pumping, coeffs, psi, work, stream);
    int variant = (path == 5 || path == 7) ? 1 : 0;
    if (path == 0 && order != 7) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long small = (long long)((cols + 31) / 32) * ((rows + 31) / 32) * batch;
        const long long tall = (long long)((cols + 31) / 32) * ((rows + 63) / 64) * batch;
        if (small > sms && tall <= sms) variant = 1;
    }
    return enqueue_rk4_2d_planar(batch, rows, cols, order, iters, dt, w, pumping, coeffs, psi, work, variant,
                                 path == 6 || path == 7, stream);
}

int enqueue_rk4_2d_staged(int batch, int rows, int cols, int order, int iters, double dt, const CrossWeights &w,
                          const double *pumping, const double *coeffs, double2 *psi, double2 *work, cudaStream_t stream)
{
    int rc = 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NLSB_CUDA(cudaStreamIsCapturing(stream, &cap));
    const int chunk = 16;
    int done = 0;
    if (cap == cudaStreamCaptureStatusNone && iters >= 2 * chunk) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        // record on a private stream (the caller's may be the legacy default stream, which cannot
        // capture); the instantiated graph is then launched on the caller's stream
        cudaStream_t rec;
        NLSB_TRY(internal_stream(&rec));
        NLSB_CUDA(cudaStreamBeginCapture(rec, cudaStreamCaptureModeThreadLocal));
        for (int s = 0; s < chunk; ++s) enqueue_step_2d(batch, rows, cols, order, dt, w, pumping, coeffs, psi, work, rec, &rc);
        cudaError_t e = cudaStreamEndCapture(rec, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
        count_launches(0ull - 4ull * chunk);   // the capture pass above only recorded, it did not run
        for (; done + chunk <= iters; done += chunk) {
            e = cudaGraphLaunch(exec, stream);
            if (e != cudaSuccess) break;
            count_launches(4ull * chunk);
        }
        // the executable graph is released once its in-flight launches complete
        cudaGraphExecDestroy(exec);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
    }
    for (; done < iters; ++done) enqueue_step_2d(batch, rows, cols, order, dt, w, pumping, coeffs, psi, work, stream, &rc);
    return rc;
}

int weights_from_host(int order, const double *wx, const double *wy, CrossWeights *w)
{
    if (!valid_order(order)) return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
    if (!wx || !wy) return fail(NLSB_EINVAL, "null stencil weights");
    std::memset(w, 0, sizeof(*w));
    for (int t = 0; t < order; ++t) {
        w->wx[t] = wx[t];
        w->wy[t] = wy[t];
    }
    return 0;
}

// ---- shared bodies of the host-buffer entry points ----------------------------------------------
int host_rk4_1d(double dt, const double *taps_host, int n, int order, int iters, const double *pumping,
                const double *coeffs, const double *u0, double *u)
{
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    double *d_taps, *d_p, *d_c;
    double2 *d_psi;
    NLSB_TRY(mem.upload(&d_taps, taps_host, (size_t)n * order));
    NLSB_TRY(mem.upload(&d_p, pumping, (size_t)n));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.upload(&d_psi, reinterpret_cast<const double2 *>(u0), (size_t)n));
    NLSB_TRY(nlsb_dev_rk4_1d(1, n, order, iters, dt, d_taps, d_p, d_c, reinterpret_cast<double *>(d_psi), s));
    NLSB_CUDA(cudaMemcpyAsync(u, d_psi, sizeof(double2) * n, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ---- host-buffer solve of a LARGE grid with the transfers overlapped ---------------------------------------------
// A solve from host buffers is H2D (24 B per node) + iters steps + D2H (16 B per node); on 8192^2 with 200 steps the
// copies are 48 of 248 ms.  A node's arithmetic does not depend on the launch that computes it (the property the slab
// decomposition rests on), so the grid can be advanced in two row ranges that run AHEAD of each other:
//   start  the top rows [0, R + 4k s) (R = a quarter to a half of the grid) are uploaded first and take s steps while the
//          other rows are still on the bus -- step j on rows [0, R + 4k (s - j)): every step loses the 4k rows whose
//          neighbours are not there yet -- then the bottom range catches up, step j on the complementary rows
//          [R + 4k (s - j), n);
//   end    the top range [0, R') (all but the last quarter to half) runs s' steps ahead again (step j on
//          [0, R' + 4k (s' - j))) and leaves for the host while the bottom range takes its last s' steps.
// Both ranges ping-pong between the same two buffers; a range's step j never overwrites rows the other range's step j
// still reads (they lie 4k rows beyond its edge).  Bit-identical to the plain loop; s, s' are even so that the middle
// part starts and ends in `psi`.  No node is computed twice (the two ranges of a step are complementary), so with
// pageable host memory -- whose copies cudaMemcpyAsync stages synchronously: nothing overlaps -- the scheme costs only
// its extra launches.
int copy_stream(cudaStream_t *out)
{
    static thread_local cudaStream_t streams[64] = {};
    int dev = 0;
    NLSB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(NLSB_EINVAL, "device ordinal %d out of range", dev);
    if (!streams[dev]) NLSB_CUDA(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
    *out = streams[dev];
    return 0;
}

struct ScopedEvent {
    cudaEvent_t ev = nullptr;
    int create()
    {
        NLSB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        return 0;
    }
    ~ScopedEvent()
    {
        if (ev) cudaEventDestroy(ev);
    }
};

bool pipelined_solve_applies(int n, int order, int iters)
{
    const int path = g_path_2d.load();
    if (path != 0 && path != 8) return false;
    (void)order;
    // every row range must be a grid the strip-marching kernel takes (>= 2^20 nodes, even column count); pipeline_plan
    // keeps the head starts inside the grid whatever the order
    return (n & 1) == 0 && (long long)(n / 4) * n >= (1ll << 20) && iters >= 160;
}

// Split points and head starts of the overlapped solve (host arithmetic only; also behind nlsb_solve_nls_2d_plan).
struct PipelinePlan {
    int r_top, s_up;     // start: rows [0, r_top + 4k s_up) are uploaded first; the top range takes s_up steps ahead
    int r_dn, s_dn;      // end: the range [0, r_dn) takes s_dn steps ahead and leaves first; rows [r_dn, n) leave last
};

PipelinePlan pipeline_plan(int n, int order, int iters)
{
    const int halo = 2 * (order - 1);
    // The range that goes first (start) / last (end) is the fraction f of the rows; what stays exposed is ITS transfer,
    // so f should be small -- but the head start it needs grows as (1 - f) / f: the other range's 24 B (16 B) per node
    // cross the bus at about 55 GB/s while a node-step takes about 1.45e-11 s, i.e. 29 (19) steps of a range as large
    // as the one on the bus.  f = 1/4 when the solve is long enough for both head starts (0.8 iters), else up to 1/2.
    double f = 1.0 / (1.0 + iters / 62.0);
    f = f < 0.25 ? 0.25 : (f > 0.5 ? 0.5 : f);
    PipelinePlan p;
    p.r_top = ((int)(f * n) + 1) & ~1;
    p.r_dn = n - p.r_top;
    p.s_up = (int)((1.0 - f) / f * 29.0) & ~1;       // even: the middle part starts and ends in `psi`
    p.s_dn = (int)((1.0 - f) / f * 19.0) & ~1;
    if (p.s_up + p.s_dn > (iters * 4) / 5) {
        const double scale = 0.8 * iters / (p.s_up + p.s_dn);
        p.s_up = (int)(p.s_up * scale) & ~1;
        p.s_dn = (int)(p.s_dn * scale) & ~1;
    }
    while (p.s_up > 0 && p.r_top + halo * p.s_up > (n * 3) / 4) p.s_up -= 2;   // the first transfer stays the smaller one
    while (p.s_dn > 0 && p.r_dn + halo * p.s_dn > n - halo) p.s_dn -= 2;       // the head start ends inside the grid
    return p;
}

int host_rk4_2d_pipelined(double dt, const CrossWeights &w, int n, int order, int iters, const double *pumping,
                          const double *coeffs, const double *u0, double *u)
{
    cudaStream_t s, c;
    NLSB_TRY(internal_stream(&s));
    NLSB_TRY(copy_stream(&c));
    Arena mem(s);
    struct DrainOnExit {     // declared after the arena: an early return must not free buffers the copy stream still uses
        cudaStream_t st;
        ~DrainOnExit() { cudaStreamSynchronize(st); }
    } drain{c};
    const size_t np = (size_t)n * n, row = (size_t)n;
    const int halo = 2 * (order - 1);
    const PipelinePlan plan = pipeline_plan(n, order, iters);
    const int R_top = plan.r_top, R_dn = plan.r_dn, s_up = plan.s_up, s_dn = plan.s_dn;
    const int R_up = R_top + halo * s_up;
    double *d_p, *d_c;
    double2 *d_psi, *d_work;
    NLSB_TRY(mem.alloc(&d_p, np));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.alloc(&d_psi, np));
    NLSB_TRY(mem.alloc(&d_work, nlsb_dev_rk4_2d_workspace(1, n, n) / sizeof(double2) + 1));
    ScopedEvent ready, top_in, bot_in, top_out;
    NLSB_TRY(ready.create()); NLSB_TRY(top_in.create()); NLSB_TRY(bot_in.create()); NLSB_TRY(top_out.create());
    NLSB_CUDA(cudaEventRecord(ready.ev, s));                   // the allocations are ordered on s
    NLSB_CUDA(cudaStreamWaitEvent(c, ready.ev, 0));
    const double2 *h_u0 = reinterpret_cast<const double2 *>(u0);
    NLSB_CUDA(cudaMemcpyAsync(d_p, pumping, sizeof(double) * R_up * row, cudaMemcpyHostToDevice, c));
    NLSB_CUDA(cudaMemcpyAsync(d_psi, h_u0, sizeof(double2) * R_up * row, cudaMemcpyHostToDevice, c));
    NLSB_CUDA(cudaEventRecord(top_in.ev, c));
    NLSB_CUDA(cudaMemcpyAsync(d_p + R_up * row, pumping + R_up * row, sizeof(double) * (n - R_up) * row, cudaMemcpyHostToDevice, c));
    NLSB_CUDA(cudaMemcpyAsync(d_psi + R_up * row, h_u0 + R_up * row, sizeof(double2) * (n - R_up) * row, cudaMemcpyHostToDevice, c));
    NLSB_CUDA(cudaEventRecord(bot_in.ev, c));

    UniformCoeffsScope uniform(coeffs);
    Fused2DStep st{1, n, n, 0, n, 0, n, nullptr, nullptr, d_p, d_c, dt, g_uniform_coeffs};
    auto range_step = [&](int j, int row0, int row1) {          // step j (1-based within its phase): buffer parity j
        st.in = (j & 1) ? d_psi : d_work;
        st.out = (j & 1) ? d_work : d_psi;
        st.out_row0 = row0;
        st.out_row1 = row1;
        return launch_interleaved_step(order, st, w, s);
    };
    NLSB_CUDA(cudaStreamWaitEvent(s, top_in.ev, 0));
    for (int j = 1; j <= s_up; ++j) NLSB_TRY(range_step(j, 0, R_top + halo * (s_up - j)));
    NLSB_CUDA(cudaStreamWaitEvent(s, bot_in.ev, 0));
    for (int j = 1; j <= s_up; ++j) NLSB_TRY(range_step(j, R_top + halo * (s_up - j), n));
    // middle: whole-grid steps (cached graphs), state in d_psi before and after
    NLSB_TRY(enqueue_rk4_2d(1, n, n, order, iters - s_up - s_dn, dt, w, d_p, d_c, d_psi, d_work, s));
    for (int j = 1; j <= s_dn; ++j) NLSB_TRY(range_step(j, 0, R_dn + halo * (s_dn - j)));
    NLSB_CUDA(cudaEventRecord(top_out.ev, s));
    NLSB_CUDA(cudaStreamWaitEvent(c, top_out.ev, 0));
    double2 *h_u = reinterpret_cast<double2 *>(u);
    NLSB_CUDA(cudaMemcpyAsync(h_u, d_psi, sizeof(double2) * R_dn * row, cudaMemcpyDeviceToHost, c));
    for (int j = 1; j <= s_dn; ++j) NLSB_TRY(range_step(j, R_dn + halo * (s_dn - j), n));
    NLSB_CUDA(cudaMemcpyAsync(h_u + R_dn * row, d_psi + R_dn * row, sizeof(double2) * (n - R_dn) * row, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(c));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int host_rk4_2d(double dt, const CrossWeights &w, int n, int order, int iters, const double *pumping,
                const double *coeffs, const double *u0, double *u)
{
    if (pipelined_solve_applies(n, order, iters)) return host_rk4_2d_pipelined(dt, w, n, order, iters, pumping, coeffs, u0, u);
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    const size_t np = (size_t)n * n;
    double *d_p, *d_c;
    double2 *d_psi, *d_work;
    NLSB_TRY(mem.upload(&d_p, pumping, np));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.upload(&d_psi, reinterpret_cast<const double2 *>(u0), np));
    NLSB_TRY(mem.alloc(&d_work, nlsb_dev_rk4_2d_workspace(1, n, n) / sizeof(double2) + 1));
    UniformCoeffsScope uniform(coeffs);
    NLSB_TRY(enqueue_rk4_2d(1, n, n, order, iters, dt, w, d_p, d_c, d_psi, d_work, s));
    NLSB_CUDA(cudaMemcpyAsync(u, d_psi, sizeof(double2) * np, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int host_hamiltonian_1d(const double *taps_host, int n, int order, const double *pumping, const double *coeffs,
                        const double *u, double *v, double *mu_out /* optional: chemical potential */, double dx)
{
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    double *d_taps, *d_p, *d_c;
    double2 *d_u, *d_v;
    NLSB_TRY(mem.upload(&d_taps, taps_host, (size_t)n * order));
    NLSB_TRY(mem.upload(&d_p, pumping, (size_t)n));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.upload(&d_u, reinterpret_cast<const double2 *>(u), (size_t)n));
    NLSB_TRY(mem.alloc(&d_v, (size_t)n));
    NLSB_TRY(launch_hamiltonian_1d(1, n, order, d_taps, d_p, d_c, d_u, d_v, s));
    if (v) NLSB_CUDA(cudaMemcpyAsync(v, d_v, sizeof(double2) * n, cudaMemcpyDeviceToHost, s));
    if (mu_out) {
        char *scratch;
        double *d_out;
        NLSB_TRY(mem.alloc(&scratch, weighted_dots_scratch_bytes()));
        NLSB_TRY(mem.alloc(&d_out, (size_t)4));
        NLSB_TRY(launch_weighted_dots((size_t)n, dx, d_u, d_v, scratch, d_out, s));
        NLSB_CUDA(cudaMemcpyAsync(mu_out, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int host_hamiltonian_2d(const CrossWeights &w, int n, int order, const double *pumping, const double *coeffs,
                        const double *u, double *v, double *mu_out)
{
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    const size_t np = (size_t)n * n;
    double *d_p, *d_c;
    double2 *d_u, *d_v;
    NLSB_TRY(mem.upload(&d_p, pumping, np));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.upload(&d_u, reinterpret_cast<const double2 *>(u), np));
    NLSB_TRY(mem.alloc(&d_v, np));
    Stage2DArgs a{1, n, n, d_u, d_u, d_p, d_c, nullptr, d_v, 0.0, 0.0, 0.0};
    NLSB_TRY(launch_stage_2d(order, kStageRhs, w, a, s));
    if (v) NLSB_CUDA(cudaMemcpyAsync(v, d_v, sizeof(double2) * np, cudaMemcpyDeviceToHost, s));
    if (mu_out) {
        char *scratch;
        double *d_out;
        NLSB_TRY(mem.alloc(&scratch, weighted_dots_scratch_bytes()));
        NLSB_TRY(mem.alloc(&d_out, (size_t)4));
        NLSB_TRY(launch_weighted_dots(np, 0.0, d_u, d_v, scratch, d_out, s));
        NLSB_CUDA(cudaMemcpyAsync(mu_out, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ---- general block-band operators (blocks memory that is not a constant-weight cross stencil) ---------------------
// The fast kernels take (wx, wy); blocks_to_weights refuses everything else with NLSB_EOPERATOR.  When the blocks
// memory still has the make_laplacian_2d LAYOUT (orders = (0, .., k, .., 0): the only layout rbbmv's hard-wired block
// offsets can address, nls.f90:421-507) the entry points fall back to the general kernels of kernels_2d.cu.
bool standard_block_orders(int m, const int *orders)
{
    if (m != 3 && m != 5 && m != 7) return false;
    for (int b = 0; b < m; ++b)
        if (orders[b] != (b == (m - 1) / 2 ? (m - 1) / 2 : 0)) return false;
    return true;
}

int host_general_hamiltonian_2d(const double *blocks, int n, int m, const double *pumping, const double *coeffs,
                                const double *u, double *v)
{
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    const size_t np = (size_t)n * n;
    double *d_b, *d_p, *d_c;
    double2 *d_u, *d_v;
    NLSB_TRY(mem.upload(&d_b, blocks, (size_t)n * (2 * m - 1)));
    NLSB_TRY(mem.upload(&d_p, pumping, np));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.upload(&d_u, reinterpret_cast<const double2 *>(u), np));
    NLSB_TRY(mem.alloc(&d_v, np));
    Stage2DArgs a{1, n, n, d_u, d_u, d_p, d_c, nullptr, d_v, 0.0, 0.0, 0.0};
    NLSB_TRY(launch_stage_2d_general(n, m, d_b, kStageRhs, a, s));
    NLSB_CUDA(cudaMemcpyAsync(v, d_v, sizeof(double2) * np, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int host_general_rk4_2d(double dt, const double *blocks, int n, int m, int iters, const double *pumping,
                        const double *coeffs, const double *u0, double *u)
{
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    const size_t np = (size_t)n * n;
    double *d_b, *d_p, *d_c;
    double2 *d_psi, *d_work;
    NLSB_TRY(mem.upload(&d_b, blocks, (size_t)n * (2 * m - 1)));
    NLSB_TRY(mem.upload(&d_p, pumping, np));
    NLSB_TRY(mem.upload(&d_c, coeffs, (size_t)23));
    NLSB_TRY(mem.upload(&d_psi, reinterpret_cast<const double2 *>(u0), np));
    NLSB_TRY(mem.alloc(&d_work, 3 * np));
    double2 *ya = d_work, *yb = d_work + np, *acc = d_work + 2 * np;
    Stage2DArgs a{1, n, n, nullptr, d_psi, d_p, d_c, acc, nullptr, 0.0, 1.0, dt / 6};
    for (int it = 0; it < iters; ++it) {      // runge_kutta_2d, nls.f90:892-899: four right-hand sides per step
        a.ysrc = d_psi; a.ydst = ya; a.cy = dt / 2;
        NLSB_TRY(launch_stage_2d_general(n, m, d_b, kStageFirst, a, s));
        a.ysrc = ya; a.ydst = yb; a.cy = dt / 2;
        NLSB_TRY(launch_stage_2d_general(n, m, d_b, kStageMid, a, s));
        a.ysrc = yb; a.ydst = ya; a.cy = dt;
        NLSB_TRY(launch_stage_2d_general(n, m, d_b, kStageMid, a, s));
        a.ysrc = ya; a.ydst = d_psi; a.cy = 0.0;
        NLSB_TRY(launch_stage_2d_general(n, m, d_b, kStageLast, a, s));
    }
    NLSB_CUDA(cudaMemcpyAsync(u, d_psi, sizeof(double2) * np, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// mu = i*E'/M from the four sums {Re M, Im M, Re E', Im E'} (nls.f90:946-947, :968-970)
void chemical_potential_from_dots(const double d[4], double mu[2])
{
    const double er = -d[3], ei = d[2];   // (0,1) * E'
    const double den = d[0] * d[0] + d[1] * d[1];
    mu[0] = (er * d[0] + ei * d[1]) / den;
    mu[1] = (ei * d[0] - er * d[1]) / den;
}

}  // namespace
}  // namespace nlsb

using namespace nlsb;

extern "C" {

const char *nlsb_last_error(void) { return g_error; }

unsigned long long nlsb_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int nlsb_set_2d_path(int path)
{
    if (path < 0 || path > 9) return fail(NLSB_EINVAL, "2D path must be 0 (auto), 1 (per-stage) or 2..9 (fused step)");
    g_path_2d.store(path);
    return 0;
}

int nlsb_trim_memory(void)
{
    {
        std::lock_guard<std::mutex> lock(g_graph_mutex);
        for (LoopGraph &g : g_graphs) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
            g = LoopGraph{};
        }
    }
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (int dev = 0; dev < 64; ++dev)
        if (g_pools[dev]) {
            cudaError_t e = cudaMemPoolTrimTo(g_pools[dev], 0);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemPoolTrimTo");
        }
    return 0;
}

int nlsb_solve_nls_2d_plan(int n, int order, int iters, int *overlapped, int *r_top, int *s_up, int *r_dn, int *s_dn)
{
    if (!overlapped || !r_top || !s_up || !r_dn || !s_dn) return fail(NLSB_EINVAL, "solve_nls_2d_plan: null argument");
    NLSB_TRY(check_order_size(n, order));
    *overlapped = pipelined_solve_applies(n, order, iters) ? 1 : 0;
    const PipelinePlan p = pipeline_plan(n, order, iters);
    *r_top = p.r_top; *s_up = p.s_up; *r_dn = p.r_dn; *s_dn = p.s_dn;
    return 0;
}

int nlsb_set_stream_tuning(int sync, int width, int iters_per_cta)
{
    if (sync < -1 || sync > 1 || width < 0 || iters_per_cta < 0)
        return fail(NLSB_EINVAL, "stream tuning: sync in -1..1, width and iterations >= 0 (0 = automatic)");
    stream_2d_set_tuning(sync, width, iters_per_cta);
    return 0;
}

int nlsb_device_available(void)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count > 0 ? 1 : 0;
}

void nlsb_version(int *major, int *minor, int *patch)
{
    if (major) *major = 0;
    if (minor) *minor = 2;
    if (patch) *patch = 0;
}

// ---- operator builders (host) ---------------------------------------------------------------------
int nlsb_make_banded_matrix(int n, int m, const double *row, double *mat)
{
    if (!row || !mat || n < 1) return fail(NLSB_EINVAL, "make_banded_matrix: bad arguments");
    return banded_from_row(n, m, row, mat);
}

int nlsb_clear_first_row_of_derivative(int n, int m, double *L1)
{
    if (!L1 || n < 1 || m < 1 || (m & 1) == 0 || n < m) return fail(NLSB_EINVAL, "clear_first_row_of_derivative: bad arguments");
    const int k = (m - 1) / 2;
    for (int j = 0; j <= k; ++j) L1[(size_t)(k - j) + (size_t)m * j] = 0.0;   // entries A(0, j)
    return 0;
}

int nlsb_divide_derivative_on_radius(int n, int m, double h, double *L1)
{
    if (!L1 || n < 1 || m < 1 || (m & 1) == 0) return fail(NLSB_EINVAL, "divide_derivative_on_radius: bad arguments");
    const int k = (m - 1) / 2;
    for (int j = 0; j < n; ++j)
        for (int b = 0; b < m; ++b) {
            const int rho = j + b - k;   // matrix row of band entry (b, j)
            if (rho > 0) L1[(size_t)b + (size_t)m * j] = L1[(size_t)b + (size_t)m * j] / ((double)rho * h);
        }
    return 0;
}

int nlsb_radial_taps(int n, int m, double h, double *taps)
{
    if (!taps || n < 1) return fail(NLSB_EINVAL, "radial_taps: bad arguments");
    return radial_taps(n, m, h, taps);
}

int nlsb_band_to_taps(int n, int m, const double *op, double *taps)
{
    if (!op || !taps || n < 1) return fail(NLSB_EINVAL, "band_to_taps: bad arguments");
    return band_to_taps(n, m, op, taps);
}

int nlsb_make_laplacian(int n, int m, double h, double *op)
{
    if (!op || n < 1) return fail(NLSB_EINVAL, "make_laplacian: bad arguments");
    NLSB_TRY(check_order_size(n, m));
    std::vector<double> taps((size_t)n * m);
    NLSB_TRY(radial_taps(n, m, h, taps.data()));
    return taps_to_band(n, m, taps.data(), op);
}
int nlsb_make_laplacian_o3(int n, double h, double *op) { return nlsb_make_laplacian(n, 3, h, op); }
int nlsb_make_laplacian_o5(int n, double h, double *op) { return nlsb_make_laplacian(n, 5, h, op); }
int nlsb_make_laplacian_o7(int n, double h, double *op) { return nlsb_make_laplacian(n, 7, h, op); }

int nlsb_make_laplacian_2d(int n, int m, double h, double *blocks, int *orders)
{
    if (!blocks || !orders || n < 1) return fail(NLSB_EINVAL, "make_laplacian_2d: bad arguments");
    return cross_blocks(n, m, h, blocks, orders);
}
int nlsb_make_laplacian_2d_o3(int n, double h, double *b, int *o) { return nlsb_make_laplacian_2d(n, 3, h, b, o); }
int nlsb_make_laplacian_2d_o5(int n, double h, double *b, int *o) { return nlsb_make_laplacian_2d(n, 5, h, b, o); }
int nlsb_make_laplacian_2d_o7(int n, double h, double *b, int *o) { return nlsb_make_laplacian_2d(n, 7, h, b, o); }

int nlsb_cross_weights(int m, double h, double *wx, double *wy)
{
    if (!wx || !wy) return fail(NLSB_EINVAL, "cross_weights: bad arguments");
    return cross_weights(m, h, wx, wy);
}

int nlsb_blocks_to_weights(int n, int m, const double *blocks, const int *orders, double *wx, double *wy)
{
    if (!blocks || !orders || !wx || !wy) return fail(NLSB_EINVAL, "blocks_to_weights: bad arguments");
    return blocks_to_weights(n, m, blocks, orders, wx, wy);
}

// ---- matvecs / reservoir (host buffers) -----------------------------------------------------------
int nlsb_rgbmv(const double *x, double *u, double sign, const double *op, int klu, int n)
{
    if (!x || !u || !op || n < 1 || klu < 0 || klu > 3) return fail(NLSB_EINVAL, "rgbmv: bad arguments");
    const int m = 2 * klu + 1;
    if (n < m) return fail(NLSB_ESIZE, "n = %d is smaller than the band width %d", n, m);
    std::vector<double> taps((size_t)n * m);
    NLSB_TRY(band_to_taps(n, m, op, taps.data()));
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    double *d_taps, *d_x, *d_u;
    NLSB_TRY(mem.upload(&d_taps, taps.data(), taps.size()));
    NLSB_TRY(mem.upload(&d_x, x, (size_t)n));
    NLSB_TRY(mem.upload(&d_u, u, (size_t)n));
    NLSB_TRY(launch_band_matvec_1d(n, m, d_taps, d_x, d_u, sign, s));
    NLSB_CUDA(cudaMemcpyAsync(u, d_u, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int nlsb_rbbmv(const double *x, double *y, double sign, const double *blocks, const int *ms, int m, int n)
{
    if (!x || !y || !blocks || !ms || n < 1) return fail(NLSB_EINVAL, "rbbmv: bad arguments");
    CrossWeights w{};
    const int cross = blocks_to_weights(n, m, blocks, ms, w.wx, w.wy);
    if (cross && !(cross == NLSB_EOPERATOR && standard_block_orders(m, ms) && n >= m)) return cross;
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    const size_t np = (size_t)n * n;
    double *d_x, *d_y;
    NLSB_TRY(mem.upload(&d_x, x, np));
    NLSB_TRY(mem.upload(&d_y, y, np));
    if (cross) {       // weights that vary along the line: the general kernel, reading the blocks memory itself
        double *d_b;
        NLSB_TRY(mem.upload(&d_b, blocks, (size_t)n * (2 * m - 1)));
        NLSB_TRY(launch_general_matvec_2d(n, m, d_b, d_x, d_y, sign, s));
    } else {
        NLSB_TRY(launch_cross_matvec_2d(n, n, m, w, d_x, d_y, sign, s));
    }
    NLSB_CUDA(cudaMemcpyAsync(y, d_y, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}
int nlsb_rbbmv_o3(const double *x, double *y, double sg, const double *b, const int *ms, int n) { return nlsb_rbbmv(x, y, sg, b, ms, 3, n); }
int nlsb_rbbmv_o5(const double *x, double *y, double sg, const double *b, const int *ms, int n) { return nlsb_rbbmv(x, y, sg, b, ms, 5, n); }
int nlsb_rbbmv_o7(const double *x, double *y, double sg, const double *b, const int *ms, int n) { return nlsb_rbbmv(x, y, sg, b, ms, 7, n); }

static int host_reservoir(const double *pumping, const double *coeffs, const double *u_sqr, double *r, size_t np)
{
    if (!pumping || !coeffs || !u_sqr || !r) return fail(NLSB_EINVAL, "revervoir: null argument");
    cudaStream_t s;
    NLSB_TRY(internal_stream(&s));
    Arena mem(s);
    double *d_p, *d_q, *d_r;
    NLSB_TRY(mem.upload(&d_p, pumping, np));
    NLSB_TRY(mem.upload(&d_q, u_sqr, np));
    NLSB_TRY(mem.alloc(&d_r, np));
    NLSB_TRY(launch_reservoir(np, rhs_coeffs_from(coeffs), d_p, d_q, d_r, s));
    NLSB_CUDA(cudaMemcpyAsync(r, d_r, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int nlsb_revervoir(const double *pumping, const double *coeffs, const double *u_sqr, double *r, int n)
{
    if (n < 1) return fail(NLSB_EINVAL, "revervoir: n must be positive");
    return host_reservoir(pumping, coeffs, u_sqr, r, (size_t)n);
}

int nlsb_revervoir_2d(const double *pumping, const double *coeffs, const double *u_sqr, double *r, int n)
{
    if (n < 1) return fail(NLSB_EINVAL, "revervoir_2d: n must be positive");
    return host_reservoir(pumping, coeffs, u_sqr, r, (size_t)n * n);
}

// ---- right-hand side, time stepping, solve (host buffers) -----------------------------------------
int nlsb_hamiltonian(const double *pumping, const double *coeffs, const double *u, double *v, const double *op,
                     int klu, int n)
{
    if (!pumping || !coeffs || !u || !v || !op) return fail(NLSB_EINVAL, "hamiltonian: null argument");
    NLSB_TRY(check_coeffs(coeffs));
    const int m = 2 * klu + 1;
    NLSB_TRY(check_order_size(n, m));
    std::vector<double> taps((size_t)n * m);
    NLSB_TRY(band_to_taps(n, m, op, taps.data()));
    return host_hamiltonian_1d(taps.data(), n, m, pumping, coeffs, u, v, nullptr, 0.0);
}

int nlsb_hamiltonian_2d(const double *pumping, const double *coeffs, const double *u, double *v, const double *blocks,
                        const int *orders, int order, int n)
{
    if (!pumping || !coeffs || !u || !v || !blocks || !orders) return fail(NLSB_EINVAL, "hamiltonian_2d: null argument");
    NLSB_TRY(check_coeffs(coeffs));
    NLSB_TRY(check_order_size(n, order));
    CrossWeights w{};
    const int cross = blocks_to_weights(n, order, blocks, orders, w.wx, w.wy);
    if (cross == NLSB_EOPERATOR && standard_block_orders(order, orders))
        return host_general_hamiltonian_2d(blocks, n, order, pumping, coeffs, u, v);
    if (cross) return cross;
    return host_hamiltonian_2d(w, n, order, pumping, coeffs, u, v, nullptr);
}

int nlsb_runge_kutta(double dt, double t0, const double *u0, const double *op, int n, int order, int iters, double *u,
                     const double *pumping, const double *coeffs)
{
    (void)t0;   // the pumping is time independent; the reference advances t but never reads it
    if (!u0 || !op || !u || !pumping || !coeffs || iters < 0) return fail(NLSB_EINVAL, "runge_kutta: bad arguments");
    NLSB_TRY(check_coeffs(coeffs));
    NLSB_TRY(check_order_size(n, order));
    std::vector<double> taps((size_t)n * order);
    NLSB_TRY(band_to_taps(n, order, op, taps.data()));
    return host_rk4_1d(dt, taps.data(), n, order, iters, pumping, coeffs, u0, u);
}

int nlsb_runge_kutta_2d(double dt, double t0, const double *u0, int n, const double *blocks, const int *orders,
                        int order, int iters, double *u, const double *pumping, const double *coeffs)
{
    (void)t0;
    if (!u0 || !blocks || !orders || !u || !pumping || !coeffs || iters < 0)
        return fail(NLSB_EINVAL, "runge_kutta_2d: bad arguments");
    NLSB_TRY(check_coeffs(coeffs));
    NLSB_TRY(check_order_size(n, order));
    CrossWeights w{};
    const int cross = blocks_to_weights(n, order, blocks, orders, w.wx, w.wy);
    if (cross == NLSB_EOPERATOR && standard_block_orders(order, orders))
        return host_general_rk4_2d(dt, blocks, n, order, iters, pumping, coeffs, u0, u);
    if (cross) return cross;
    return host_rk4_2d(dt, w, n, order, iters, pumping, coeffs, u0, u);
}

int nlsb_solve_nls(double dt, double dx, int n, int order, int iters, const double *pumping, const double *coeffs,
                   const double *u0, double *u)
{
    if (!pumping || !coeffs || !u0 || !u || iters < 0) return fail(NLSB_EINVAL, "solve_nls: bad arguments");
    NLSB_TRY(check_coeffs(coeffs));
    NLSB_TRY(check_order_size(n, order));
    std::vector<double> taps((size_t)n * order);
    NLSB_TRY(radial_taps(n, order, dx, taps.data()));
    return host_rk4_1d(dt, taps.data(), n, order, iters, pumping, coeffs, u0, u);
}

int nlsb_solve_nls_1d(double dt, double dx, int n, int order, int iters, const double *pumping, const double *coeffs,
                      const double *u0, double *u)
{
    return nlsb_solve_nls(dt, dx, n, order, iters, pumping, coeffs, u0, u);
}

int nlsb_solve_nls_2d(double dt, double dx, int n, int order, int iters, const double *pumping, const double *coeffs,
                      const double *u0, double *u)
{
    if (!pumping || !coeffs || !u0 || !u || iters < 0) return fail(NLSB_EINVAL, "solve_nls_2d: bad arguments");
    NLSB_TRY(check_coeffs(coeffs));
    NLSB_TRY(check_order_size(n, order));
    CrossWeights w{};
    NLSB_TRY(cross_weights(order, dx, w.wx, w.wy));
    return host_rk4_2d(dt, w, n, order, iters, pumping, coeffs, u0, u);
}

int nlsb_chemical_potential_1d(double dx, int n, const double *pumping, const double *coeffs, const double *u0,
                               double *mu)
{
    if (!pumping || !coeffs || !u0 || !mu) return fail(NLSB_EINVAL, "chemical_potential_1d: null argument");
    NLSB_TRY(check_coeffs(coeffs));
    const int order = 5;   // hard-wired in the reference (nls.f90:931)
    NLSB_TRY(check_order_size(n, order));
    if (!(dx > 0.0)) return fail(NLSB_EINVAL, "chemical_potential_1d: dx must be positive");
    std::vector<double> taps((size_t)n * order);
    NLSB_TRY(radial_taps(n, order, dx, taps.data()));
    double dots[4];
    NLSB_TRY(host_hamiltonian_1d(taps.data(), n, order, pumping, coeffs, u0, nullptr, dots, dx));
    chemical_potential_from_dots(dots, mu);
    return 0;
}

int nlsb_chemical_potential_2d(double dx, int n, const double *pumping, const double *coeffs, const double *u0,
                               double *mu)
{
    if (!pumping || !coeffs || !u0 || !mu) return fail(NLSB_EINVAL, "chemical_potential_2d: null argument");
    NLSB_TRY(check_coeffs(coeffs));
    const int order = 5;   // nls.f90:958
    NLSB_TRY(check_order_size(n, order));
    CrossWeights w{};
    NLSB_TRY(cross_weights(order, dx, w.wx, w.wy));
    double dots[4], both[2];
    NLSB_TRY(host_hamiltonian_2d(w, n, order, pumping, coeffs, u0, nullptr, dots));
    chemical_potential_from_dots(dots, both);
    *mu = both[0];   // the reference's result is real(sp): the imaginary part is dropped (nls.f90:955)
    return 0;
}

// ---- device-resident entry points -----------------------------------------------------------------
int nlsb_dev_rk4_1d(int batch, int n, int order, int iters, double dt, const double *taps, const double *pumping,
                    const double *coeffs, double *psi, nlsb_stream_t stream)
{
    if (!taps || !pumping || !coeffs || !psi || batch < 1 || iters < 0) return fail(NLSB_EINVAL, "dev_rk4_1d: bad arguments");
    NLSB_TRY(check_order_size(n, order));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double2 *p = reinterpret_cast<double2 *>(psi);
    if (n <= kMaxResident1D) {
        NLSB_TRY(launch_rk4_1d(batch, n, order, iters, dt, taps, pumping, coeffs, p, s));
        return 0;
    }
    Arena mem(s);
    double2 *work;
    NLSB_TRY(mem.alloc(&work, 3 * (size_t)batch * n));
    NLSB_TRY(launch_rk4_1d_staged(batch, n, order, iters, dt, taps, pumping, coeffs, p, work, s));
    return 0;
}

int nlsb_dev_rk4_1d_diag(int batch, int n, int order, int iters, double dt, double dx, const double *taps,
                         const double *pumping, const double *coeffs, double *psi, void *diag_scratch, double *out8,
                         nlsb_stream_t stream)
{
    if (!taps || !pumping || !coeffs || !psi || !diag_scratch || !out8 || batch < 1 || iters < 1 || !(dx > 0.0))
        return fail(NLSB_EINVAL, "dev_rk4_1d_diag: bad arguments (iters must be >= 1)");
    NLSB_TRY(check_order_size(n, order));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double2 *p = reinterpret_cast<double2 *>(psi);
    if (n <= kMaxResident1D) {
        // The reduction belongs to the LAST step only, and the instantiation that carries it pays for it in every step
        // (its accumulators and branch cost registers the time loop then spills: measured +20 % on a 200-step chunk of
        // the C3 ensemble).  So: iters - 1 steps with the plain kernel, one step with the diagnostics-carrying one --
        // the field makes one extra round trip through HBM (32 B per node, once per chunk).
        if (iters > 1) NLSB_TRY(launch_rk4_1d(batch, n, order, iters - 1, dt, taps, pumping, coeffs, p, s, 0.0, nullptr));
        NLSB_TRY(launch_rk4_1d(batch, n, order, 1, dt, taps, pumping, coeffs, p, s, dx, out8));
        return 0;
    }
    // systems too large for one CTA: the stand-alone reduction on the state entering the last step
    if (iters > 1) NLSB_TRY(nlsb_dev_rk4_1d(batch, n, order, iters - 1, dt, taps, pumping, coeffs, psi, stream));
    NLSB_TRY(launch_diagnostics_1d(batch, n, order, dx, taps, pumping, coeffs, p, diag_scratch, out8, s));
    NLSB_TRY(nlsb_dev_rk4_1d(batch, n, order, 1, dt, taps, pumping, coeffs, psi, stream));
    return 0;
}

int nlsb_dev_hamiltonian_1d(int batch, int n, int order, const double *taps, const double *pumping,
                            const double *coeffs, const double *u, double *v, nlsb_stream_t stream)
{
    if (!taps || !pumping || !coeffs || !u || !v || batch < 1) return fail(NLSB_EINVAL, "dev_hamiltonian_1d: bad arguments");
    NLSB_TRY(check_order_size(n, order));
    NLSB_TRY(launch_hamiltonian_1d(batch, n, order, taps, pumping, coeffs, reinterpret_cast<const double2 *>(u),
                                   reinterpret_cast<double2 *>(v), static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_band_matvec_1d(int n, int order, const double *taps, const double *x, double *u, double sign,
                            nlsb_stream_t stream)
{
    if (!taps || !x || !u || n < 1) return fail(NLSB_EINVAL, "dev_band_matvec_1d: bad arguments");
    NLSB_TRY(launch_band_matvec_1d(n, order, taps, x, u, sign, static_cast<cudaStream_t>(stream)));
    return 0;
}

size_t nlsb_dev_rk4_2d_workspace(int batch, int rows, int cols)
{
    if (batch < 1 || rows < 1 || cols < 1) return 0;
    const size_t staged = sizeof(double2) * 3 * (size_t)batch * rows * cols;
    const size_t planar = sizeof(double) * (2 * planar_psi_doubles(batch, rows, cols) + planar_cp_doubles(batch, rows, cols));
    return (staged > planar ? staged : planar) + 256;
}

int nlsb_dev_rk4_2d(int batch, int rows, int cols, int order, int iters, double dt, const double *wx,
                    const double *wy, const double *pumping, const double *coeffs, const double *shared_coeffs_host,
                    double *psi, void *workspace, size_t workspace_bytes, nlsb_stream_t stream)
{
    UniformCoeffsScope uniform(shared_coeffs_host);
    if (!pumping || !coeffs || !psi || !workspace || batch < 1 || iters < 0) return fail(NLSB_EINVAL, "dev_rk4_2d: bad arguments");
    if (shared_coeffs_host) NLSB_TRY(check_coeffs(shared_coeffs_host));
    NLSB_TRY(check_order_size(rows < cols ? rows : cols, order));
    if (workspace_bytes < nlsb_dev_rk4_2d_workspace(batch, rows, cols))
        return fail(NLSB_EINVAL, "dev_rk4_2d: workspace of %zu bytes is too small", workspace_bytes);
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    NLSB_TRY(enqueue_rk4_2d(batch, rows, cols, order, iters, dt, w, pumping, coeffs, reinterpret_cast<double2 *>(psi),
                            static_cast<double2 *>(workspace), static_cast<cudaStream_t>(stream)));
    return 0;
}

size_t nlsb_dev_rk4_2d_diag_scratch(int batch, int rows, int cols, int order)
{
    if (batch < 1 || rows < 1 || cols < 1) return 0;
    size_t bytes = diagnostics_scratch_bytes(batch);
    const int parts = stream_2d_diag_parts(order, batch, rows, cols);
    const size_t fused = sizeof(double) * 8 * (size_t)(parts > 0 ? parts : 0) * batch;
    return (fused > bytes ? fused : bytes) + 256;
}

int nlsb_dev_rk4_2d_diag(int batch, int rows, int cols, int order, int iters, double dt, double dx, const double *wx,
                         const double *wy, const double *pumping, const double *coeffs, const double *shared_coeffs_host,
                         double *psi, void *workspace, size_t workspace_bytes, void *diag_scratch, double *out8,
                         nlsb_stream_t stream)
{
    UniformCoeffsScope uniform(shared_coeffs_host);
    if (!pumping || !coeffs || !psi || !workspace || !diag_scratch || !out8 || batch < 1 || iters < 1 || !(dx > 0.0))
        return fail(NLSB_EINVAL, "dev_rk4_2d_diag: bad arguments (iters must be >= 1)");
    if (shared_coeffs_host) NLSB_TRY(check_coeffs(shared_coeffs_host));
    NLSB_TRY(check_order_size(rows < cols ? rows : cols, order));
    if (workspace_bytes < nlsb_dev_rk4_2d_workspace(batch, rows, cols))
        return fail(NLSB_EINVAL, "dev_rk4_2d_diag: workspace of %zu bytes is too small", workspace_bytes);
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double2 *p = reinterpret_cast<double2 *>(psi);
    double2 *work = static_cast<double2 *>(workspace);
    const int path = g_path_2d.load();
    const bool fused = (path == 8 || (path == 0 && stream_preferred(order, batch, rows, cols))) && batch <= 65535;
    if (fused) {
        // the reduction rides in the first stage of the last step's launch; a tiny second launch sums the CTAs' partials
        const DiagRequest req{diag_scratch, dx * dx};
        NLSB_TRY(enqueue_rk4_2d_fused(batch, rows, cols, order, iters, dt, w, pumping, coeffs, p, work, s, &req));
        NLSB_TRY(launch_finish_diagnostics(batch, stream_2d_diag_parts(order, batch, rows, cols), diag_scratch, out8, s));
        return 0;
    }
    // kernels without the fused reduction: the stand-alone pass, on the same state (the one entering the last step)
    if (iters > 1) NLSB_TRY(enqueue_rk4_2d(batch, rows, cols, order, iters - 1, dt, w, pumping, coeffs, p, work, s));
    NLSB_TRY(launch_diagnostics_2d(batch, rows, cols, order, dx, w, pumping, coeffs, p, diag_scratch, out8, s));
    NLSB_TRY(enqueue_rk4_2d(batch, rows, cols, order, 1, dt, w, pumping, coeffs, p, work, s));
    return 0;
}

int nlsb_dev_rk4_step_2d_slab(int rows_alloc, int cols, int order, double dt, const double *wx, const double *wy,
                              int global_row0, int global_rows, int out_row0, int out_row1, const double *pumping,
                              const double *coeffs_host, const double *psi_in, double *psi_out, nlsb_stream_t stream)
{
    if (!pumping || !coeffs_host || !psi_in || !psi_out || psi_in == psi_out || rows_alloc < 1 || cols < 1)
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab: bad arguments");
    NLSB_TRY(check_coeffs(coeffs_host));
    if (out_row0 < 0 || out_row1 > rows_alloc || out_row0 > out_row1)
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab: output rows [%d, %d) outside the slab of %d rows", out_row0,
                    out_row1, rows_alloc);
    NLSB_TRY(check_order_size(global_rows < cols ? global_rows : cols, order));
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    const RhsCoeffs shared = rhs_coeffs_from(coeffs_host);
    Fused2DStep s{1, rows_alloc, cols, global_row0, global_rows, out_row0, out_row1,
                  reinterpret_cast<const double2 *>(psi_in), reinterpret_cast<double2 *>(psi_out), pumping, nullptr, dt,
                  &shared};
    NLSB_TRY(launch_interleaved_step(order, s, w, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_rk4_step_2d_slab_exchange(int rows_alloc, int cols, int order, double dt, const double *wx, const double *wy,
                                       int global_row0, int global_rows, int out_row0, int out_row1, const double *pumping,
                                       const double *coeffs_host, const double *psi_in, double *psi_out, int halo_rows,
                                       int up_row0, double *up_dst, int dn_row0, double *dn_dst, void *state,
                                       void *flags_mine, void *flags_up, void *flags_down, double timeout_seconds,
                                       nlsb_stream_t stream)
{
    if (!pumping || !coeffs_host || !psi_in || !psi_out || psi_in == psi_out || rows_alloc < 1 || cols < 1 || !state || !flags_mine)
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab_exchange: bad arguments");
    if (out_row0 < 0 || out_row1 > rows_alloc || out_row0 > out_row1 || halo_rows < 1)
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab_exchange: output rows [%d, %d) outside the slab of %d rows", out_row0,
                    out_row1, rows_alloc);
    if ((flags_up && (!up_dst || up_row0 < out_row0 || up_row0 + halo_rows > out_row1)) ||
        (flags_down && (!dn_dst || dn_row0 < out_row0 || dn_row0 + halo_rows > out_row1)))
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab_exchange: the rows a neighbour receives must be rows this step writes");
    NLSB_TRY(check_coeffs(coeffs_host));
    NLSB_TRY(check_order_size(global_rows < cols ? global_rows : cols, order));
    const int path = g_path_2d.load();
    if (!(path == 8 || (path == 0 && stream_preferred(order, 1, out_row1 - out_row0, cols))))
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab_exchange: this launch would not take the strip-marching kernel "
                                 "(use nlsb_dev_rk4_step_2d_slab + nlsb_dev_halo_exchange)");
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    const RhsCoeffs shared = rhs_coeffs_from(coeffs_host);
    PeerStep peer{};
    const char *out_bytes = reinterpret_cast<const char *>(psi_out);
    const size_t row_bytes = sizeof(double2) * (size_t)cols;
    peer.up_row0 = up_row0; peer.dn_row0 = dn_row0; peer.nrows = halo_rows;
    peer.up_delta = flags_up ? reinterpret_cast<const char *>(up_dst) - (out_bytes + row_bytes * up_row0) : 0;
    peer.dn_delta = flags_down ? reinterpret_cast<const char *>(dn_dst) - (out_bytes + row_bytes * dn_row0) : 0;
    peer.state = static_cast<unsigned long long *>(state);
    peer.flags_mine = static_cast<unsigned long long *>(flags_mine);
    peer.flags_up = static_cast<unsigned long long *>(flags_up);
    peer.flags_down = static_cast<unsigned long long *>(flags_down);
    peer.timeout_cycles = (long long)((timeout_seconds > 0.0 ? timeout_seconds : 5.0) * 1.9e9);
    Fused2DStep s{1, rows_alloc, cols, global_row0, global_rows, out_row0, out_row1,
                  reinterpret_cast<const double2 *>(psi_in), reinterpret_cast<double2 *>(psi_out), pumping, nullptr, dt,
                  &shared};
    s.peer = &peer;
    NLSB_TRY(launch_rk4_step_stream_2d(order, s, w, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_planar_pitch(int cols) { return cols > 0 ? planar_pitch(cols) : 0; }

int nlsb_dev_rk4_step_2d_slab_planar(int rows_alloc, int cols, int order, double dt, const double *wx,
                                     const double *wy, int global_row0, int global_rows, int out_row0, int out_row1,
                                     const double *cp, const double *coeffs_host, double *planes_in, double *planes_out,
                                     nlsb_stream_t stream)
{
    if (!cp || !coeffs_host || !planes_in || !planes_out || planes_in == planes_out || rows_alloc < 1 || cols < 1)
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab_planar: bad arguments");
    NLSB_TRY(check_coeffs(coeffs_host));
    if (out_row0 < 0 || out_row1 > rows_alloc || out_row0 > out_row1)
        return fail(NLSB_EINVAL, "dev_rk4_step_2d_slab_planar: output rows [%d, %d) outside the slab of %d rows",
                    out_row0, out_row1, rows_alloc);
    NLSB_TRY(check_order_size(global_rows < cols ? global_rows : cols, order));
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    const RhsCoeffs shared = rhs_coeffs_from(coeffs_host);
    const int variant = (g_path_2d.load() == 5 || g_path_2d.load() == 7) ? 1 : 0;
    Fused2DPlanar p;
    p.batch = 1; p.rows = rows_alloc; p.cols = cols; p.pitch = planar_pitch(cols);
    p.grow0 = global_row0; p.grows = global_rows; p.out_row0 = out_row0; p.out_row1 = out_row1;
    p.psi_a = planes_in; p.psi_b = planes_out; p.cp = const_cast<double *>(cp);
    p.coeffs = nullptr; p.uniform = &shared; p.dt = dt;
    PlanarMaps maps;
    NLSB_TRY(make_planar_maps(order, variant, p, &maps));
    NLSB_TRY(launch_rk4_step_fused_2d_planar(order, variant, p, maps, true, false, w, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_hamiltonian_2d(int batch, int rows, int cols, int order, const double *wx, const double *wy,
                            const double *pumping, const double *coeffs, const double *u, double *v,
                            nlsb_stream_t stream)
{
    if (!pumping || !coeffs || !u || !v || batch < 1) return fail(NLSB_EINVAL, "dev_hamiltonian_2d: bad arguments");
    NLSB_TRY(check_order_size(rows < cols ? rows : cols, order));
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    const double2 *uu = reinterpret_cast<const double2 *>(u);
    Stage2DArgs a{batch, rows, cols, uu, uu, pumping, coeffs, nullptr, reinterpret_cast<double2 *>(v), 0.0, 0.0, 0.0};
    NLSB_TRY(launch_stage_2d(order, kStageRhs, w, a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_cross_matvec_2d(int rows, int cols, int order, const double *wx, const double *wy, const double *x,
                             double *y, double sign, nlsb_stream_t stream)
{
    if (!x || !y || rows < 1 || cols < 1) return fail(NLSB_EINVAL, "dev_cross_matvec_2d: bad arguments");
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    NLSB_TRY(launch_cross_matvec_2d(rows, cols, order, w, x, y, sign, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_rk4_2d_plan(int batch, int rows, int cols, int order, int *kernel, int *threads, int *strips,
                         int *chunk_rows)
{
    if (!kernel || !threads || !strips || !chunk_rows || batch < 1) return fail(NLSB_EINVAL, "dev_rk4_2d_plan: bad arguments");
    NLSB_TRY(check_order_size(rows < cols ? rows : cols, order));
    const int path = g_path_2d.load();
    *strips = *chunk_rows = 0;
    if (path == 8 || (path == 0 && stream_preferred(order, batch, rows, cols))) {
        *kernel = 2;                          // odd widths too: the time loop hands over a pitched copy of the pumping
        return stream_2d_plan(order, batch, rows, cols, threads, strips, chunk_rows);
    }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaGetLastError();
    const long long small = (long long)((cols + 31) / 32) * ((rows + 31) / 32) * batch;
    const long long tall = (long long)((cols + 31) / 32) * ((rows + 63) / 64) * batch;
    const bool tall_tiles = path == 3 || path == 5 || path == 7 || (path == 0 && order != 7 && small > sms && tall <= sms);
    *kernel = path == 1 ? 3 : path == 9 ? 4 : tall_tiles ? 1 : 0;
    *threads = tall_tiles ? 512 : 256;
    return 0;
}

int nlsb_dev_pumping_profiles(int dim, int kind, int batch, int n, double dx, const double *params_host, double *out,
                              nlsb_stream_t stream)
{
    if (!params_host || !out || batch < 1 || n < 1 || (dim != 1 && dim != 2) || (kind != 0 && kind != 1) || !(dx > 0.0))
        return fail(NLSB_EINVAL, "dev_pumping_profiles: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Arena mem(s);
    double *d_params;
    NLSB_TRY(mem.upload(&d_params, params_host, (size_t)5 * batch));
    NLSB_TRY(launch_pumping_profiles(dim, kind, batch, n, dx, d_params, out, s));
    // the parameter table is pageable host memory: the copy above must have left it before we return
    NLSB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

size_t nlsb_dev_diagnostics_scratch(int batch) { return batch > 0 ? diagnostics_scratch_bytes(batch) : 0; }

int nlsb_dev_diagnostics_1d(int batch, int n, int order, double dx, const double *taps, const double *pumping,
                            const double *coeffs, const double *psi, void *scratch, double *out8, nlsb_stream_t stream)
{
    if (!taps || !pumping || !coeffs || !psi || !scratch || !out8 || batch < 1 || !(dx > 0.0))
        return fail(NLSB_EINVAL, "dev_diagnostics_1d: bad arguments");
    NLSB_TRY(check_order_size(n, order));
    NLSB_TRY(launch_diagnostics_1d(batch, n, order, dx, taps, pumping, coeffs, reinterpret_cast<const double2 *>(psi),
                                   scratch, out8, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_diagnostics_2d(int batch, int rows, int cols, int order, double dx, const double *wx, const double *wy,
                            const double *pumping, const double *coeffs, const double *psi, void *scratch, double *out8,
                            nlsb_stream_t stream)
{
    if (!pumping || !coeffs || !psi || !scratch || !out8 || batch < 1 || !(dx > 0.0))
        return fail(NLSB_EINVAL, "dev_diagnostics_2d: bad arguments");
    NLSB_TRY(check_order_size(rows < cols ? rows : cols, order));
    CrossWeights w{};
    NLSB_TRY(weights_from_host(order, wx, wy, &w));
    NLSB_TRY(launch_diagnostics_2d(batch, rows, cols, order, dx, w, pumping, coeffs, reinterpret_cast<const double2 *>(psi),
                                   scratch, out8, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_divide_check(size_t n, const double *a, const double *b, double *fast, double *exact, nlsb_stream_t stream)
{
    if (!a || !b || !fast || !exact) return fail(NLSB_EINVAL, "dev_divide_check: null argument");
    NLSB_TRY(launch_divide_check(n, a, b, fast, exact, static_cast<cudaStream_t>(stream)));
    return 0;
}

int nlsb_dev_reservoir(size_t npts, const double *coeffs_host, const double *pumping, const double *u_sqr, double *r,
                       nlsb_stream_t stream)
{
    if (!coeffs_host || !pumping || !u_sqr || !r) return fail(NLSB_EINVAL, "dev_reservoir: null argument");
    NLSB_TRY(launch_reservoir(npts, rhs_coeffs_from(coeffs_host), pumping, u_sqr, r, static_cast<cudaStream_t>(stream)));
    return 0;
}

}  // extern "C"
