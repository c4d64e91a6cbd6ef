// resident_2d.cu -- the whole RK4 time loop of a small 2D grid in ONE cooperative launch (sm_100a): one patch
// per CTA, the field in registers, edge nodes exchanged with the neighbouring CTAs once per RK stage through
// sequence-numbered packets in L2 (resident_2d_core.cuh).  Stands in for runge_kutta_2d (nls.f90:873-901) on
// grids whose patches are all resident at once (<= 128 x 15 nodes per CTA: 512 x 512 on the 148 SMs of a B200).
//
// EXPERIMENTAL path (nlsb_set_2d_path(9)); not chosen automatically.  Measured on B200, 512 x 512: 10.0 us per RK
// step against 8.6 us for one tile-kernel launch per step.  No work is redundant and nothing is launched, but the
// four exchanges per step each cost a store -> L2 -> poll -> barrier chain of about 3000 SM cycles that phase A
// (400 cycles) cannot hide: the step is bounded by L2 latency, not by arithmetic (FP64 pipe 24 % busy).
// profiles/r1_ncu_resident_c2_512_experimental.txt.

#include "kernels.h"
#include "resident_2d_core.cuh"

#include <cooperative_groups.h>

namespace nlsb {

namespace {

using namespace resident2d;

struct ResidentArgs {
    Layout layout;
    const double2 *in;       // [batch][rows][cols]
    double2 *out;            // [batch][rows][cols] (may alias `in`: every CTA reads its frame before the first exchange)
    const double *pumping;   // [batch][rows][cols]
    const double *coeffs;    // [batch][23] -- unused when UNIFORM
    RhsCoeffs cu;
    Packet *mail;            // zeroed; Cfg::mailbox_bytes(patches)
    double dt;
    int steps;
    uint32_t seq0;           // sequence number before the first stage (packets of earlier launches are older)
};

template <int K>
struct ResidentWeights {
    double wx[2 * K + 1];
    double wy[2 * K + 1];
};

template <typename C, bool UNIFORM>
__global__ void __launch_bounds__(C::T, 1)
rk4_resident_2d_kernel(const __grid_constant__ ResidentArgs a, const __grid_constant__ ResidentWeights<C::K> wa)
{
    constexpr int K = C::K, RT = C::RT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *frame0 = reinterpret_cast<double2 *>(smem_raw), *frame1 = frame0 + C::FRAME;
    double2 *psi_plane = frame1 + C::FRAME;
    double *cp_plane = reinterpret_cast<double *>(psi_plane + C::PLANE);

    const int tid = threadIdx.x;
    const int x = tid % C::TX, r0 = (tid / C::TX) * RT;
    const Patch p = make_patch(a.layout, (int)blockIdx.x);
    const size_t plane = (size_t)a.layout.rows * a.layout.cols;
    const double2 *__restrict__ in = a.in + p.member * plane;
    const double *__restrict__ P = a.pumping + p.member * plane;
    double2 *out = a.out + p.member * plane;
    RhsCoeffs cl;
    if (!UNIFORM) cl = load_rhs_coeffs(a.coeffs + (size_t)p.member * 23);
    const RhsCoeffs &c = UNIFORM ? a.cu : cl;
    const double half_dt = a.dt / 2, dt6 = a.dt / 6;

    // ---- frame 0 = psi on the patch and its halo (zero outside the grid: the truncated stencil), frame 1 = 0 ----
    for (int i = tid; i < C::FRAME; i += C::T) {
        const int fr = i / C::FP - K, fc = i % C::FP - K;
        const int gr = p.row0 + fr, gc = p.col0 + fc;
        double2 v = make_double2(0.0, 0.0);
        if (gr >= 0 && gr < a.layout.rows && gc >= 0 && gc < a.layout.cols) v = in[(size_t)gr * a.layout.cols + gc];
        frame0[i] = v;
        frame1[i] = make_double2(0.0, 0.0);
    }
    State<C> s;
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int row = r0 + i;
        const bool inside = row < p.ph && x < p.pw;
        const size_t g = (size_t)(p.row0 + row) * a.layout.cols + p.col0 + x;
        s.y[i] = inside ? in[g] : make_double2(0.0, 0.0);
        psi_plane[row * C::TX + x] = s.y[i];
        cp_plane[row * C::TX + x] = inside ? c.c12 * P[g] : 0.0;     // c12 * P, rounded once (nls.f90:580 association)
        s.acc[i] = make_double2(0.0, 0.0);
    }
    // the halo cells this thread fetches every stage: source packets (parity 0) and frame index
    const Packet *src[C::NCELL];
    int dst[C::NCELL];
#pragma unroll
    for (int n = 0; n < C::NCELL; ++n) {
        const int cell = tid + n * C::T;
        int d = 0, sp = -1, sc = 0;
        const bool ok = cell < C::MB_CELLS && halo_cell<C>(p, cell, d, sp, sc);
        src[n] = ok ? mailbox_cell<C>(a.mail, sp, 0, sc) : nullptr;
        dst[n] = d;
    }
    // every CTA must have read its input frame before anyone overwrites `out` (which may alias `in`)
    cooperative_groups::this_grid().sync();

    uint32_t g = a.seq0;                 // sequence number of the current stage input
    bool first = true;
    auto stage = [&](auto tag, bool last) {
        constexpr int S = decltype(tag)::value;
        double2 *cur = (g & 1u) ? frame1 : frame0, *nxt = (g & 1u) ? frame0 : frame1;
        // all of this thread's halo packets are requested at once, BEFORE phase A (whose arithmetic hides part of
        // the L2 round trip); afterwards only the packets that do not carry sequence number g yet are requested
        // again.  (Measured alternatives: polling after phase A, or one polling lane per warp -- both slower.)
        Packet re[C::NCELL], im[C::NCELL];
        const size_t par = (size_t)(g & 1u) * C::MB_CELLS * 2;
        if (!first) {
#pragma unroll
            for (int n = 0; n < C::NCELL; ++n)
                if (src[n]) {
                    re[n] = packet_read(src[n] + par);
                    im[n] = packet_read(src[n] + par + 1);
                }
        }
        phase_a<C>(s, c, cp_plane, x, r0, wa.wx[C::K]);
        if (!first) {
            const long long t0 = clock64();
            for (;;) {
                bool all = true;
#pragma unroll
                for (int n = 0; n < C::NCELL; ++n)
                    if (src[n] && !(packet_ok(re[n], g) && packet_ok(im[n], g))) {
                        all = false;
                        re[n] = packet_read(src[n] + par);
                        im[n] = packet_read(src[n] + par + 1);
                    }
                if (all) break;
                if (clock64() - t0 > 4000000000ll) __trap();    // a neighbour never published: fail loudly
            }
#pragma unroll
            for (int n = 0; n < C::NCELL; ++n)
                if (src[n]) cur[dst[n]] = make_double2(packet_value(re[n]), packet_value(im[n]));
        }
        __syncthreads();
        phase_b<C, S>(s, p, x, r0, psi_plane, cur, nxt, a.mail, g + 1, last ? out : nullptr, (size_t)a.layout.cols, wa.wx, wa.wy,
                      half_dt, a.dt, dt6);
        g += 1;
        first = false;
    };
    for (int n = 0; n < a.steps; ++n) {
        stage(std::integral_constant<int, 1>{}, false);
        stage(std::integral_constant<int, 2>{}, false);
        stage(std::integral_constant<int, 3>{}, false);
        stage(std::integral_constant<int, 4>{}, n == a.steps - 1);
    }
}

template <typename C, bool UNIFORM>
int resident_capacity(long long *capacity)
{
    static long long cached[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return fail(NLSB_EINVAL, "device ordinal %d out of range", dev);
    if (!cached[dev]) {
        e = cudaFuncSetAttribute(rk4_resident_2d_kernel<C, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0, sms = 0, coop = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rk4_resident_2d_kernel<C, UNIFORM>, C::T, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (e != cudaSuccess) return (int)e;
        cached[dev] = (coop && per_sm * sms > 0) ? (long long)per_sm * sms : -1;
    }
    *capacity = cached[dev];
    return 0;
}

template <typename C, bool UNIFORM>
int resident_cfg(const Resident2D &r, const CrossWeights &w, bool query_only, bool *fits, size_t *mail_bytes,
                 cudaStream_t stream)
{
    long long capacity = 0;
    int rc = resident_capacity<C, UNIFORM>(&capacity);
    if (rc) return rc;
    Layout l;
    const bool ok = capacity > 0 && r.batch > 0 && r.batch <= capacity && make_layout<C>(r.batch, r.rows, r.cols, capacity, &l);
    if (fits) *fits = ok;
    if (!ok) return query_only ? 0 : fail(NLSB_ESIZE, "resident 2D kernel: %d x %d x %d does not fit on the device", r.batch, r.rows, r.cols);
    const long long patches = (long long)l.npx * l.npy * l.batch;
    if (mail_bytes) *mail_bytes = C::mailbox_bytes(patches);
    if (query_only) return 0;
    if (r.steps <= 0) return 0;
    ResidentArgs a{};
    a.layout = l;
    a.in = r.psi; a.out = r.psi; a.pumping = r.pumping; a.coeffs = r.coeffs;
    if (UNIFORM) a.cu = *r.uniform;
    a.mail = static_cast<Packet *>(r.mailbox);
    a.dt = r.dt; a.steps = r.steps; a.seq0 = r.seq0;
    ResidentWeights<C::K> wa;
    for (int i = 0; i < C::NW; ++i) { wa.wx[i] = w.wx[i]; wa.wy[i] = w.wy[i]; }
    void *args[] = {&a, &wa};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)rk4_resident_2d_kernel<C, UNIFORM>, dim3((unsigned)patches),
                                                dim3(C::T), args, C::SMEM, stream);
    if (e != cudaSuccess) return (int)e;
    count_launches(1);
    return 0;
}

template <int K>
int resident_k(const Resident2D &r, const CrossWeights &w, bool query_only, bool *fits, size_t *mail_bytes,
               cudaStream_t stream)
{
    using C = Cfg<K>;
    return r.uniform ? resident_cfg<C, true>(r, w, query_only, fits, mail_bytes, stream)
                     : resident_cfg<C, false>(r, w, query_only, fits, mail_bytes, stream);
}

int resident_dispatch(int order, const Resident2D &r, const CrossWeights &w, bool query_only, bool *fits,
                      size_t *mail_bytes, cudaStream_t stream)
{
    switch (order) {
    case 3: return resident_k<1>(r, w, query_only, fits, mail_bytes, stream);
    case 5: return resident_k<2>(r, w, query_only, fits, mail_bytes, stream);
    case 7: return resident_k<3>(r, w, query_only, fits, mail_bytes, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

}  // namespace

int resident_2d_query(int order, const Resident2D &r, bool *fits, size_t *mailbox_bytes)
{
    CrossWeights w{};
    return resident_dispatch(order, r, w, true, fits, mailbox_bytes, nullptr);
}

int launch_rk4_resident_2d(int order, const Resident2D &r, const CrossWeights &w, cudaStream_t stream)
{
    return resident_dispatch(order, r, w, false, nullptr, nullptr, stream);
}

}  // namespace nlsb
