// stream_2d_core.cuh -- per-thread body of the STREAMING fused RK4 step of the 2D solver.
//
// Reference semantics: one iteration of runge_kutta_2d (nls.f90:892-899) = four hamiltonian_2d evaluations
// (:841-870; cross stencil of make_laplacian_2d :297-385; reservoir :829-839) and the update
// u + (k1 + 2 k2 + 2 k3 + k4) dt/6.
//
// Decomposition: a CTA owns a STRIP of W = T - 8K columns and marches down a CHUNK of rows, one row per
// iteration.  Thread t owns frame column t (the strip plus a 4K-column halo each side).  The four RK stages
// run skewed by K rows: in iteration `it` (row j = jstart + it)
//       stage 1 evaluates row j, stage 2 row j-K, stage 3 row j-2K, stage 4 row j-3K,
// so every y-neighbour a stage needs has been produced by THIS thread (this or an earlier iteration) and
// lives in a register window; only the 2K x-neighbours of each stage come from shared memory (rings the
// owning threads publish to).  Per node-step: 4*2K + 1 shared loads and 3 shared stores of 16 bytes, the
// redundant work is (T / W) * (1 + 6K / H) instead of the tile kernel's 1.43.
//
// Register windows are circular with compile-time indices: the march loop is unrolled U = 2(2K+1) times
// and `ph` = it mod U is a constant in every unrolled copy, so no register is ever moved.
//     psi, acc, cp : period U     row j+d  <->  index (ph + d + 3K) mod U      (psi: d in [-3K, K])
//     y2, y3, y4   : period 2K+1  newest row at (ph + 2K), centre at (ph + K), oldest at ph
// The shared-memory rings use the same periods, so every shared address of an unrolled copy is one of two base
// registers plus an immediate: stage rings have NW = 2K+1 rows (row written in iteration `it` lives in slot
// it mod NW and is read K iterations later), the psi ring has 2U rows = 4 TMA batches of NW rows (the row
// `rel` rows after the first row of batch 0 lives in slot rel mod 2U; `half` = (it / U) mod 2 selects which
// half of the ring the unrolled copy starts in).
//
// The arithmetic of a node (order of the FMA chain of the stencil, stage algebra, domain mask) is the one of
// fused_2d.cu, so both kernels produce bit-identical fields.
//
// This header is compiled for the device (stream_2d.cu) and for the host (tests/emu/stream_emu.cu runs the
// same body thread by thread so that the indexing can be tested without a GPU).
#pragma once

#include "device_math.cuh"
#include "diag_acc.cuh"

#define NLSB_HD __host__ __device__ __forceinline__

namespace nlsb {
namespace stream2d {

template <int K_, int T_>
struct Cfg {
    static constexpr int K = K_, T = T_;
    static constexpr int NW = 2 * K_ + 1;       // taps per axis = period of the y windows and of the stage rings
    static constexpr int U = 2 * NW;            // unroll of the march = period of the psi / acc / cp windows
    static constexpr int RB = NW, NB = 4;       // rows per TMA batch, batches in the psi ring
    static constexpr int CTAS_PER_SM = T_ <= 128 ? 2 : 1;   // register budget: 65536 / (T * CTAS_PER_SM) >= 255
    static constexpr int HALO = 4 * K_;         // frame columns each side of the strip
    static constexpr int W = T_ - 2 * HALO;     // columns a strip produces
    static constexpr int RING = RB * NB;        // rows of the psi ring (= 2U)
    static constexpr int YS = NW;               // rows of a stage ring: written in iteration it, read in it + K
    static constexpr int YP = T_ + 2 * K_;      // pitch of a stage ring (K pad columns each side)
    static constexpr int SKEW = 3 * K_;         // rows between stage 1 and stage 4
    // shared memory: [3 stage rings][psi ring][pumping ring][mbarriers]; a psi x-neighbour read of an edge thread
    // may leave its row by K elements (into the stage rings below / the pumping ring above): harmless, those
    // threads carry garbage
    static constexpr size_t YRING_BYTES = sizeof(double2) * YS * YP;
    static constexpr size_t RING_OFFSET = (3 * YRING_BYTES + 127) / 128 * 128;
    static constexpr size_t RING_BYTES = sizeof(double2) * RING * T_;
    static constexpr size_t PRING_OFFSET = RING_OFFSET + RING_BYTES;
    static constexpr size_t PRING_BYTES = sizeof(double) * RING * T_;
    static constexpr size_t BAR_OFFSET = PRING_OFFSET + PRING_BYTES;
    static constexpr size_t SMEM = BAR_OFFSET + 8 * (NB + NW) + 64;   // NB batch barriers + NW split-phase barriers
    static_assert(RING == 2 * U, "psi ring = two unrolled march bodies");
    static_assert(RB >= 2 * K_, "the first TMA batch must hold the 2K rows the march starts from");
    static_assert((sizeof(double) * RB * T_) % 128 == 0, "TMA batches must stay 128-byte aligned");
    static_assert(W > 0, "strip narrower than its halo");
    static_assert(4 * K_ < U, "windows do not fit their period");
    // a chunk of H rows takes H + 6K iterations: H is chosen so that this is a multiple of U
    static constexpr int chunk_rows(int target_iters) { return (target_iters + U - 1) / U * U - 6 * K_; }
};

// Geometry of one CTA (uniform over its threads).
struct Chunk {
    int r0, r1;        // local rows [r0, r1) this CTA produces
    int jstart;        // row of stage 1 in iteration 0 (= r0 - 3K)
    int base;          // first row of TMA batch 0 (= jstart - K)
    int niter;         // iterations (multiple of U)
    int nbatches;      // TMA batches the march consumes
    int c0;            // first column of the strip
};

template <class C>
NLSB_HD Chunk make_chunk(int strip, int chunk, int rows_per_chunk, int out_row0, int out_row1)
{
    Chunk g;
    g.r0 = out_row0 + chunk * rows_per_chunk;
    g.r1 = g.r0 + rows_per_chunk < out_row1 ? g.r0 + rows_per_chunk : out_row1;
    g.jstart = g.r0 - C::SKEW;
    g.base = g.jstart - C::K;
    g.niter = (g.r1 - g.r0 + 2 * C::SKEW + C::U - 1) / C::U * C::U;
    g.nbatches = (g.niter + 2 * C::K + C::RB - 1) / C::RB;   // rows base .. base + niter - 1 + 2K
    g.c0 = strip * C::W;
    return g;
}

// Threads of a strip that have work.  The frame columns to the right of the domain's last column hold zeros in every
// stage (truncated stencil): the TMA fills them with zeros in the psi ring and the stage rings keep the zeros they
// are initialised with when nobody writes them, so the warps that would own only such columns leave right after the
// set-up.  Only the LAST strip of a row of strips is narrower than T: a 1024-column member cut into 112-column
// strips ends in a strip of 16 columns = one warp of 128 threads (1184 swept columns per row instead of 1280).
template <class C>
NLSB_HD int active_threads(const Chunk &g, int cols)
{
    const int in_domain = cols - (g.c0 - C::HALO);          // frame columns [0, in_domain) lie left of the domain's end
    const int need = in_domain < C::T ? in_domain : C::T;
    return (need + 31) / 32 * 32;
}

// What a thread keeps in registers for the whole march.
template <class C>
struct State {
    double2 psi[C::U], acc[C::U];
    double cp[C::U];
    double2 y2[C::NW], y3[C::NW], y4[C::NW];
    double2 *onext;            // output of this column, row j - 3K
};

// Loop-invariant values of one thread.
template <class C>
struct Lane {
    const double2 *ring;       // psi ring, [RING][T]
    double2 *yr;               // three stage rings, [3][YS][YP]
    double2 *out;              // the member's output, already offset to this thread's column
    size_t pitch;              // elements between rows of P / out (= cols)
    int fx;                    // frame column = thread index
    int dlo, dspan;            // local rows [dlo, dlo + dspan) lie inside the local array AND the global domain
    bool col_in;               // the column lies inside the domain
    bool col_owned;            // ... and inside the strip: this thread writes the new psi
    double half_dt, dt, dt6;
    // PEER (multi-GPU slabs, last step before a halo exchange): rows [up0, up0 + nup) / [dn0, dn0 + ndn) of the new psi
    // are ALSO stored into the neighbouring ranks' halo rows, at the address of the local store plus these byte offsets
    // (peer-mapped memory over NVLink); 0 rows = no such neighbour
    ptrdiff_t up_delta, dn_delta;
    int up0, nup, dn0, ndn;
};

template <class C>
NLSB_HD Lane<C> make_lane(const Chunk &g, int tid, const double2 *ring, double2 *yr, double2 *out, int rows, int cols,
                          int grow0, int grows, double dt)
{
    Lane<C> L;
    L.ring = ring;
    L.yr = yr;
    L.fx = tid;
    const int gx = g.c0 - C::HALO + tid;
    L.col_in = gx >= 0 && gx < cols;
    L.col_owned = L.col_in && tid >= C::HALO && tid < C::T - C::HALO;
    L.out = out + (L.col_in ? gx : 0);
    L.pitch = (size_t)cols;
    L.dlo = grow0 < 0 ? -grow0 : 0;
    const int dhi = rows < grows - grow0 ? rows : grows - grow0;
    L.dspan = dhi > L.dlo ? dhi - L.dlo : 0;
    L.half_dt = dt / 2; L.dt = dt; L.dt6 = dt / 6;
    L.up_delta = L.dn_delta = 0;
    L.up0 = L.nup = L.dn0 = L.ndn = 0;
    return L;
}

template <class C>
NLSB_HD bool row_in_domain(const Lane<C> &L, int ly)
{
    return (unsigned)(ly - L.dlo) < (unsigned)L.dspan;
}

NLSB_HD void store_if(double2 *p, double2 v, bool put)
{
#if defined(__CUDA_ARCH__)
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.v2.f64 [%0], {%1, %2};\n\t}"
                 ::"l"(p), "d"(v.x), "d"(v.y), "r"((int)put) : "memory");
#else
    if (put) *p = v;
#endif
}

// Cross stencil of one node WITHOUT its centre tap (rhs_point_c folds wx[K] into the pointwise part): y taps from the
// register window `w` (period PER, centre index CI), x taps from `xn` (xn[K], the centre, is not read).  Order of
// the chain: the K rows above, the 2K x taps, the K rows below (fused_2d.cu).
template <class C, int PER>
NLSB_HD void cross_stencil(const double2 (&w)[PER], int ci, const double2 (&xn)[C::NW], const double (&wx)[C::NW],
                           const double (&wy)[C::NW], double &lr, double &li)
{
    constexpr int K = C::K;
    lr = wy[0] * w[(ci - K + PER) % PER].x;
    li = wy[0] * w[(ci - K + PER) % PER].y;
#pragma unroll
    for (int d = -K + 1; d < 0; ++d) {
        lr = fma(wy[d + K], w[(ci + d + PER) % PER].x, lr);
        li = fma(wy[d + K], w[(ci + d + PER) % PER].y, li);
    }
#pragma unroll
    for (int tp = 0; tp < C::NW; ++tp) {
        if (tp == K) continue;
        lr = fma(wx[tp], xn[tp].x, lr);
        li = fma(wx[tp], xn[tp].y, li);
    }
#pragma unroll
    for (int d = 1; d <= K; ++d) {
        lr = fma(wy[d + K], w[(ci + d) % PER].x, lr);
        li = fma(wy[d + K], w[(ci + d) % PER].y, li);
    }
}

// Values of the window registers before the first iteration: psi rows jstart-K .. jstart+K-1 from TMA batch 0,
// everything else zero (finite garbage never reaches a node that is kept: see the validity argument in
// DESIGN.md 3.5).
template <class C>
NLSB_HD void march_begin(State<C> &s, const Lane<C> &L, const Chunk &g)
{
    constexpr int K = C::K, U = C::U;
#pragma unroll
    for (int i = 0; i < U; ++i) {
        s.psi[i] = make_double2(0.0, 0.0);
        s.acc[i] = make_double2(0.0, 0.0);
        s.cp[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < C::NW; ++i) s.y2[i] = s.y3[i] = s.y4[i] = make_double2(0.0, 0.0);
#pragma unroll
    for (int d = -K; d < K; ++d) s.psi[(d + 3 * K) % U] = L.ring[(d + K) * C::T + L.fx];   // it = 0: ring slot = rel
    s.onext = L.out + (ptrdiff_t)(g.jstart - 3 * K) * (ptrdiff_t)L.pitch;
}

// Row `rel` (counted from the first row of TMA batch 0) of the psi ring lives in slot rel mod 2U.  An unrolled
// body starts at it = itb (a multiple of U): with rh = ring + (itb mod 2U) T and ro = the other half, the row
// rel = itb + srel (srel a compile-time constant in [0, 2U)) is at a constant offset from rh or ro.
template <class C, class E>
NLSB_HD const E *ring_row(const E *rh, const E *ro, int srel)
{
    return srel < C::U ? rh + srel * C::T : ro + (srel - C::U) * C::T;
}

// One iteration of the march.  `ph` must equal it mod U, rh / ro are the halves of the psi ring and ph_ / po_ of
// the pumping ring (already offset to this thread's column) as ring_row wants them; in the kernel `ph` is a
// compile-time constant of the unrolled copy and the halves swap once per unrolled body.
//
// DIAG: stage 1 holds k1 = H(psi) of the state ENTERING the step, so the scalar diagnostics of that state (chemical
// potential sums, damping integral, particle number, peak density / reservoir: diag_acc.cuh) are accumulated there
// for the nodes this CTA produces -- no extra pass over the field, no stencil evaluated twice.
//
// MASKED = false: the caller guarantees that every ROW this iteration evaluates for later use lies inside the domain
// (rows j - 3K .. j of an interior block of iterations), so the "zero outside the domain" selects vanish: per
// iteration 12 FSEL, the register moves that feed their results to the stage-ring stores and the row comparisons --
// about a third of the loop's non-FP64 instructions.  COLUMNS outside the domain (the left halo of the first strip,
// the tail of the last) need no select either: a thread that owns one simply never publishes -- its stage-ring
// entries keep the zeros the rings are initialised with, which is what its in-domain neighbours must read, and
// whatever it computes for itself is never stored (predicated STS: no extra instruction).  On B200 an FP64
// instruction holds the scheduler's issue port for two cycles and EVERY other instruction for one
// (tools/micro/fp64_issue.cu), so those instructions are paid for in FP64 throughput.
//
// PEER: see Lane -- the new psi rows a neighbouring rank needs go straight into its halo rows from this kernel's
// store stage (the halo exchange is the epilogue of the step, not a copy afterwards).
template <class C, bool DIAG = false, bool MASKED = true, bool PEER = false>
NLSB_HD void march_iter(State<C> &s, const Lane<C> &L, const Chunk &g, const RhsCoeffs &c, const double (&wx)[C::NW],
                        const double (&wy)[C::NW], int it, int ph, const double2 *rh, const double2 *ro, const double *ph_,
                        const double *po_, DiagAcc *diag = nullptr, double area = 0.0)
{
    constexpr int K = C::K, U = C::U, NW = C::NW, YP = C::YP;
    const int j = g.jstart + it;
    const int q = ph % NW;                // stage-ring row written in this iteration
    const int qr = (ph + NW - K) % NW;    // stage-ring row written K iterations ago

    // ---- loads ---------------------------------------------------------------------------------------
    s.psi[(ph + 4 * K) % U] = *ring_row<C>(rh, ro, ph + 2 * K);                        // psi(j + K)
    double2 x1[NW], x2[NW], x3[NW], x4[NW];
    {
        const double2 *row = ring_row<C>(rh, ro, ph + K);                              // psi(j)
        const double2 *r2 = L.yr + (0 * C::YS + qr) * YP + K + L.fx;
        const double2 *r3 = L.yr + (1 * C::YS + qr) * YP + K + L.fx;
        const double2 *r4 = L.yr + (2 * C::YS + qr) * YP + K + L.fx;
#pragma unroll
        for (int tp = -K; tp <= K; ++tp) {
            if (tp == 0) continue;
            x1[tp + K] = row[tp];      // edge threads read up to K elements outside their row (see Cfg)
            x2[tp + K] = r2[tp];
            x3[tp + K] = r3[tp];
            x4[tp + K] = r4[tp];
        }
    }
    const bool in1 = !MASKED || (L.col_in && row_in_domain(L, j));
    const bool in2 = !MASKED || (L.col_in && row_in_domain(L, j - K));
    const bool in3 = !MASKED || (L.col_in && row_in_domain(L, j - 2 * K));

    // ---- stage 1, row j -------------------------------------------------------------------------------
    {
        const int ci = (ph + 3 * K) % U;
        s.cp[ci] = c.c12 * *ring_row<C>(ph_, po_, ph + K);    // c12 * P(j), rounded once (nls.f90:580 association)
        const double2 u = s.psi[ci];
        double lr, li;
        cross_stencil<C, U>(s.psi, ci, x1, wx, wy, lr, li);
        const double2 k = rhs_point_c(c, s.cp[ci], u, wx[K], lr, li);
        if (DIAG) {
            if (L.col_owned && (unsigned)(j - g.r0) < (unsigned)(g.r1 - g.r0) && (!MASKED || row_in_domain(L, j)))
                diag_accumulate(*diag, c, s.cp[ci], u, k, 1.0, area);
        }
        double2 y;
        y.x = in1 ? fma(k.x, L.half_dt, u.x) : 0.0;
        y.y = in1 ? fma(k.y, L.half_dt, u.y) : 0.0;
        s.acc[ci] = k;
        s.y2[(ph + 2 * K) % NW] = y;
        if (MASKED || L.col_in) L.yr[(0 * C::YS + q) * YP + K + L.fx] = y;     // see MASKED below: outside columns stay zero
    }
    // ---- stage 2, row j - K ---------------------------------------------------------------------------
    {
        const int ci = (ph + K) % NW, cu = (ph + 2 * K) % U;
        const double2 u = s.y2[ci];
        double lr, li;
        cross_stencil<C, NW>(s.y2, ci, x2, wx, wy, lr, li);
        const double2 k = rhs_point_c(c, s.cp[cu], u, wx[K], lr, li);
        double2 y;
        y.x = in2 ? fma(k.x, L.half_dt, s.psi[cu].x) : 0.0;
        y.y = in2 ? fma(k.y, L.half_dt, s.psi[cu].y) : 0.0;
        s.acc[cu].x = fma(2.0, k.x, s.acc[cu].x);
        s.acc[cu].y = fma(2.0, k.y, s.acc[cu].y);
        s.y3[(ph + 2 * K) % NW] = y;
        if (MASKED || L.col_in) L.yr[(1 * C::YS + q) * YP + K + L.fx] = y;     // see MASKED below: outside columns stay zero
    }
    // ---- stage 3, row j - 2K --------------------------------------------------------------------------
    {
        const int ci = (ph + K) % NW, cu = (ph + K) % U;
        const double2 u = s.y3[ci];
        double lr, li;
        cross_stencil<C, NW>(s.y3, ci, x3, wx, wy, lr, li);
        const double2 k = rhs_point_c(c, s.cp[cu], u, wx[K], lr, li);
        double2 y;
        y.x = in3 ? fma(k.x, L.dt, s.psi[cu].x) : 0.0;
        y.y = in3 ? fma(k.y, L.dt, s.psi[cu].y) : 0.0;
        s.acc[cu].x = fma(2.0, k.x, s.acc[cu].x);
        s.acc[cu].y = fma(2.0, k.y, s.acc[cu].y);
        s.y4[(ph + 2 * K) % NW] = y;
        if (MASKED || L.col_in) L.yr[(2 * C::YS + q) * YP + K + L.fx] = y;     // see MASKED below: outside columns stay zero
    }
    // ---- stage 4, row j - 3K: the new psi ----------------------------------------------------------------
    {
        const int ci = (ph + K) % NW, cu = ph % U;
        const int r = j - 3 * K;
        const double2 u = s.y4[ci];
        double lr, li;
        cross_stencil<C, NW>(s.y4, ci, x4, wx, wy, lr, li);
        const double2 k = rhs_point_c(c, s.cp[cu], u, wx[K], lr, li);
        double2 v;
        v.x = fma(s.acc[cu].x + k.x, L.dt6, s.psi[cu].x);
        v.y = fma(s.acc[cu].y + k.y, L.dt6, s.psi[cu].y);
        const bool put = L.col_owned && (unsigned)(r - g.r0) < (unsigned)(g.r1 - g.r0) && (!MASKED || row_in_domain(L, r));
        store_if(s.onext, v, put);
        if (PEER) {
            store_if(reinterpret_cast<double2 *>(reinterpret_cast<char *>(s.onext) + L.up_delta), v,
                     put && (unsigned)(r - L.up0) < (unsigned)L.nup);
            store_if(reinterpret_cast<double2 *>(reinterpret_cast<char *>(s.onext) + L.dn_delta), v,
                     put && (unsigned)(r - L.dn0) < (unsigned)L.ndn);
        }
        s.onext += L.pitch;
    }
}

}  // namespace stream2d
}  // namespace nlsb
