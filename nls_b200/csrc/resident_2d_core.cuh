// resident_2d_core.cuh -- per-thread body of the RESIDENT 2D solver: the whole RK4 time loop of a small grid in
// one launch, the field held in registers, neighbouring CTAs exchanging their edge nodes once per RK stage.
//
// Reference semantics: runge_kutta_2d (nls.f90:873-901) = `iters` times four hamiltonian_2d evaluations
// (:841-870; cross stencil of make_laplacian_2d :297-385; reservoir :829-839) and the RK4 update.
//
// Why: a grid like BASELINE config 2 (512 x 512) is 1771 nodes per SM.  One launch per step (fused_2d.cu)
// spends most of a step on launch latency, tile fill and the redundant halo ring (1.43x) -- about 8 us per step.
// Here the grid is cut into one PATCH per CTA (<= 128 columns x <= 15 rows; 4 x 37 = 148 patches for 512^2),
// every thread keeps the RK accumulator and the current stage input of its 5 nodes in registers (psi and c12*P in
// thread-private shared-memory slots) for all steps, and nothing is recomputed.  Per stage a CTA
//   A  forms the coefficients a, b of its nodes (division, reservoir; needs no neighbour),
//   H  copies the K-deep halo of the stage input from its <= 4 neighbours' mailboxes into its frame,
//   B  evaluates the stencil + stage algebra, writes the next stage input into the other frame and its own
//      edge nodes into its mailbox.
// Phase A overlaps the latency of the exchange.  One __syncthreads per stage (between H and B).
//
// Mailboxes (global memory, L2 resident): 16-byte packets {lo32(value), seq, hi32(value), seq}; a reader spins on
// the packet itself until both sequence words equal the stage counter g it waits for -- no separate flag, no
// fence (NCCL's LL protocol).  Two parities: a patch can publish stage g+2 only after it has consumed its
// neighbours' stage g+1, which they publish only after consuming this patch's stage g.  A reader that would see
// any other sequence number than g spins until the kernel's timeout traps: protocol errors are loud.
//
// The arithmetic of a node (FMA chain of the stencil, rhs, stage algebra) is the one of fused_2d.cu /
// stream_2d.cu: the three kernels produce bit-identical fields.
//
// Compiled for the device (resident_2d.cu) and for the host (tests/emu/resident_emu.cu: CTAs run as coroutines
// in random order, which exercises the mailbox protocol without a GPU).
#pragma once

#include "device_math.cuh"

#include <cstdint>
#include <cstring>

#ifndef NLSB_HD
#define NLSB_HD __host__ __device__ __forceinline__
#endif

namespace nlsb {
namespace resident2d {

struct alignas(16) Packet {
    uint32_t lo, seq0, hi, seq1;
};

template <int K_>
struct Cfg {
    static constexpr int K = K_, NW = 2 * K_ + 1;
    static constexpr int TX = 128;                 // patch columns = threads per row group
    static constexpr int RT = 5;                   // rows a thread owns
    static constexpr int RG = 3;                   // row groups
    static constexpr int T = TX * RG;              // threads per CTA
    static constexpr int PHMAX = RT * RG;          // patch rows
    static constexpr int FP = TX + 2 * K_;         // frame pitch (K halo columns each side)
    static constexpr int FH = PHMAX + 2 * K_;
    static constexpr int FRAME = FP * FH;          // cells of one frame; two frames ping-pong
    // mailbox of one patch and parity, in cells (a cell = 2 packets: re, im)
    static constexpr int MB_TOP = 0;                           // rows 0 .. K-1            [K][TX]
    static constexpr int MB_BOTTOM = K_ * TX;                  // rows ph-K .. ph-1        [K][TX]
    static constexpr int MB_LEFT = 2 * K_ * TX;                // columns 0 .. K-1         [PHMAX][K]
    static constexpr int MB_RIGHT = 2 * K_ * TX + PHMAX * K_;  // columns pw-K .. pw-1     [PHMAX][K]
    static constexpr int MB_CELLS = 2 * K_ * TX + 2 * PHMAX * K_;
    static constexpr int NCELL = (MB_CELLS + T - 1) / T;       // halo cells a thread copies per stage
    static constexpr int PLANE = PHMAX * TX;       // cells of the psi / c12*P planes (patch interior, thread-private slots)
    static constexpr size_t SMEM = sizeof(double2) * (2 * FRAME + PLANE) + sizeof(double) * PLANE;
    static constexpr size_t mailbox_bytes(long long patches) { return sizeof(Packet) * 2 * MB_CELLS * 2 * (size_t)patches; }
};

// How a grid is cut into patches (host side; the same numbers reach the kernel through Layout).
struct Layout {
    int rows, cols, batch;
    int npx, npy;        // patches per member
    int pw, ph;          // nominal patch size (the last column / row of patches may be smaller, never below K)
};

template <class C>
inline bool make_layout(int batch, int rows, int cols, long long capacity, Layout *out)
{
    Layout l;
    l.rows = rows; l.cols = cols; l.batch = batch;
    l.npx = (cols + C::TX - 1) / C::TX;
    l.pw = (cols + l.npx - 1) / l.npx;
    l.npx = (cols + l.pw - 1) / l.pw;
    if (cols - (l.npx - 1) * l.pw < C::K) return false;
    const long long per_member = capacity / batch;
    long long npy_max = per_member / l.npx;
    if (npy_max < 1) return false;
    const int thin = rows / 4 > 0 ? rows / 4 : 1;             // patches of at least 4 rows
    if (npy_max > thin) npy_max = thin;
    for (int ph = (int)((rows + npy_max - 1) / npy_max); ph <= C::PHMAX; ++ph) {
        const int npy = (rows + ph - 1) / ph;
        if (ph < C::K || rows - (npy - 1) * ph < C::K) continue;
        l.ph = ph; l.npy = npy;
        *out = l;
        return true;
    }
    return false;
}

// One CTA's patch (uniform over its threads).
struct Patch {
    int member, px, py;
    int pw, ph;          // actual extent of this patch
    int col0, row0;      // global coordinates of its first node
    int id;              // index of its mailbox
    int up, down, left, right;   // mailbox index of the neighbours, -1 at the domain edge
};

NLSB_HD Patch make_patch(const Layout &l, int cta)
{
    Patch p;
    const int per = l.npx * l.npy;
    p.member = cta / per;
    const int r = cta % per;
    p.py = r / l.npx; p.px = r % l.npx;
    p.col0 = p.px * l.pw; p.row0 = p.py * l.ph;
    p.pw = l.cols - p.col0 < l.pw ? l.cols - p.col0 : l.pw;
    p.ph = l.rows - p.row0 < l.ph ? l.rows - p.row0 : l.ph;
    p.id = cta;
    p.up = p.py > 0 ? cta - l.npx : -1;
    p.down = p.py < l.npy - 1 ? cta + l.npx : -1;
    p.left = p.px > 0 ? cta - 1 : -1;
    p.right = p.px < l.npx - 1 ? cta + 1 : -1;
    return p;
}

template <class C>
NLSB_HD int frame_index(int row, int col) { return (row + C::K) * C::FP + col + C::K; }

// What a thread keeps in registers for the whole time loop: the RK accumulator and the current stage input of
// its nodes (plus the coefficients a, b between the two phases of a stage).  psi and c12*P, read once per stage,
// sit in thread-private slots of shared memory (planes indexed row * TX + x).
template <class C>
struct State {
    double2 acc[C::RT], y[C::RT];
    double a[C::RT], b[C::RT];
};

// Halo cell `c` (0 <= c < MB_CELLS) of a patch: where it lands in the frame and which cell of which neighbour's
// mailbox holds it.  Returns false when the cell does not exist for this patch (domain edge, narrow patch).
template <class C>
NLSB_HD bool halo_cell(const Patch &p, int c, int &dst, int &src_patch, int &src_cell)
{
    constexpr int K = C::K, TX = C::TX;
    if (c < C::MB_BOTTOM) {                              // my top halo <- the bottom rows of the patch above
        const int hr = c / TX, hx = c % TX;
        dst = frame_index<C>(-K + hr, hx);
        src_patch = p.up; src_cell = C::MB_BOTTOM + hr * TX + hx;
        return p.up >= 0 && hx < p.pw;
    }
    if (c < C::MB_LEFT) {                                // my bottom halo <- the top rows of the patch below
        const int hr = (c - C::MB_BOTTOM) / TX, hx = (c - C::MB_BOTTOM) % TX;
        dst = frame_index<C>(p.ph + hr, hx);
        src_patch = p.down; src_cell = C::MB_TOP + hr * TX + hx;
        return p.down >= 0 && hx < p.pw;
    }
    if (c < C::MB_RIGHT) {                               // my left halo <- the right columns of the patch to the left
        const int r = (c - C::MB_LEFT) / K, hc = (c - C::MB_LEFT) % K;
        dst = frame_index<C>(r, -K + hc);
        src_patch = p.left; src_cell = C::MB_RIGHT + r * K + hc;
        return p.left >= 0 && r < p.ph;
    }
    const int r = (c - C::MB_RIGHT) / K, hc = (c - C::MB_RIGHT) % K;
    dst = frame_index<C>(r, p.pw + hc);
    src_patch = p.right; src_cell = C::MB_LEFT + r * K + hc;
    return p.right >= 0 && r < p.ph;
}

NLSB_HD unsigned long long bits_of(double v)
{
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b;
    memcpy(&b, &v, sizeof(b));
    return b;
#endif
}

NLSB_HD double double_of(unsigned long long b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, sizeof(v));
    return v;
#endif
}

NLSB_HD void packet_store(Packet *dst, double v, uint32_t seq)
{
    const unsigned long long bits = bits_of(v);
    const uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
#if defined(__CUDA_ARCH__)
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
#else
    dst->lo = lo; dst->seq0 = seq; dst->hi = hi; dst->seq1 = seq;
#endif
}

// One 16-byte read of a packet (straight from L2 on the device); packet_ok tells whether it carries sequence
// number `seq`, packet_value extracts the double.
NLSB_HD Packet packet_read(const Packet *src)
{
    Packet r;
#if defined(__CUDA_ARCH__)
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.lo), "=r"(r.seq0), "=r"(r.hi), "=r"(r.seq1) : "l"(src) : "memory");
#else
    r = *src;
#endif
    return r;
}
NLSB_HD bool packet_ok(const Packet &r, uint32_t seq) { return r.seq0 == seq && r.seq1 == seq; }
NLSB_HD double packet_value(const Packet &r) { return double_of(((unsigned long long)r.hi << 32) | r.lo); }

// true (and the value) when the packet carries sequence number `seq`
NLSB_HD bool packet_load(const Packet *src, uint32_t seq, double &v)
{
    const Packet r = packet_read(src);
    v = packet_value(r);
    return packet_ok(r, seq);
}

// Mailbox cell `cell` of patch `patch`, parity `par`: two packets (re, im).
template <class C>
NLSB_HD Packet *mailbox_cell(Packet *mail, int patch, int par, int cell)
{
    return mail + (((size_t)patch * 2 + par) * C::MB_CELLS + cell) * 2;
}

// Phase A: the coefficients a, b of the thread's nodes from the stage input y (registers only).
template <class C>
NLSB_HD void phase_a(State<C> &s, const RhsCoeffs &c, const double *cp_plane, int x, int r0, double w0)
{
#pragma unroll
    for (int i = 0; i < C::RT; ++i) rhs_abm(c, cp_plane[(r0 + i) * C::TX + x], s.y[i], w0, s.a[i], s.b[i]);   // b holds b - w0
}

// Phase B of stage S (1..4) for the thread owning column x, rows r0 .. r0 + RT - 1 of patch p.
// cur / nxt: frames holding the stage input (interior + halo) / receiving the next stage input.
// seq_next: sequence number the edge nodes are published under; out != nullptr: last stage of the last step,
// the new psi goes to the output array instead (nothing is published).
template <class C, int S>
NLSB_HD void phase_b(State<C> &s, const Patch &p, int x, int r0, double2 *psi_plane, const double2 *cur, double2 *nxt, Packet *mail,
                     uint32_t seq_next, double2 *out, size_t out_pitch, const double (&wx)[C::NW],
                     const double (&wy)[C::NW], double half_dt, double dt, double dt6)
{
    constexpr int K = C::K, RT = C::RT;
    // the column of the stage input this thread needs: rows r0 - K .. r0 + RT - 1 + K; its own valid nodes
    // from registers, everything else (halo rows, rows of the other row group) from the frame
    double2 ext[RT + 2 * K];
#pragma unroll
    for (int e = 0; e < RT + 2 * K; ++e) {
        const int i = e - K, row = r0 + i;
        if (i >= 0 && i < RT) {
            ext[e] = s.y[i];
            if (row >= p.ph) ext[e] = cur[frame_index<C>(row, x)];
        } else {
            ext[e] = cur[frame_index<C>(row, x)];
        }
    }
    const int par = (int)(seq_next & 1u);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int row = r0 + i;
        const double2 *line = cur + frame_index<C>(row, x);
        double lr = wy[0] * ext[i].x, li = wy[0] * ext[i].y;             // rows above: ext[i] is row - K
#pragma unroll
        for (int d = -K + 1; d < 0; ++d) {
            lr = fma(wy[d + K], ext[i + K + d].x, lr);
            li = fma(wy[d + K], ext[i + K + d].y, li);
        }
#pragma unroll
        for (int tp = -K; tp <= K; ++tp) {
            if (tp == 0) continue;          // the centre tap is folded into b (rhs_abm)
            const double2 v = line[tp];
            lr = fma(wx[tp + K], v.x, lr);
            li = fma(wx[tp + K], v.y, li);
        }
#pragma unroll
        for (int d = 1; d <= K; ++d) {
            lr = fma(wy[d + K], ext[i + K + d].x, lr);
            li = fma(wy[d + K], ext[i + K + d].y, li);
        }
        const double2 u = ext[i + K];
        const double2 k = rhs_apply(s.a[i], s.b[i], u, lr, li);
        const bool inside = row < p.ph && x < p.pw;
        const double2 psi = psi_plane[row * C::TX + x];
        double2 y;
        if (S < 4) {
            const double cy = (S == 3) ? dt : half_dt;
            y.x = inside ? fma(k.x, cy, psi.x) : 0.0;
            y.y = inside ? fma(k.y, cy, psi.y) : 0.0;
            if (S == 1) {
                s.acc[i] = k;
            } else {
                s.acc[i].x = fma(2.0, k.x, s.acc[i].x);
                s.acc[i].y = fma(2.0, k.y, s.acc[i].y);
            }
        } else {
            y.x = inside ? fma(s.acc[i].x + k.x, dt6, psi.x) : 0.0;
            y.y = inside ? fma(s.acc[i].y + k.y, dt6, psi.y) : 0.0;
            psi_plane[row * C::TX + x] = y;
        }
        s.y[i] = y;
        if (!inside) continue;
        if (out) {
            out[(size_t)(p.row0 + row) * out_pitch + p.col0 + x] = y;
            continue;
        }
        nxt[frame_index<C>(row, x)] = y;
        if (row < K) {
            Packet *q = mailbox_cell<C>(mail, p.id, par, C::MB_TOP + row * C::TX + x);
            packet_store(q, y.x, seq_next); packet_store(q + 1, y.y, seq_next);
        }
        if (row >= p.ph - K) {
            Packet *q = mailbox_cell<C>(mail, p.id, par, C::MB_BOTTOM + (row - (p.ph - K)) * C::TX + x);
            packet_store(q, y.x, seq_next); packet_store(q + 1, y.y, seq_next);
        }
        if (x < K) {
            Packet *q = mailbox_cell<C>(mail, p.id, par, C::MB_LEFT + row * K + x);
            packet_store(q, y.x, seq_next); packet_store(q + 1, y.y, seq_next);
        }
        if (x >= p.pw - K) {
            Packet *q = mailbox_cell<C>(mail, p.id, par, C::MB_RIGHT + row * K + (x - (p.pw - K)));
            packet_store(q, y.x, seq_next); packet_store(q + 1, y.y, seq_next);
        }
    }
}

}  // namespace resident2d
}  // namespace nlsb
