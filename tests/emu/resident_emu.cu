// resident_emu.cu -- HOST emulation of rk4_resident_2d_kernel (nls_b200/csrc/resident_2d.cu): the same per-thread
// body (resident_2d_core.cuh), CTAs run as coroutines in a seeded random order.  A CTA may run its next stage only
// when every halo packet it needs carries exactly the sequence number it waits for, as on the device; a packet
// overwritten too early (protocol error) is reported, and so is a deadlock.  Test infrastructure: lets the CPU
// suite check patch layout, halo mapping and the mailbox protocol without a GPU.
#include "../../nls_b200/csrc/resident_2d_core.cuh"

#include <cstdio>
#include <random>
#include <vector>

using namespace nlsb;
using namespace nlsb::resident2d;

template <class C>
struct Cta {
    Patch p;
    std::vector<double2> frames;        // 2 * FRAME
    std::vector<double2> psi;           // PLANE
    std::vector<double> cp;             // PLANE
    std::vector<State<C>> st;           // per thread
    uint32_t g;                         // sequence number of the next stage input
    int stage;                          // stages completed
};

template <class C>
static int run(int batch, int rows, int cols, long long capacity, int steps, unsigned seed, const double2 *in,
               const double *P, const double *coeffs, const double *wx_, const double *wy_, double dt, double2 *out,
               int *layout_out)
{
    Layout l;
    if (!make_layout<C>(batch, rows, cols, capacity, &l)) return 1;
    if (layout_out) { layout_out[0] = l.npx; layout_out[1] = l.npy; layout_out[2] = l.pw; layout_out[3] = l.ph; }
    double wx[C::NW], wy[C::NW];
    for (int i = 0; i < C::NW; ++i) { wx[i] = wx_[i]; wy[i] = wy_[i]; }
    const int ncta = l.npx * l.npy * l.batch;
    if (ncta > capacity) return 2;
    const size_t plane = (size_t)rows * cols;
    std::vector<Packet> mail(C::mailbox_bytes(ncta) / sizeof(Packet), Packet{0, 0, 0, 0});
    std::vector<Cta<C>> ctas(ncta);
    std::vector<double2> result((size_t)batch * plane);
    for (int id = 0; id < ncta; ++id) {
        Cta<C> &t = ctas[id];
        t.p = make_patch(l, id);
        t.frames.assign(2 * C::FRAME, make_double2(0.0, 0.0));
        t.st.resize(C::T);
        t.psi.assign(C::PLANE, make_double2(0.0, 0.0));
        t.cp.assign(C::PLANE, 0.0);
        t.g = 0; t.stage = 0;
        const double2 *min = in + t.p.member * plane;
        const double *Pm = P + t.p.member * plane;
        const RhsCoeffs c = rhs_coeffs_from(coeffs + (size_t)t.p.member * 23);
        for (int i = 0; i < C::FRAME; ++i) {
            const int fr = i / C::FP - C::K, fc = i % C::FP - C::K, gr = t.p.row0 + fr, gc = t.p.col0 + fc;
            if (gr >= 0 && gr < rows && gc >= 0 && gc < cols) t.frames[i] = min[(size_t)gr * cols + gc];
        }
        for (int tid = 0; tid < C::T; ++tid) {
            const int x = tid % C::TX, r0 = (tid / C::TX) * C::RT;
            State<C> &s = t.st[tid];
            for (int i = 0; i < C::RT; ++i) {
                const int row = r0 + i;
                const bool inside = row < t.p.ph && x < t.p.pw;
                const size_t g = (size_t)(t.p.row0 + row) * cols + t.p.col0 + x;
                s.y[i] = inside ? min[g] : make_double2(0.0, 0.0);
                t.psi[row * C::TX + x] = s.y[i];
                t.cp[row * C::TX + x] = inside ? c.c12 * Pm[g] : 0.0;
                s.acc[i] = make_double2(0.0, 0.0);
            }
        }
    }
    std::mt19937 rng(seed);
    const int total = 4 * steps;
    long long remaining = (long long)ncta * total, idle = 0;
    while (remaining > 0) {
        const int id = (int)(rng() % ncta);
        Cta<C> &t = ctas[id];
        if (t.stage == total) continue;
        double2 *cur = t.frames.data() + (t.g & 1u) * C::FRAME, *nxt = t.frames.data() + ((t.g & 1u) ^ 1u) * C::FRAME;
        // halo: every packet must carry exactly t.g
        bool ready = true;
        if (t.stage > 0) {
            for (int cell = 0; cell < C::MB_CELLS && ready; ++cell) {
                int d, sp, sc;
                if (!halo_cell<C>(t.p, cell, d, sp, sc)) continue;
                const Packet *q = mailbox_cell<C>(mail.data(), sp, (int)(t.g & 1u), sc);
                double re, im;
                if (q->seq0 > t.g || q[1].seq0 > t.g) return 3;           // overwritten before it was consumed
                if (!packet_load(q, t.g, re) || !packet_load(q + 1, t.g, im)) ready = false;
            }
        }
        if (!ready) {
            if (++idle > 100000ll * ncta) return 4;    // deadlock
            continue;
        }
        idle = 0;
        if (t.stage > 0)
            for (int cell = 0; cell < C::MB_CELLS; ++cell) {
                int d, sp, sc;
                if (!halo_cell<C>(t.p, cell, d, sp, sc)) continue;
                const Packet *q = mailbox_cell<C>(mail.data(), sp, (int)(t.g & 1u), sc);
                double re, im;
                packet_load(q, t.g, re); packet_load(q + 1, t.g, im);
                cur[d] = make_double2(re, im);
            }
        const RhsCoeffs c = rhs_coeffs_from(coeffs + (size_t)t.p.member * 23);
        const int S = t.stage % 4 + 1;
        const bool last = t.stage == total - 1;
        double2 *o = last ? result.data() + t.p.member * plane : nullptr;
        for (int tid = 0; tid < C::T; ++tid) phase_a<C>(t.st[tid], c, t.cp.data(), tid % C::TX, (tid / C::TX) * C::RT, wx[C::K]);
        for (int tid = 0; tid < C::T; ++tid) {
            const int x = tid % C::TX, r0 = (tid / C::TX) * C::RT;
            State<C> &s = t.st[tid];
            switch (S) {
            case 1: phase_b<C, 1>(s, t.p, x, r0, t.psi.data(), cur, nxt, mail.data(), t.g + 1, o, cols, wx, wy, dt / 2, dt, dt / 6); break;
            case 2: phase_b<C, 2>(s, t.p, x, r0, t.psi.data(), cur, nxt, mail.data(), t.g + 1, o, cols, wx, wy, dt / 2, dt, dt / 6); break;
            case 3: phase_b<C, 3>(s, t.p, x, r0, t.psi.data(), cur, nxt, mail.data(), t.g + 1, o, cols, wx, wy, dt / 2, dt, dt / 6); break;
            default: phase_b<C, 4>(s, t.p, x, r0, t.psi.data(), cur, nxt, mail.data(), t.g + 1, o, cols, wx, wy, dt / 2, dt, dt / 6); break;
            }
        }
        t.g += 1; t.stage += 1; --remaining;
    }
    for (size_t i = 0; i < result.size(); ++i) out[i] = result[i];
    return 0;
}

extern "C" int emu_resident(int order, int batch, int rows, int cols, long long capacity, int steps, unsigned seed,
                            const double *in, const double *P, const double *coeffs, const double *wx, const double *wy,
                            double dt, double *out, int *layout_out)
{
    const double2 *i2 = reinterpret_cast<const double2 *>(in);
    double2 *o2 = reinterpret_cast<double2 *>(out);
    switch (order) {
    case 3: return run<Cfg<1>>(batch, rows, cols, capacity, steps, seed, i2, P, coeffs, wx, wy, dt, o2, layout_out);
    case 5: return run<Cfg<2>>(batch, rows, cols, capacity, steps, seed, i2, P, coeffs, wx, wy, dt, o2, layout_out);
    case 7: return run<Cfg<3>>(batch, rows, cols, capacity, steps, seed, i2, P, coeffs, wx, wy, dt, o2, layout_out);
    }
    return 9;
}
