// stream_emu.cu -- HOST emulation of rk4_stream_kernel (nls_b200/csrc/stream_2d.cu): the same per-thread body
// (stream_2d_core.cuh) run thread by thread, with the TMA batches and the barriers replaced by their sequential
// meaning.  Test infrastructure: it lets the CPU suite check the window / ring indexing of the kernel against
// the oracle on a machine without a GPU.  A refill is emulated at the earliest moment the kernel may issue it,
// so a batch that is overwritten while still needed shows up as a wrong result.
#include "../../nls_b200/csrc/stream_2d_core.cuh"

#include <cstdio>
#include <vector>

using namespace nlsb;
using namespace nlsb::stream2d;

template <class C>
static void fill_batch(double2 *ring, double *pring, const Chunk &g, int b, const double2 *in, const double *P, int rows,
                       int cols)
{
    for (int r = 0; r < C::RB; ++r)
        for (int x = 0; x < C::T; ++x) {
            const int ly = g.base + b * C::RB + r, lx = g.c0 - C::HALO + x;
            double2 v = make_double2(0.0, 0.0);
            double p = 0.0;
            if (ly >= 0 && ly < rows && lx >= 0 && lx < cols) {
                v = in[(size_t)ly * cols + lx];
                p = P[(size_t)ly * cols + lx];
            }
            ring[((b % C::NB) * C::RB + r) * C::T + x] = v;
            pring[((b % C::NB) * C::RB + r) * C::T + x] = p;
        }
}

template <class C>
static int run(int rows, int cols, int grow0, int grows, int out_row0, int out_row1, int chunk_rows, const double2 *in,
               const double *P, const double *coeffs, const double *wx_, const double *wy_, double dt, double2 *out)
{
    double wx[C::NW], wy[C::NW];
    for (int i = 0; i < C::NW; ++i) { wx[i] = wx_[i]; wy[i] = wy_[i]; }
    const RhsCoeffs c = rhs_coeffs_from(coeffs);
    const int strips = (cols + C::W - 1) / C::W;
    const int chunks = (out_row1 - out_row0 + chunk_rows - 1) / chunk_rows;
    for (int chunk = 0; chunk < chunks; ++chunk)
        for (int strip = 0; strip < strips; ++strip) {
            const Chunk g = make_chunk<C>(strip, chunk, chunk_rows, out_row0, out_row1);
            // K pad elements each side: edge threads read x neighbours outside their row (as in the kernel's layout)
            std::vector<double2> ring_store((size_t)C::RING * C::T + 2 * C::K, make_double2(1e300, 1e300));
            double2 *ring = ring_store.data() + C::K;
            std::vector<double> pring((size_t)C::RING * C::T, 1e300);
            std::vector<double2> yr((size_t)3 * C::YS * C::YP, make_double2(0.0, 0.0));
            std::vector<State<C>> st(C::T);
            std::vector<Lane<C>> lane(C::T);
            for (int b = 0; b < C::NB && b < g.nbatches; ++b) fill_batch<C>(ring, pring.data(), g, b, in, P, rows, cols);
            const int nact = active_threads<C>(g, cols);      // the warps right of the domain leave (as in the kernel)
            for (int t = 0; t < nact; ++t) {
                lane[t] = make_lane<C>(g, t, ring, yr.data(), out, rows, cols, grow0, grows, dt);
                march_begin<C>(st[t], lane[t], g);
            }
            int highest_waited = 0;
            for (int it = 0; it < g.niter; ++it) {
                if ((it + 2 * C::K) % C::RB == 0) {
                    highest_waited = (it + 2 * C::K) / C::RB;
                    if (highest_waited >= g.nbatches) return 1;     // the kernel would wait for a batch never issued
                }
                // the kernel's choice of loop body, per block of U iterations: the domain masks only when one of the
                // block's rows j - 3K .. j lies outside the domain (stream_2d.cu)
                const int j0 = g.jstart + it / C::U * C::U;
                const bool masked = j0 - C::SKEW < lane[0].dlo || j0 + C::U > lane[0].dlo + lane[0].dspan;
                for (int t = 0; t < nact; ++t) {
                    const int half = (it / C::U) & 1;
                    const double2 *rh = ring + half * C::U * C::T + t, *ro = ring + (half ^ 1) * C::U * C::T + t;
                    const double *ph = pring.data() + half * C::U * C::T + t, *po = pring.data() + (half ^ 1) * C::U * C::T + t;
                    if (masked)
                        march_iter<C, false, true>(st[t], lane[t], g, c, wx, wy, it, it % C::U, rh, ro, ph, po);
                    else
                        march_iter<C, false, false>(st[t], lane[t], g, c, wx, wy, it, it % C::U, rh, ro, ph, po);
                }
                if ((it + C::K + 1) % C::RB == 0) {
                    const int nb = (it + C::K + 1) / C::RB - 1 + C::NB;
                    if (nb < g.nbatches) fill_batch<C>(ring, pring.data(), g, nb, in, P, rows, cols);
                }
            }
        }
    return 0;
}

extern "C" int emu_stream_step(int order, int threads, int rows, int cols, int grow0, int grows, int out_row0,
                               int out_row1, int chunk_rows, const double *in, const double *P, const double *coeffs,
                               const double *wx, const double *wy, double dt, double *out)
{
    const double2 *i2 = reinterpret_cast<const double2 *>(in);
    double2 *o2 = reinterpret_cast<double2 *>(out);
#define EMU_CASE(K, T)                                                                                        \
    if (order == 2 * K + 1 && threads == T) {                                                                         \
        using C = Cfg<K, T>;                                                                                  \
        if (chunk_rows <= 0) chunk_rows = C::chunk_rows(140);                                                         \
        if ((chunk_rows + 6 * K) % C::U) return 3;                                                                    \
        return run<C>(rows, cols, grow0, grows, out_row0, out_row1, chunk_rows, i2, P, coeffs, wx, wy, dt, o2);       \
    }
    EMU_CASE(1, 256)
    EMU_CASE(2, 256)
    EMU_CASE(2, 128)
    EMU_CASE(3, 192)
    EMU_CASE(1, 64)
    EMU_CASE(2, 64)
    EMU_CASE(3, 64)
    return 2;
}
