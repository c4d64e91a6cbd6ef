// fused_2d.cu -- one launch = one whole RK4 step of the 2D solver (sm_100a).
//
// Reference semantics: one iteration of runge_kutta_2d (nls.f90:892-899): four evaluations of
// hamiltonian_2d (:841-870, cross stencil of make_laplacian_2d :297-385, reservoir :829-839) and
// the update u + (k1 + 2 k2 + 2 k3 + k4) dt/6.
//
// Why: in the per-stage formulation (kernels_2d.cu) every RK stage streams psi, the stage input
// and the accumulator through L2/HBM: ~300 B per node-step.  Here a CTA loads a tile of psi with
// a 4K-node halo ONCE (40 B per node-step of compulsory traffic), runs all four stages out of
// shared memory -- recomputing the shrinking halo ring (stage s is evaluated on tile + (4-s)K) --
// and writes the tile of the new psi once.  The kernel is then bounded by the FP64 pipe and the
// shared-memory crossbar, not by HBM (DESIGN.md "fused step").
//
// Shared memory: frame B0 = psi on tile + 4K halo; B1, B2 = stage inputs on the owned region
// (tile + 3K), ping-pong; CP = c12*P on the owned region.  Every field is stored as separate re / im
// planes of doubles, so one 16-byte LDS/STS moves the same component of two x-adjacent nodes and
// consecutive lanes touch consecutive 16-byte words.  A thread works on micro-tiles of 2 (x) x MB (y)
// nodes, which cuts the stencil's shared-memory loads from 9 to 5 per node (MB = 2, order 5).
// Work list of a stage: first the tile's own micro-tiles -- one per thread, the same thread in
// every stage, so psi, c12*P and the RK accumulator of those nodes stay in registers for the whole
// step -- then the micro-tiles of the halo ring of that stage (top band, bottom band, left and
// right columns), which carry no state and are dealt round-robin.
//
// Outside the square the field is identically zero at every stage (the reference's truncated
// band matrix): out-of-domain nodes are forced to zero after each stage.  Every node's value is
// produced by the same sequence of roundings whatever tile, CTA or GPU computes it, so results
// do not depend on the tiling or on a slab decomposition.

#include "device_math.cuh"
#include "kernels.h"

#include <cuda.h>
#include <cstdint>
#include <cstring>

namespace nlsb {

namespace {

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

template <int K_, int TY_, int MB_, int THREADS_, int MINBLOCKS_>
struct FusedCfg {
    static constexpr int K = K_, TY = TY_, MB = MB_, THREADS = THREADS_, MINBLOCKS = MINBLOCKS_;
    static constexpr int TX = 32;
    static constexpr int PH = (K_ + 1) / 2;              // x-neighbour pairs on each side
    static constexpr int HXL = 4 * K_ + 2 * (K_ & 1);    // frame columns left of the tile (even)
    static constexpr int OX = HXL - 3 * K_ - (K_ & 1);   // first owned column (even, <= tile - 3K)
    static constexpr int W0 = TX + 2 * HXL;              // frame width (even)
    static constexpr int NMX = (HXL + TX + 3 * K_ - OX + 1) / 2;   // micro-tiles per row of the owned region
    static constexpr int DY = MB_ * cdiv(3 * K_, MB_);   // owned rows above the tile (multiple of MB)
    static constexpr int OY = K_;                        // first owned row
    static constexpr int FY0 = OY + DY;                  // first tile row
    static constexpr int NMY = DY / MB_ + TY_ / MB_ + cdiv(3 * K_, MB_);
    static constexpr int H0 = OY + MB_ * NMY + K_;       // frame height
    static constexpr int W1 = 2 * NMX, H1 = MB_ * NMY;   // owned region
    // plane sizes rounded up to 128 bytes: every plane is a legal TMA destination
    static constexpr int PLANE0 = (W0 * H0 + 15) / 16 * 16, PLANE1 = (W1 * H1 + 15) / 16 * 16;
    static constexpr size_t SMEM = sizeof(double) * (2 * PLANE0 + 5 * PLANE1) + 16;   // + the TMA mbarrier
    static constexpr int TMX0 = (HXL - OX) / 2;          // first tile micro-tile column / row
    static constexpr int TMY0 = DY / MB_;
    static constexpr int TMW = TX / 2, TMH = TY_ / MB_;
    static constexpr int NTILE = TMW * TMH;
    static_assert(TY_ % MB_ == 0, "tile height must be a multiple of the micro-tile height");
    static_assert(NTILE == THREADS_, "one tile micro-tile per thread");
    static_assert(OX % 2 == 0 && W0 % 2 == 0, "pairs must be 16-byte aligned");
    static_assert(OX >= 2 * PH && OX + 2 * NMX + 2 * PH <= W0, "x halo of the frame too small");
    static_assert(SMEM <= 227 * 1024, "frames do not fit in shared memory");
};

struct FusedArgs {
    int rows, cols;          // extent of the local arrays (rows may include slab halo rows)
    int grow0, grows;        // global row index of local row 0, global number of rows
    int out_row0, out_row1;  // local rows [out_row0, out_row1) are written
    int tiles_x, tiles_y;    // tile grid covering cols x (out_row1 - out_row0)
    const double2 *in;       // [batch][rows][cols]   (interleaved kernels)
    double2 *out;            // [batch][rows][cols]
    const double *pumping;   // [batch][rows][cols]
    double *pout;            // planar kernels: output planes [batch][2][rows][pitch]
    int pitch;
    const double *coeffs;    // [batch][23] (per-member coefficients) -- unused when UNIFORM
    RhsCoeffs cu;            // coefficients shared by every member (UNIFORM kernels read them from the
                             // constant bank instead of pinning 14 registers)
    double dt, half_dt, dt6;
};

__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void sts2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// Ring micro-tile `idx` (0-based, after the tile's own micro-tiles) of stage S (1-based).
template <typename C, int S>
struct StageGrid {
    static constexpr int E = (4 - S) * C::K;
    static constexpr int MX0 = (C::HXL - E - C::OX) / 2;
    static constexpr int MX1 = (C::HXL + C::TX + E - C::OX + 1) / 2;
    static constexpr int MY0 = (C::FY0 - E - C::OY) / C::MB;
    static constexpr int MY1 = (C::FY0 + C::TY + E - C::OY + C::MB - 1) / C::MB;
    static constexpr int GW = MX1 - MX0;
    static constexpr int NTOP = GW * (C::TMY0 - MY0);
    static constexpr int NBOT = GW * (MY1 - (C::TMY0 + C::TMH));
    static constexpr int LW = C::TMX0 - MX0;
    static constexpr int RW = MX1 - (C::TMX0 + C::TMW);
    static constexpr int SW = LW + RW;                  // left and right columns of one micro-row side by side
    static constexpr int NSIDE = SW * C::TMH;
    static constexpr int NRING = NTOP + NBOT + NSIDE;
    static_assert(MX0 >= 0 && MX1 <= C::NMX && MY0 >= 0 && MY1 <= C::NMY, "stage region leaves the owned region");

    __device__ static __forceinline__ void locate(int idx, int &mx, int &my)
    {
        if (idx < NTOP) {
            mx = MX0 + idx % GW;
            my = MY0 + idx / GW;
            return;
        }
        idx -= NTOP;
        if (idx < NBOT) {
            mx = MX0 + idx % GW;
            my = C::TMY0 + C::TMH + idx / GW;
            return;
        }
        idx -= NBOT;
        constexpr int sw = SW > 0 ? SW : 1;
        const int c = idx % sw;
        my = C::TMY0 + idx / sw;
        mx = c < LW ? MX0 + c : C::TMX0 + C::TMW + (c - LW);
    }
};

template <typename C>
struct Smem {
    double *base;
    // frame 0 (psi): pitch W0; frames 1, 2 and cp: owned region, pitch W1
    __device__ __forceinline__ double *re0() const { return base; }
    __device__ __forceinline__ double *im0() const { return base + C::PLANE0; }
    __device__ __forceinline__ double *re(int f) const { return base + 2 * C::PLANE0 + (2 * (f - 1)) * C::PLANE1; }
    __device__ __forceinline__ double *im(int f) const { return base + 2 * C::PLANE0 + (2 * (f - 1) + 1) * C::PLANE1; }
    __device__ __forceinline__ double *cp() const { return base + 2 * C::PLANE0 + 4 * C::PLANE1; }
};

struct TileCtx {
    int x0, y0;          // local array coordinates of frame node (0, 0)
    int cols, grow0, grows;
    int out_row0, out_row1;
    double half_dt, dt, dt6;
};

// Where the new psi goes: interleaved complex (aos) or separate planes (planar working copy).
struct OutRef {
    double2 *aos;
    double *re, *im;
    int pitch;
};

// Per-thread state of the tile micro-tile a thread owns for the whole step.
template <int MB>
struct Owned {
    double ure[MB][2], uim[MB][2], cp[MB][2];
    double yre[MB][2], yim[MB][2];   // the stage input at the owned nodes (what this thread stored last stage)
    double2 acc[MB][2];
};

// One micro-tile of stage S.  OWNED: the thread's own tile micro-tile (state in registers,
// writes the new psi in stage 4); otherwise a stateless ring micro-tile.
template <typename C, int S, bool OWNED, bool PLANAR>
__device__ __forceinline__ void micro_tile(const Smem<C> &sm, const TileCtx &t, const RhsCoeffs &c,
                                           const double (&wx)[2 * C::K + 1], const double (&wy)[2 * C::K + 1],
                                           int mx, int my, Owned<C::MB> &own, const OutRef &out)
{
    constexpr int K = C::K, PH = C::PH, MB = C::MB;
    // stage input: S=1 psi frame; S=2 frame 1; S=3 frame 2; S=4 frame 1.  Output: 1, 2, 1.
    constexpr int SRC = (S == 1) ? 0 : (S == 3) ? 2 : 1;
    constexpr int DST = (S == 2) ? 2 : 1;
    constexpr int WS = (SRC == 0) ? C::W0 : C::W1;
    const int fx = C::OX + 2 * mx, fy = C::OY + MB * my;     // frame coordinates
    const int o1 = (MB * my) * C::W1 + 2 * mx;                 // offset in owned-region planes
    const double *sre = (SRC == 0) ? sm.re0() : sm.re(SRC);
    const double *sim = (SRC == 0) ? sm.im0() : sm.im(SRC);
    const int os = (SRC == 0) ? fy * C::W0 + fx : o1;

    double lre[MB][2], lim[MB][2], cre[MB][2], cim[MB][2];
#pragma unroll
    for (int r = -K; r < MB + K; ++r) {
        const int o = os + r * WS;
        double2 vr, vi;
        if (OWNED && S > 1 && r >= 0 && r < MB) {
            // the centre pair of an owned row is what this thread wrote itself one stage ago
            vr = make_double2(own.yre[r][0], own.yre[r][1]);
            vi = make_double2(own.yim[r][0], own.yim[r][1]);
        } else {
            vr = lds2(sre + o);
            vi = lds2(sim + o);
        }
        if (r >= 0 && r < MB) {
            cre[r][0] = vr.x; cre[r][1] = vr.y;
            cim[r][0] = vi.x; cim[r][1] = vi.y;
        }
        // scatter input row r into the output rows it touches: output row j receives, in this order,
        // the taps of the K rows above it, its own row (x taps and centre), the K rows below it
#pragma unroll
        for (int j = 0; j < MB; ++j) {
            const int d = r - j;               // input row = output row + d
            if (d == 0) {
                double xr[2 * (2 * PH + 1)], xi[2 * (2 * PH + 1)];
#pragma unroll
                for (int q = -PH; q <= PH; ++q) {
                    double2 pr, pi;
                    if (q == 0) {
                        pr = vr; pi = vi;
                    } else {
                        pr = lds2(sre + o + 2 * q);
                        pi = lds2(sim + o + 2 * q);
                    }
                    xr[2 * (q + PH)] = pr.x; xr[2 * (q + PH) + 1] = pr.y;
                    xi[2 * (q + PH)] = pi.x; xi[2 * (q + PH) + 1] = pi.y;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    double ar = lre[j][i], ai = lim[j][i];   // rows j-K .. j-1 have already contributed
#pragma unroll
                    for (int tp = -K; tp <= K; ++tp) {
                        if (tp == 0) continue;        // the centre tap is folded into the pointwise part (rhs_point_c)
                        ar = fma(wx[tp + K], xr[2 * PH + i + tp], ar);
                        ai = fma(wx[tp + K], xi[2 * PH + i + tp], ai);
                    }
                    lre[j][i] = ar;
                    lim[j][i] = ai;
                }
            } else if (d >= -K && d <= K) {
                const double w = wy[d + K];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double a = i ? vr.y : vr.x, b = i ? vi.y : vi.x;
                    if (d == -K) {             // first contribution to output row j
                        lre[j][i] = w * a;
                        lim[j][i] = w * b;
                    } else {
                        lre[j][i] = fma(w, a, lre[j][i]);
                        lim[j][i] = fma(w, b, lim[j][i]);
                    }
                }
            }
        }
    }

    // pointwise part, stage algebra, domain mask
    double *dre = sm.re(DST), *dim_ = sm.im(DST);
    const int gx = t.x0 + fx;
    const bool colin0 = gx >= 0 && gx < t.cols, colin1 = gx + 1 >= 0 && gx + 1 < t.cols;
#pragma unroll
    for (int j = 0; j < MB; ++j) {
        const int ly = t.y0 + fy + j;
        const int gy = ly + t.grow0;
        const bool rowin = gy >= 0 && gy < t.grows;
        double u_re[2], u_im[2], cpv[2];
        if (OWNED) {
            if (S == 1) {
                const double2 p = lds2(sm.cp() + o1 + j * C::W1);
                own.cp[j][0] = p.x; own.cp[j][1] = p.y;
#pragma unroll
                for (int i = 0; i < 2; ++i) { own.ure[j][i] = cre[j][i]; own.uim[j][i] = cim[j][i]; }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) { u_re[i] = own.ure[j][i]; u_im[i] = own.uim[j][i]; cpv[i] = own.cp[j][i]; }
        } else {
            const double2 p = lds2(sm.cp() + o1 + j * C::W1);
            cpv[0] = p.x; cpv[1] = p.y;
            if (S == 1) {
#pragma unroll
                for (int i = 0; i < 2; ++i) { u_re[i] = cre[j][i]; u_im[i] = cim[j][i]; }
            } else {
                const int o0 = (fy + j) * C::W0 + fx;
                const double2 a = lds2(sm.re0() + o0), b = lds2(sm.im0() + o0);
                u_re[0] = a.x; u_re[1] = a.y; u_im[0] = b.x; u_im[1] = b.y;
            }
        }
        double yr[2], yi[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double2 k = rhs_point_c(c, cpv[i], make_double2(cre[j][i], cim[j][i]), wx[K], lre[j][i], lim[j][i]);
            const bool inside = rowin && (i ? colin1 : colin0);
            if (S < 4) {
                const double cy = (S == 3) ? t.dt : t.half_dt;
                yr[i] = inside ? fma(k.x, cy, u_re[i]) : 0.0;
                yi[i] = inside ? fma(k.y, cy, u_im[i]) : 0.0;
            }
            if (OWNED) {
                if (S == 1) {
                    own.acc[j][i] = k;
                } else if (S < 4) {
                    own.acc[j][i].x = fma(2.0, k.x, own.acc[j][i].x);
                    own.acc[j][i].y = fma(2.0, k.y, own.acc[j][i].y);
                } else {
                    yr[i] = fma(own.acc[j][i].x + k.x, t.dt6, u_re[i]);
                    yi[i] = fma(own.acc[j][i].y + k.y, t.dt6, u_im[i]);
                }
            }
        }
        if (S < 4) {
            sts2(dre + o1 + j * C::W1, yr[0], yr[1]);
            sts2(dim_ + o1 + j * C::W1, yi[0], yi[1]);
            if (OWNED) {
#pragma unroll
                for (int i = 0; i < 2; ++i) { own.yre[j][i] = yr[i]; own.yim[j][i] = yi[i]; }
            }
        } else if (OWNED) {
            if (rowin && ly >= t.out_row0 && ly < t.out_row1) {
                if (PLANAR) {
                    const size_t o = (size_t)ly * out.pitch + gx;      // even: 16-byte aligned pair
                    if (colin1) {
                        *reinterpret_cast<double2 *>(out.re + o) = make_double2(yr[0], yr[1]);
                        *reinterpret_cast<double2 *>(out.im + o) = make_double2(yi[0], yi[1]);
                    } else if (colin0) {
                        out.re[o] = yr[0];
                        out.im[o] = yi[0];
                    }
                } else {
                    double2 *q = out.aos + (size_t)ly * t.cols + gx;
                    if (colin0) q[0] = make_double2(yr[0], yi[0]);
                    if (colin1) q[1] = make_double2(yr[1], yi[1]);
                }
            }
        }
    }
}

template <typename C, int S, bool PLANAR>
__device__ __forceinline__ void run_stage(const Smem<C> &sm, const TileCtx &t, const RhsCoeffs &c,
                                          const double (&wx)[2 * C::K + 1], const double (&wy)[2 * C::K + 1],
                                          Owned<C::MB> &own, const OutRef &out)
{
    using G = StageGrid<C, S>;
    const int tid = threadIdx.x;
    micro_tile<C, S, true, PLANAR>(sm, t, c, wx, wy, C::TMX0 + tid % C::TMW, C::TMY0 + tid / C::TMW, own, out);
    if (S < 4) {
        for (int idx = tid; idx < G::NRING; idx += C::THREADS) {
            int mx, my;
            G::locate(idx, mx, my);
            micro_tile<C, S, false, PLANAR>(sm, t, c, wx, wy, mx, my, own, out);
        }
    }
}

template <int K>
struct WeightsArg {
    double wx[2 * K + 1];
    double wy[2 * K + 1];
};

// ---- TMA / mbarrier primitives (PTX as emitted by CUTLASS's SM90_TMA_LOAD_3D and ClusterTransactionBarrier) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *map, uint32_t bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

struct TensorMap {
    alignas(64) unsigned char bytes[128];
};

template <typename C, bool UNIFORM, bool TMA>
__global__ void __launch_bounds__(C::THREADS, C::MINBLOCKS)
rk4_step_fused_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ WeightsArg<C::K> wa,
                      const __grid_constant__ TensorMap map_in, const __grid_constant__ TensorMap map_cp)
{
    constexpr int K = C::K;
    extern __shared__ __align__(128) double smem_raw[];
    Smem<C> sm{smem_raw};

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % a.tiles_x, tile_y = blockIdx.x / a.tiles_x;
    const size_t member = blockIdx.y;
    const size_t plane = (size_t)a.rows * a.cols;
    const double2 *__restrict__ in = a.in + member * plane;
    const double *__restrict__ P = a.pumping + member * plane;
    OutRef out;
    if (TMA) {
        const size_t pplane = (size_t)a.rows * a.pitch;
        out.aos = nullptr;
        out.re = a.pout + (2 * member) * pplane;
        out.im = a.pout + (2 * member + 1) * pplane;
        out.pitch = a.pitch;
    } else {
        out.aos = a.out + member * plane;
        out.re = out.im = nullptr;
        out.pitch = 0;
    }
    const RhsCoeffs c = UNIFORM ? a.cu : load_rhs_coeffs(a.coeffs + member * 23);

    TileCtx t;
    t.x0 = tile_x * C::TX - C::HXL;
    t.y0 = a.out_row0 + tile_y * C::TY - C::FY0;
    t.cols = a.cols; t.grow0 = a.grow0; t.grows = a.grows;
    t.out_row0 = a.out_row0; t.out_row1 = a.out_row1;
    t.half_dt = a.half_dt; t.dt = a.dt; t.dt6 = a.dt6;

    double wx[2 * K + 1], wy[2 * K + 1];
#pragma unroll
    for (int i = 0; i < 2 * K + 1; ++i) { wx[i] = wa.wx[i]; wy[i] = wa.wy[i]; }

    if (TMA) {
        // ---- fill by TMA: three boxes (re frame, im frame, c12*P); elements outside the arrays arrive
        // as zeros, which is exactly the truncated (zero outside the square) stencil boundary ----
        // Programmatic dependent launch: the next step's grid may start (and run this prologue, including the
        // fetch of c12*P, which no step writes) while this grid is still running; psi is requested only after
        // griddepcontrol.wait, i.e. once the previous step has completed and its stores are visible.  Every
        // global store of this grid comes after the psi frame has arrived, hence after that wait too.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const uint32_t bar = smem_u32(smem_raw + 2 * C::PLANE0 + 5 * C::PLANE1);
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            constexpr uint32_t bytes = sizeof(double) * (2 * C::W0 * C::H0 + C::W1 * C::H1);
            mbar_expect_tx(bar, bytes);
            tma_load_3d(smem_u32(sm.cp()), &map_cp, bar, t.x0 + C::OX, t.y0 + C::OY, (int)member);
            asm volatile("griddepcontrol.wait;" ::: "memory");
            tma_load_3d(smem_u32(sm.re0()), &map_in, bar, t.x0, t.y0, (int)(2 * member));
            tma_load_3d(smem_u32(sm.im0()), &map_in, bar, t.x0, t.y0, (int)(2 * member + 1));
        }
        mbar_wait(bar, 0);
    } else {
    // ---- fill: psi frame (zero outside the local array / the domain), loads batched 4 deep -------
    {
        double *b0r = sm.re0(), *b0i = sm.im0();
        constexpr int PAIRS = C::W0 / 2, TOTAL = PAIRS * C::H0, UNROLL = 4;
        for (int base = tid; base < TOTAL; base += UNROLL * C::THREADS) {
            double2 v0[UNROLL], v1[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int i = base + u * C::THREADS;
                v0[u] = make_double2(0.0, 0.0);
                v1[u] = v0[u];
                if (i < TOTAL) {
                    const int fy = i / PAIRS, fx = 2 * (i % PAIRS);
                    const int lx = t.x0 + fx, ly = t.y0 + fy, gy = ly + t.grow0;
                    if (ly >= 0 && ly < a.rows && gy >= 0 && gy < t.grows) {
                        const double2 *g = in + (size_t)ly * t.cols + lx;
                        if (lx >= 0 && lx < t.cols) v0[u] = g[0];
                        if (lx + 1 >= 0 && lx + 1 < t.cols) v1[u] = g[1];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int i = base + u * C::THREADS;
                if (i < TOTAL) {
                    const int o = (i / PAIRS) * C::W0 + 2 * (i % PAIRS);
                    sts2(b0r + o, v0[u].x, v1[u].x);
                    sts2(b0i + o, v0[u].y, v1[u].y);
                }
            }
        }
        // c12*P on the owned region
        double *cpp = sm.cp();
        constexpr int PAIRS1 = C::W1 / 2, TOTAL1 = PAIRS1 * C::H1;
        for (int base = tid; base < TOTAL1; base += UNROLL * C::THREADS) {
            double p0[UNROLL], p1[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int i = base + u * C::THREADS;
                p0[u] = 0.0; p1[u] = 0.0;
                if (i < TOTAL1) {
                    const int ry = i / PAIRS1, rx = 2 * (i % PAIRS1);
                    const int lx = t.x0 + C::OX + rx, ly = t.y0 + C::OY + ry, gy = ly + t.grow0;
                    if (ly >= 0 && ly < a.rows && gy >= 0 && gy < t.grows) {
                        const double *g = P + (size_t)ly * t.cols + lx;
                        if (lx >= 0 && lx < t.cols) p0[u] = g[0];
                        if (lx + 1 >= 0 && lx + 1 < t.cols) p1[u] = g[1];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int i = base + u * C::THREADS;
                if (i < TOTAL1) sts2(cpp + (i / PAIRS1) * C::W1 + 2 * (i % PAIRS1), c.c12 * p0[u], c.c12 * p1[u]);
            }
        }
    }
    __syncthreads();
    }

    Owned<C::MB> own;
    run_stage<C, 1, TMA>(sm, t, c, wx, wy, own, out);
    __syncthreads();
    run_stage<C, 2, TMA>(sm, t, c, wx, wy, own, out);
    __syncthreads();
    run_stage<C, 3, TMA>(sm, t, c, wx, wy, own, out);
    __syncthreads();
    run_stage<C, 4, TMA>(sm, t, c, wx, wy, own, out);
}

// ---- persistent variant: the whole time loop in ONE launch ------------------------------------------
// For grids whose tiles all fit on the GPU at once (512^2: 256 CTAs on 148 SMs x 2) every CTA keeps its
// tile for all steps and synchronises only with its (up to 8) neighbouring tiles through per-tile step
// counters in global memory: a CTA may start step s once its neighbours have published step s-1 (which
// also means they no longer read the buffer this CTA is about to overwrite).  No launch gaps, no
// end-of-kernel tail, neighbours drift freely.  Launched cooperatively so that co-residency -- which the
// spin-waits rely on -- is guaranteed by the driver; a wait that exceeds ~1 s traps instead of hanging.
__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}

template <typename C, bool UNIFORM>
__global__ void __launch_bounds__(C::THREADS, C::MINBLOCKS)
rk4_persistent_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ WeightsArg<C::K> wa,
                      const __grid_constant__ TensorMap map_a, const __grid_constant__ TensorMap map_b,
                      const __grid_constant__ TensorMap map_cp, double *buf_a, double *buf_b, int *flags, int steps)
{
    constexpr int K = C::K;
    extern __shared__ __align__(128) double smem_raw[];
    Smem<C> sm{smem_raw};

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % a.tiles_x, tile_y = blockIdx.x / a.tiles_x;
    const size_t member = blockIdx.y;
    const size_t pplane = (size_t)a.rows * a.pitch;
    const RhsCoeffs c = UNIFORM ? a.cu : load_rhs_coeffs(a.coeffs + member * 23);

    TileCtx t;
    t.x0 = tile_x * C::TX - C::HXL;
    t.y0 = tile_y * C::TY - C::FY0;
    t.cols = a.cols; t.grow0 = 0; t.grows = a.rows;
    t.out_row0 = 0; t.out_row1 = a.rows;
    t.half_dt = a.half_dt; t.dt = a.dt; t.dt6 = a.dt6;

    double wx[2 * K + 1], wy[2 * K + 1];
#pragma unroll
    for (int i = 0; i < 2 * K + 1; ++i) { wx[i] = wa.wx[i]; wy[i] = wa.wy[i]; }

    // step counters: one per tile; thread i < 8 watches neighbour i
    int *my_flag = flags + (member * gridDim.x + blockIdx.x);
    const int *watch = nullptr;
    if (tid < 8) {
        const int d = tid < 4 ? tid : tid + 1;            // skip (0, 0)
        const int nx = tile_x + d % 3 - 1, ny = tile_y + d / 3 - 1;
        if (nx >= 0 && nx < a.tiles_x && ny >= 0 && ny < a.tiles_y)
            watch = flags + (member * gridDim.x + (size_t)ny * a.tiles_x + nx);
    }

    const uint32_t bar = smem_u32(smem_raw + 2 * C::PLANE0 + 5 * C::PLANE1);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {      // c12*P does not change: fetched once
        mbar_expect_tx(bar, sizeof(double) * C::W1 * C::H1);
        tma_load_3d(smem_u32(sm.cp()), &map_cp, bar, t.x0 + C::OX, t.y0 + C::OY, (int)member);
    }
    mbar_wait(bar, 0);

    Owned<C::MB> own;
    for (int s = 0; s < steps; ++s) {
        if (s > 0 && watch) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(watch) < s)
                if (clock64() - t0 > 4000000000ll) __trap();   // a neighbour never arrived: fail loudly
        }
        __syncthreads();   // neighbours are ready; every warp of this CTA is done with the previous step
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, sizeof(double) * 2 * C::W0 * C::H0);
            const TensorMap *src = (s & 1) ? &map_b : &map_a;
            tma_load_3d(smem_u32(sm.re0()), src, bar, t.x0, t.y0, (int)(2 * member));
            tma_load_3d(smem_u32(sm.im0()), src, bar, t.x0, t.y0, (int)(2 * member + 1));
        }
        mbar_wait(bar, (s + 1) & 1);

        OutRef out;
        double *dst = (s & 1) ? buf_a : buf_b;
        out.aos = nullptr;
        out.re = dst + (2 * member) * pplane;
        out.im = dst + (2 * member + 1) * pplane;
        out.pitch = a.pitch;

        run_stage<C, 1, true>(sm, t, c, wx, wy, own, out);
        __syncthreads();
        run_stage<C, 2, true>(sm, t, c, wx, wy, own, out);
        __syncthreads();
        run_stage<C, 3, true>(sm, t, c, wx, wy, own, out);
        __syncthreads();
        run_stage<C, 4, true>(sm, t, c, wx, wy, own, out);

        // publish: the CTA barrier orders every thread's stores before thread 0, whose gpu-scope fence
        // and release store then make them visible (cumulativity) to whoever acquires the counter
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            st_release_gpu(my_flag, s + 1);
        }
    }
}

template <typename C, bool UNIFORM, bool TMA>
int configure_fused()
{
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(rk4_step_fused_kernel<C, UNIFORM, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(rk4_step_fused_kernel<C, UNIFORM, TMA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    return 0;
}

template <typename C>
WeightsArg<C::K> pack_weights(const CrossWeights &w)
{
    WeightsArg<C::K> wa;
    for (int i = 0; i < 2 * C::K + 1; ++i) { wa.wx[i] = w.wx[i]; wa.wy[i] = w.wy[i]; }
    return wa;
}

template <typename C, bool UNIFORM>
int launch_fused_cfg(const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    int rc = configure_fused<C, UNIFORM, false>();
    if (rc) return rc;
    FusedArgs a{};
    a.rows = s.rows; a.cols = s.cols; a.grow0 = s.grow0; a.grows = s.grows;
    a.out_row0 = s.out_row0; a.out_row1 = s.out_row1;
    a.tiles_x = (s.cols + C::TX - 1) / C::TX;
    a.tiles_y = (s.out_row1 - s.out_row0 + C::TY - 1) / C::TY;
    a.in = s.in; a.out = s.out; a.pumping = s.pumping; a.coeffs = s.coeffs;
    if (UNIFORM) a.cu = *s.uniform;
    a.dt = s.dt; a.half_dt = s.dt / 2; a.dt6 = s.dt / 6;
    if (a.tiles_x <= 0 || a.tiles_y <= 0) return 0;
    const dim3 grid((unsigned)(a.tiles_x * a.tiles_y), (unsigned)s.batch);
    TensorMap none{};
    rk4_step_fused_kernel<C, UNIFORM, false><<<grid, C::THREADS, C::SMEM, stream>>>(a, pack_weights<C>(w), none, none);
    count_launches(1);
    return (int)cudaGetLastError();
}

template <typename C, bool UNIFORM>
int launch_fused_planar_cfg(const Fused2DPlanar &p, const PlanarMaps &maps, bool a_to_b, bool overlap, const CrossWeights &w,
                            cudaStream_t stream)
{
    int rc = configure_fused<C, UNIFORM, true>();
    if (rc) return rc;
    FusedArgs a{};
    a.rows = p.rows; a.cols = p.cols; a.grow0 = p.grow0; a.grows = p.grows;
    a.out_row0 = p.out_row0; a.out_row1 = p.out_row1;
    a.tiles_x = (p.cols + C::TX - 1) / C::TX;
    a.tiles_y = (p.out_row1 - p.out_row0 + C::TY - 1) / C::TY;
    if (a.tiles_x <= 0 || a.tiles_y <= 0) return 0;
    a.pout = a_to_b ? p.psi_b : p.psi_a;
    a.pitch = p.pitch;
    a.coeffs = p.coeffs;
    if (UNIFORM) a.cu = *p.uniform;
    a.dt = p.dt; a.half_dt = p.dt / 2; a.dt6 = p.dt / 6;
    TensorMap in, cp;
    std::memcpy(in.bytes, a_to_b ? maps.psi_a : maps.psi_b, 128);
    std::memcpy(cp.bytes, maps.cp, 128);
    const dim3 grid((unsigned)(a.tiles_x * a.tiles_y), (unsigned)p.batch);
    // programmatic stream serialization: this grid may begin before the previous one in the stream has drained
    // (the kernel orders its own reads and writes with griddepcontrol.wait)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = overlap ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const WeightsArg<C::K> wa = pack_weights<C>(w);
    cudaError_t e = cudaLaunchKernelEx(&cfg, rk4_step_fused_kernel<C, UNIFORM, true>, a, wa, in, cp);
    count_launches(1);
    return (int)e;
}

template <typename C, bool UNIFORM>
int persistent_cfg(const Fused2DPlanar &p, const PlanarMaps *maps, int steps, int *flags, const CrossWeights *w,
                   cudaStream_t stream, long long *tiles_out, long long *capacity_out)
{
    static int capacity[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return fail(NLSB_EINVAL, "device ordinal %d out of range", dev);
    if (!capacity[dev]) {
        e = cudaFuncSetAttribute(rk4_persistent_kernel<C, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(rk4_persistent_kernel<C, UNIFORM>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rk4_persistent_kernel<C, UNIFORM>, C::THREADS, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        capacity[dev] = per_sm * sms > 0 ? per_sm * sms : -1;
    }
    const int tiles_x = (p.cols + C::TX - 1) / C::TX, tiles_y = (p.rows + C::TY - 1) / C::TY;
    *tiles_out = (long long)tiles_x * tiles_y * p.batch;
    *capacity_out = capacity[dev];
    if (!maps) return 0;                     // query only
    if (*tiles_out > *capacity_out) return fail(NLSB_ESIZE, "persistent kernel: %lld tiles exceed the %lld resident CTAs",
                                                *tiles_out, *capacity_out);
    FusedArgs a{};
    a.rows = p.rows; a.cols = p.cols; a.grow0 = 0; a.grows = p.rows; a.out_row0 = 0; a.out_row1 = p.rows;
    a.tiles_x = tiles_x; a.tiles_y = tiles_y;
    a.pitch = p.pitch; a.coeffs = p.coeffs;
    if (UNIFORM) a.cu = *p.uniform;
    a.dt = p.dt; a.half_dt = p.dt / 2; a.dt6 = p.dt / 6;
    WeightsArg<C::K> wa = pack_weights<C>(*w);
    TensorMap ma, mb, mc;
    std::memcpy(ma.bytes, maps->psi_a, 128);
    std::memcpy(mb.bytes, maps->psi_b, 128);
    std::memcpy(mc.bytes, maps->cp, 128);
    double *buf_a = p.psi_a, *buf_b = p.psi_b;
    void *args[] = {&a, &wa, &ma, &mb, &mc, &buf_a, &buf_b, &flags, &steps};
    const dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)p.batch);
    e = cudaLaunchCooperativeKernel((const void *)rk4_persistent_kernel<C, UNIFORM>, grid, dim3(C::THREADS), args, C::SMEM,
                                    stream);
    if (e != cudaSuccess) return (int)e;
    count_launches(1);
    return 0;
}

// Tile shapes.  variant 0: 32x32 tiles, 256 threads, two CTAs per SM (one CTA's fill and barriers
// hide behind the other's arithmetic); variant 1: 32x64 tiles, 512 threads, one CTA per SM (less
// redundant halo work).
template <int K>
struct Shapes {
    using Small = FusedCfg<K, 32, 2, 256, (K == 3) ? 1 : 2>;
    using Tall = FusedCfg<(K == 3) ? 2 : K, 64, 2, 512, 1>;
};

template <int K>
int launch_fused_k(int variant, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    using Small = typename Shapes<K>::Small;
    using Tall = typename Shapes<K>::Tall;
    if (K == 3 || variant == 0)
        return s.uniform ? launch_fused_cfg<Small, true>(s, w, stream) : launch_fused_cfg<Small, false>(s, w, stream);
    return s.uniform ? launch_fused_cfg<Tall, true>(s, w, stream) : launch_fused_cfg<Tall, false>(s, w, stream);
}

template <int K>
int launch_fused_planar_k(int variant, const Fused2DPlanar &p, const PlanarMaps &maps, bool a_to_b, bool overlap,
                          const CrossWeights &w, cudaStream_t stream)
{
    using Small = typename Shapes<K>::Small;
    using Tall = typename Shapes<K>::Tall;
    if (K == 3 || variant == 0)
        return p.uniform ? launch_fused_planar_cfg<Small, true>(p, maps, a_to_b, overlap, w, stream)
                         : launch_fused_planar_cfg<Small, false>(p, maps, a_to_b, overlap, w, stream);
    return p.uniform ? launch_fused_planar_cfg<Tall, true>(p, maps, a_to_b, overlap, w, stream)
                     : launch_fused_planar_cfg<Tall, false>(p, maps, a_to_b, overlap, w, stream);
}

template <int K>
int persistent_k(int variant, const Fused2DPlanar &p, const PlanarMaps *maps, int steps, int *flags,
                 const CrossWeights *w, cudaStream_t stream, long long *tiles, long long *capacity)
{
    using Small = typename Shapes<K>::Small;
    using Tall = typename Shapes<K>::Tall;
    if (K == 3 || variant == 0)
        return p.uniform ? persistent_cfg<Small, true>(p, maps, steps, flags, w, stream, tiles, capacity)
                         : persistent_cfg<Small, false>(p, maps, steps, flags, w, stream, tiles, capacity);
    return p.uniform ? persistent_cfg<Tall, true>(p, maps, steps, flags, w, stream, tiles, capacity)
                     : persistent_cfg<Tall, false>(p, maps, steps, flags, w, stream, tiles, capacity);
}

int persistent_dispatch(int order, int variant, const Fused2DPlanar &p, const PlanarMaps *maps, int steps, int *flags,
                        const CrossWeights *w, cudaStream_t stream, long long *tiles, long long *capacity)
{
    switch (order) {
    case 3: return persistent_k<1>(variant, p, maps, steps, flags, w, stream, tiles, capacity);
    case 5: return persistent_k<2>(variant, p, maps, steps, flags, w, stream, tiles, capacity);
    case 7: return persistent_k<3>(variant, p, maps, steps, flags, w, stream, tiles, capacity);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_plane_map(void *out128, double *base, int cols, int rows, int planes, int pitch, int box_w, int box_h)
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
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(double), (cuuint64_t)pitch * rows * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSB_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    std::memcpy(out128, &map, 128);
    return 0;
}

template <typename C>
int make_maps_cfg(const Fused2DPlanar &p, PlanarMaps *maps)
{
    int rc = encode_plane_map(maps->psi_a, p.psi_a, p.cols, p.rows, 2 * p.batch, p.pitch, C::W0, C::H0);
    if (!rc) rc = encode_plane_map(maps->psi_b, p.psi_b, p.cols, p.rows, 2 * p.batch, p.pitch, C::W0, C::H0);
    if (!rc) rc = encode_plane_map(maps->cp, p.cp, p.cols, p.rows, p.batch, p.pitch, C::W1, C::H1);
    return rc;
}

template <int K>
int make_maps_k(int variant, const Fused2DPlanar &p, PlanarMaps *maps)
{
    if (K == 3 || variant == 0) return make_maps_cfg<typename Shapes<K>::Small>(p, maps);
    return make_maps_cfg<typename Shapes<K>::Tall>(p, maps);
}

// ---- interleaved <-> planar ---------------------------------------------------------------------
__global__ void split_planar_kernel(int rows, int cols, int pitch, const double2 *__restrict__ psi,
                                    const double *__restrict__ pumping, const double *__restrict__ coeffs,
                                    double *__restrict__ planes, double *__restrict__ cp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const size_t member = blockIdx.z;
    if (x >= pitch) return;
    const size_t pplane = (size_t)rows * pitch, o = (size_t)y * pitch + x;
    double2 v = make_double2(0.0, 0.0);
    double c = 0.0;
    if (x < cols) {
        const size_t g = (member * rows + y) * cols + x;
        v = psi[g];
        c = coeffs[member * 23 + 11] * pumping[g];     // c12 * P, rounded once (nls.f90:580 association)
    }
    planes[(2 * member) * pplane + o] = v.x;
    planes[(2 * member + 1) * pplane + o] = v.y;
    cp[member * pplane + o] = c;
}

__global__ void join_planar_kernel(int rows, int cols, int pitch, const double *__restrict__ planes,
                                   double2 *__restrict__ psi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const size_t member = blockIdx.z;
    if (x >= cols) return;
    const size_t pplane = (size_t)rows * pitch, o = (size_t)y * pitch + x;
    psi[(member * rows + y) * cols + x] = make_double2(planes[(2 * member) * pplane + o], planes[(2 * member + 1) * pplane + o]);
}

}  // namespace

int make_planar_maps(int order, int variant, const Fused2DPlanar &p, PlanarMaps *maps)
{
    switch (order) {
    case 3: return make_maps_k<1>(variant, p, maps);
    case 5: return make_maps_k<2>(variant, p, maps);
    case 7: return make_maps_k<3>(variant, p, maps);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

int persistent_2d_fits(int order, int variant, const Fused2DPlanar &p, bool *fits, long long *tiles)
{
    long long capacity = 0;
    int rc = persistent_dispatch(order, variant, p, nullptr, 0, nullptr, nullptr, nullptr, tiles, &capacity);
    if (rc) return rc;
    *fits = p.batch <= 65535 && *tiles <= capacity;
    return 0;
}

int launch_rk4_persistent_2d_planar(int order, int variant, const Fused2DPlanar &p, const PlanarMaps &maps, int steps,
                                    int *flags, const CrossWeights &w, cudaStream_t stream)
{
    long long tiles = 0, capacity = 0;
    return persistent_dispatch(order, variant, p, &maps, steps, flags, &w, stream, &tiles, &capacity);
}

int launch_split_planar(const Fused2DPlanar &p, const double2 *psi, const double *pumping, cudaStream_t stream)
{
    if (p.rows > 65535 || p.batch > 65535) return fail(NLSB_ESIZE, "planar split: rows/batch exceed the grid limits");
    const dim3 block(128), grid((p.pitch + 127) / 128, p.rows, p.batch);
    split_planar_kernel<<<grid, block, 0, stream>>>(p.rows, p.cols, p.pitch, psi, pumping, p.coeffs, p.psi_a, p.cp);
    count_launches(1);
    return (int)cudaGetLastError();
}

int launch_join_planar(const Fused2DPlanar &p, bool from_b, double2 *psi, cudaStream_t stream)
{
    const dim3 block(128), grid((p.cols + 127) / 128, p.rows, p.batch);
    join_planar_kernel<<<grid, block, 0, stream>>>(p.rows, p.cols, p.pitch, from_b ? p.psi_b : p.psi_a, psi);
    count_launches(1);
    return (int)cudaGetLastError();
}

int launch_rk4_step_fused_2d(int order, int variant, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    if (s.batch > 65535) return fail(NLSB_ESIZE, "batch = %d exceeds the grid y-limit 65535", s.batch);
    switch (order) {
    case 3: return launch_fused_k<1>(variant, s, w, stream);
    case 5: return launch_fused_k<2>(variant, s, w, stream);
    case 7: return launch_fused_k<3>(variant, s, w, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

int launch_rk4_step_fused_2d_planar(int order, int variant, const Fused2DPlanar &p, const PlanarMaps &maps, bool a_to_b,
                                    bool overlap, const CrossWeights &w, cudaStream_t stream)
{
    if (p.batch > 32767) return fail(NLSB_ESIZE, "batch = %d exceeds the planar-path limit 32767", p.batch);
    switch (order) {
    case 3: return launch_fused_planar_k<1>(variant, p, maps, a_to_b, overlap, w, stream);
    case 5: return launch_fused_planar_k<2>(variant, p, maps, a_to_b, overlap, w, stream);
    case 7: return launch_fused_planar_k<3>(variant, p, maps, a_to_b, overlap, w, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

}  // namespace nlsb
