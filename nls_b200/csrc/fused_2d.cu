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
// Data layout in shared memory: three frames (B0 = psi, B1, B2 = stage inputs, ping-pong) of
// W x H nodes stored as separate re / im planes of doubles, plus one plane of c12*P.  A thread
// works on micro-tiles of 2 (x) x 4 (y) nodes: with the planar layout a 16-byte LDS/STS moves the
// same component of two x-adjacent nodes, consecutive lanes touch consecutive 16-byte words
// (conflict-free), and a micro-tile needs 4 loads per node and stage instead of 9.
// The first 256 micro-tiles of every stage's work list are the tile itself, always mapped to the
// same thread, so the RK accumulator of a node lives in that thread's registers across stages;
// the halo-ring micro-tiles follow in the list and carry no state.
//
// Outside the square the field is identically zero at every stage (the reference's truncated
// band matrix): out-of-domain nodes are forced to zero after each stage.  Every node's value is
// produced by the same sequence of roundings whatever tile, CTA or GPU computes it, so results
// do not depend on the tiling or on a slab decomposition.

#include "device_math.cuh"
#include "kernels.h"

namespace nlsb {

namespace {

constexpr int kFusedThreads = 256;

template <int K_>
struct FusedCfg {
    static constexpr int K = K_;
    static constexpr int TX = 32;
    static constexpr int TY = (K_ == 3) ? 32 : 64;
    static constexpr int PH = (K_ + 1) / 2;              // x-neighbour pairs on each side
    static constexpr int HXL = 4 * K_ + 2 * (K_ & 1);    // frame columns left of the tile (even)
    static constexpr int OX = HXL - 3 * K_ - (K_ & 1);   // first owned column (even, <= tile - 3K)
    static constexpr int W = TX + 2 * HXL;               // frame width (even)
    static constexpr int NMX = (HXL + TX + 3 * K_ - OX + 1) / 2;   // micro-tiles per row of the owned region
    static constexpr int DY = 4 * ((3 * K_ + 3) / 4);    // owned rows above the tile (multiple of 4)
    static constexpr int OY = K_;                        // first owned row
    static constexpr int FY0 = OY + DY;                  // first tile row
    static constexpr int NMY = DY / 4 + TY / 4 + (3 * K_ + 3) / 4;
    static constexpr int H = OY + 4 * NMY + K_;          // frame height
    static constexpr int PLANE = W * H;                  // doubles per plane
    static constexpr int NPLANES = 7;                    // 3 frames x (re, im) + c12*P
    static constexpr size_t SMEM = sizeof(double) * PLANE * NPLANES;
    static constexpr int TMX0 = (HXL - OX) / 2;          // first tile micro-tile column
    static constexpr int TMY0 = DY / 4;
    static constexpr int TMW = TX / 2, TMH = TY / 4;
    static constexpr int NTILE = TMW * TMH;
    static_assert(OX % 2 == 0 && W % 2 == 0, "pairs must be 16-byte aligned");
    static_assert(OX >= 2 * PH && OX + 2 * NMX + 2 * PH <= W, "x halo of the frame too small");
    static_assert(SMEM <= 227 * 1024, "frame does not fit in shared memory");
};

struct FusedArgs {
    int rows, cols;          // extent of the local arrays (rows may include slab halo rows)
    int grow0, grows;        // global row index of local row 0, global number of rows
    int out_row0, out_row1;  // local rows [out_row0, out_row1) are written
    int tiles_x, tiles_y;    // tile grid covering cols x (out_row1 - out_row0)
    const double2 *in;       // [batch][rows][cols]
    double2 *out;            // [batch][rows][cols]
    const double *pumping;   // [batch][rows][cols]
    const double *coeffs;    // [batch][23]
    double dt;
};

__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void sts2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// Micro-tile `idx` of stage S (1-based): tile micro-tiles first, then the halo ring of the stage
// enumerated as top band, bottom band, left columns, right columns.
template <typename C, int S>
struct StageGrid {
    static constexpr int E = (4 - S) * C::K;
    static constexpr int MX0 = (C::HXL - E - C::OX) / 2;
    static constexpr int MX1 = (C::HXL + C::TX + E - C::OX + 1) / 2;
    static constexpr int MY0 = (C::FY0 - E - C::OY) / 4;
    static constexpr int MY1 = (C::FY0 + C::TY + E - C::OY + 3) / 4;
    static constexpr int GW = MX1 - MX0;
    static constexpr int NTOP = GW * (C::TMY0 - MY0);
    static constexpr int NBOT = GW * (MY1 - (C::TMY0 + C::TMH));
    static constexpr int LW = C::TMX0 - MX0;
    static constexpr int RW = MX1 - (C::TMX0 + C::TMW);
    static constexpr int NLEFT = LW * C::TMH;
    static constexpr int NRIGHT = RW * C::TMH;
    static constexpr int COUNT = C::NTILE + NTOP + NBOT + NLEFT + NRIGHT;
    static_assert(MX0 >= 0 && MX1 <= C::NMX && MY0 >= 0 && MY1 <= C::NMY, "stage region leaves the owned region");

    __device__ static __forceinline__ void locate(int idx, int &mx, int &my)
    {
        if (idx < C::NTILE) {
            mx = C::TMX0 + idx % C::TMW;
            my = C::TMY0 + idx / C::TMW;
            return;
        }
        idx -= C::NTILE;
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
        if (idx < NLEFT) {
            mx = MX0 + idx % (LW > 0 ? LW : 1);
            my = C::TMY0 + idx / (LW > 0 ? LW : 1);
            return;
        }
        idx -= NLEFT;
        mx = C::TMX0 + C::TMW + idx % (RW > 0 ? RW : 1);
        my = C::TMY0 + idx / (RW > 0 ? RW : 1);
    }
};

template <typename C>
struct Smem {
    double *plane;
    __device__ __forceinline__ double *re(int frame) const { return plane + (2 * frame) * C::PLANE; }
    __device__ __forceinline__ double *im(int frame) const { return plane + (2 * frame + 1) * C::PLANE; }
    __device__ __forceinline__ double *cp() const { return plane + 6 * C::PLANE; }
};

struct TileCtx {
    int x0, y0;          // local array coordinates of frame node (0, 0)
    int rows, cols, grow0, grows;
    int out_row0, out_row1;
    double half_dt, dt, dt6;
};

// One micro-tile of stage S.  src/dst: frame indices of the stage input / next stage input.
template <typename C, int S, bool WITH_ACC>
__device__ __forceinline__ void micro_tile(const Smem<C> &sm, const TileCtx &t, const RhsCoeffs &c,
                                           const double (&wx)[2 * C::K + 1], const double (&wy)[2 * C::K + 1],
                                           int mx, int my, bool is_tile, double2 (&acc)[4][2], double2 *__restrict__ out)
{
    constexpr int K = C::K, W = C::W, PH = C::PH;
    constexpr int src = (S == 1) ? 0 : (S == 2) ? 1 : (S == 3) ? 2 : 1;
    constexpr int dst = (S == 1) ? 1 : (S == 2) ? 2 : 1;
    const int fx = C::OX + 2 * mx, fy = C::OY + 4 * my;
    const double *sre = sm.re(src), *sim = sm.im(src);

    double lre[4][2], lim[4][2], cre[4][2], cim[4][2];
#pragma unroll
    for (int r = -K; r < 4 + K; ++r) {
        const int o = (fy + r) * W + fx;
        const double2 vr = lds2(sre + o), vi = lds2(sim + o);
        if (r >= 0 && r < 4) {
            cre[r][0] = vr.x; cre[r][1] = vr.y;
            cim[r][0] = vi.x; cim[r][1] = vi.y;
        }
        // scatter input row r into the output rows it touches: output row j receives, in this order,
        // the taps of the K rows above it, its own row (x taps and centre), the K rows below it
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = r - j;               // input row = output row + d
            if (d == 0) {
                // x direction (includes the centre weight)
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
                    // rows j-K .. j-1 have already contributed (r starts at -K), so lre/lim are initialised
                    double ar = lre[j][i], ai = lim[j][i];
#pragma unroll
                    for (int tp = -K; tp <= K; ++tp) {
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
    const double *ure = sm.re(0), *uim = sm.im(0), *cpp = sm.cp();
    double *dre = sm.re(dst), *dim_ = sm.im(dst);
    const int gx = t.x0 + fx;
    const bool colin0 = gx >= 0 && gx < t.cols, colin1 = gx + 1 >= 0 && gx + 1 < t.cols;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = (fy + j) * W + fx;
        const int ly = t.y0 + fy + j;
        const int gy = ly + t.grow0;
        const bool rowin = gy >= 0 && gy < t.grows;
        const double2 cpv = lds2(cpp + o);
        double2 ur, ui;
        if (S == 1) {
            ur = make_double2(cre[j][0], cre[j][1]);
            ui = make_double2(cim[j][0], cim[j][1]);
        } else {
            ur = lds2(ure + o);
            ui = lds2(uim + o);
        }
        double yr[2], yi[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double2 y = make_double2(cre[j][i], cim[j][i]);
            const double2 k = rhs_point(c, i ? cpv.y : cpv.x, y, lre[j][i], lim[j][i]);
            const double u_re = i ? ur.y : ur.x, u_im = i ? ui.y : ui.x;
            const bool inside = rowin && (i ? colin1 : colin0);
            if (S < 4) {
                const double cy = (S == 3) ? t.dt : t.half_dt;
                yr[i] = inside ? fma(k.x, cy, u_re) : 0.0;
                yi[i] = inside ? fma(k.y, cy, u_im) : 0.0;
            }
            if (WITH_ACC) {
                if (S == 1) {
                    acc[j][i] = k;
                } else if (S < 4) {
                    acc[j][i].x = fma(2.0, k.x, acc[j][i].x);
                    acc[j][i].y = fma(2.0, k.y, acc[j][i].y);
                } else {
                    yr[i] = fma(acc[j][i].x + k.x, t.dt6, u_re);
                    yi[i] = fma(acc[j][i].y + k.y, t.dt6, u_im);
                }
            }
        }
        if (S < 4) {
            sts2(dre + o, yr[0], yr[1]);
            sts2(dim_ + o, yi[0], yi[1]);
        } else if (WITH_ACC) {
            if (is_tile && rowin && ly >= t.out_row0 && ly < t.out_row1) {
                double2 *q = out + (size_t)ly * t.cols + gx;
                if (colin0) q[0] = make_double2(yr[0], yi[0]);
                if (colin1) q[1] = make_double2(yr[1], yi[1]);
            }
        }
    }
}

template <typename C, int S>
__device__ __forceinline__ void run_stage(const Smem<C> &sm, const TileCtx &t, const RhsCoeffs &c,
                                          const double (&wx)[2 * C::K + 1], const double (&wy)[2 * C::K + 1],
                                          double2 (&acc)[4][2], double2 *__restrict__ out)
{
    using G = StageGrid<C, S>;
    const int tid = threadIdx.x;
    int mx, my;
    // round 0: the tile's own micro-tiles (fixed thread mapping, carries the RK accumulator)
    if (tid < G::COUNT) {
        G::locate(tid, mx, my);
        micro_tile<C, S, true>(sm, t, c, wx, wy, mx, my, tid < C::NTILE, acc, out);
    }
    if (S < 4) {
        double2 none[4][2];
        for (int idx = tid + kFusedThreads; idx < G::COUNT; idx += kFusedThreads) {
            G::locate(idx, mx, my);
            micro_tile<C, S, false>(sm, t, c, wx, wy, mx, my, false, none, out);
        }
    }
}

template <int K>
struct WeightsArg {
    double wx[2 * K + 1];
    double wy[2 * K + 1];
};

template <int K>
__global__ void __launch_bounds__(kFusedThreads, 1)
rk4_step_fused_kernel(FusedArgs a, WeightsArg<K> wa)
{
    using C = FusedCfg<K>;
    static_assert(C::NTILE <= kFusedThreads, "tile micro-tiles must fit one round");
    extern __shared__ __align__(16) double smem_raw[];
    Smem<C> sm{smem_raw};

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % a.tiles_x, tile_y = blockIdx.x / a.tiles_x;
    const size_t member = blockIdx.y;
    const size_t plane = (size_t)a.rows * a.cols;
    const double2 *__restrict__ in = a.in + member * plane;
    double2 *__restrict__ out = a.out + member * plane;
    const double *__restrict__ P = a.pumping + member * plane;
    const RhsCoeffs c = load_rhs_coeffs(a.coeffs + member * 23);

    TileCtx t;
    t.x0 = tile_x * C::TX - C::HXL;
    t.y0 = a.out_row0 + tile_y * C::TY - C::FY0;
    t.rows = a.rows; t.cols = a.cols; t.grow0 = a.grow0; t.grows = a.grows;
    t.out_row0 = a.out_row0; t.out_row1 = a.out_row1;
    t.half_dt = a.dt / 2; t.dt = a.dt; t.dt6 = a.dt / 6;

    double wx[2 * K + 1], wy[2 * K + 1];
#pragma unroll
    for (int i = 0; i < 2 * K + 1; ++i) { wx[i] = wa.wx[i]; wy[i] = wa.wy[i]; }

    // ---- fill: psi frame (zero outside the local array / the domain) and c12*P ----------------
    {
        double *b0r = sm.re(0), *b0i = sm.im(0), *cpp = sm.cp();
        constexpr int PAIRS = C::W / 2;
        for (int i = tid; i < PAIRS * C::H; i += kFusedThreads) {
            const int fy = i / PAIRS, fx = 2 * (i % PAIRS);
            const int lx = t.x0 + fx, ly = t.y0 + fy;
            const int gy = ly + t.grow0;
            const bool rowok = ly >= 0 && ly < t.rows && gy >= 0 && gy < t.grows;
            double2 v0 = make_double2(0.0, 0.0), v1 = v0;
            double p0 = 0.0, p1 = 0.0;
            if (rowok) {
                const size_t g = (size_t)ly * t.cols + lx;
                if (lx >= 0 && lx < t.cols) { v0 = in[g]; p0 = c.c12 * P[g]; }
                if (lx + 1 >= 0 && lx + 1 < t.cols) { v1 = in[g + 1]; p1 = c.c12 * P[g + 1]; }
            }
            const int o = fy * C::W + fx;
            sts2(b0r + o, v0.x, v1.x);
            sts2(b0i + o, v0.y, v1.y);
            sts2(cpp + o, p0, p1);
        }
    }
    __syncthreads();

    double2 acc[4][2];
    run_stage<C, 1>(sm, t, c, wx, wy, acc, out);
    __syncthreads();
    run_stage<C, 2>(sm, t, c, wx, wy, acc, out);
    __syncthreads();
    run_stage<C, 3>(sm, t, c, wx, wy, acc, out);
    __syncthreads();
    run_stage<C, 4>(sm, t, c, wx, wy, acc, out);
}

template <int K>
int launch_fused_k(const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    using C = FusedCfg<K>;
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(rk4_step_fused_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    FusedArgs a;
    a.rows = s.rows; a.cols = s.cols; a.grow0 = s.grow0; a.grows = s.grows;
    a.out_row0 = s.out_row0; a.out_row1 = s.out_row1;
    a.tiles_x = (s.cols + C::TX - 1) / C::TX;
    a.tiles_y = (s.out_row1 - s.out_row0 + C::TY - 1) / C::TY;
    a.in = s.in; a.out = s.out; a.pumping = s.pumping; a.coeffs = s.coeffs; a.dt = s.dt;
    if (a.tiles_x <= 0 || a.tiles_y <= 0) return 0;
    WeightsArg<K> wa;
    for (int i = 0; i < 2 * K + 1; ++i) { wa.wx[i] = w.wx[i]; wa.wy[i] = w.wy[i]; }
    const dim3 grid((unsigned)(a.tiles_x * a.tiles_y), (unsigned)s.batch);
    rk4_step_fused_kernel<K><<<grid, kFusedThreads, C::SMEM, stream>>>(a, wa);
    count_launches(1);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_rk4_step_fused_2d(int order, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream)
{
    if (s.batch > 65535) return fail(NLSB_ESIZE, "batch = %d exceeds the grid y-limit 65535", s.batch);
    switch (order) {
    case 3: return launch_fused_k<1>(s, w, stream);
    case 5: return launch_fused_k<2>(s, w, stream);
    case 7: return launch_fused_k<3>(s, w, stream);
    }
    return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", order);
}

}  // namespace nlsb
