// kernels.h -- launchers implemented in the .cu files; called by api.cu.  All pointers are device
// pointers unless stated otherwise.  Each launcher returns a cudaError_t-compatible int (0 = ok) or
// a negative NLSB_E* code.
#pragma once

#include "internal.h"

#include <cuda_runtime.h>

namespace nlsb {

enum StageMode { kStageRhs = 0, kStageFirst = 1, kStageMid = 2, kStageLast = 3 };

// kernels_1d.cu
constexpr int kMaxResident1D = 2048;   // largest n the CTA-resident 1D kernel handles (8 nodes x 256 threads)
// diag_out8 non-null (and iters >= 1): the launch also writes the 8 scalar diagnostics per member (diag_acc.cuh) of
// the state ENTERING its last step, reduced inside that step's first stage; dx is the radial step of their weights
int launch_rk4_1d(int batch, int n, int order, int iters, double dt, const double *taps, const double *pumping,
                  const double *coeffs, double2 *psi, cudaStream_t stream, double dx = 0.0, double *diag_out8 = nullptr);
// n > kMaxResident1D: per-stage launches through global memory; work holds 3*batch*n complex values
int launch_rk4_1d_staged(int batch, int n, int order, int iters, double dt, const double *taps,
                         const double *pumping, const double *coeffs, double2 *psi, double2 *work,
                         cudaStream_t stream);
int launch_hamiltonian_1d(int batch, int n, int order, const double *taps, const double *pumping,
                          const double *coeffs, const double2 *u, double2 *v, cudaStream_t stream);
int launch_band_matvec_1d(int n, int order, const double *taps, const double *x, double *u, double sign,
                          cudaStream_t stream);

// kernels_2d.cu
struct Stage2DArgs {
    int batch, rows, cols;
    const double2 *ysrc;    // stage input (stencil source)
    const double2 *ubase;   // u at the beginning of the step
    const double *pumping;
    const double *coeffs;   // [batch][23]
    double2 *acc;           // k1 + 2 k2 + 2 k3 (read and written in place)
    double2 *ydst;          // next stage input, or v for kStageRhs, or u for kStageLast
    double cy;              // dt/2 or dt
    double cacc;            // weight of k in acc (1 or 2)
    double dt6;             // dt/6
};
int launch_stage_2d(int order, StageMode mode, const CrossWeights &w, const Stage2DArgs &a, cudaStream_t stream);
// General block-band operator (any blocks memory of the make_laplacian_2d layout, m = 3 / 5 / 7; kernels_2d.cu):
// blocks_dev is the device copy of the blocks memory; arrays in the reference's memory order (line = slow index).
int launch_stage_2d_general(int n, int m, const double *blocks_dev, StageMode mode, const Stage2DArgs &a, cudaStream_t stream);
int launch_general_matvec_2d(int n, int m, const double *blocks_dev, const double *x, double *y, double sign,
                             cudaStream_t stream);
// The halo exchange a slab's step launch performs itself (strip-marching kernel, multi-GPU; peer.cu / stream_2d.cu):
// new psi rows [up_row0, up_row0 + nrows) also go to (address of the local store + up_delta bytes) = the halo rows of
// the rank above, rows [dn_row0, dn_row0 + nrows) to the rank below; flags as in peer_flags.cuh (null: no neighbour).
struct PeerStep {
    long long up_delta, dn_delta;
    int up_row0, dn_row0, nrows;
    unsigned long long *state, *flags_mine, *flags_up, *flags_down;
    long long timeout_cycles;
};

// fused_2d.cu: one launch advances psi by one full RK4 step (in -> out, distinct buffers).
struct Fused2DStep {
    int batch, rows, cols;   // local arrays are [batch][rows][cols]
    int grow0, grows;        // global row of local row 0 and global row count (slab decomposition; else 0, rows)
    int out_row0, out_row1;  // local rows to produce
    const double2 *in;
    double2 *out;
    const double *pumping;
    const double *coeffs;    // [batch][23] on the device
    double dt;
    const RhsCoeffs *uniform;   // host: non-null when every member shares these coefficients
    // strip-marching kernel only: when non-null the step also reduces the scalar diagnostics of the state it reads
    // (`in`) into [batch][stream_2d_diag_parts] partial sums (diag_acc.cuh), area element diag_area
    void *diag_partial = nullptr;
    double diag_area = 0.0;
    // strip-marching kernel only: elements between rows of `pumping` (0 = cols).  The pumping rows are fetched by
    // TMA, whose row stride must be a multiple of 16 bytes: grids with an odd number of columns hand over a copy of
    // the pumping with an even pitch (api.cu makes it once per time loop)
    int p_pitch = 0;
    const PeerStep *peer = nullptr;   // strip-marching kernel only: this launch also performs the slab's halo exchange
};
// variant 0: 32x32 tiles, two CTAs per SM; variant 1: 32x64 tiles, one CTA of 512 threads per SM
int launch_rk4_step_fused_2d(int order, int variant, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream);
// stream_2d.cu: the same step as a strip-marching kernel (skewed stages, register windows, TMA row ring); reads
// and writes interleaved psi (grids with an odd number of columns are handed to the tile kernel: same bits).
int launch_rk4_step_stream_2d(int order, const Fused2DStep &s, const CrossWeights &w, cudaStream_t stream);
int stream_2d_plan(int order, int batch, int out_rows, int cols, int *threads, int *strips, int *chunk_rows);
bool stream_2d_takes(const Fused2DStep &s);
int stream_2d_diag_parts(int order, int batch, int out_rows, int cols);
void stream_2d_set_tuning(int sync, int width, int iters);
void stream_2d_get_tuning(int *t3);
// resident_2d.cu: `steps` RK4 steps of a SMALL grid in one cooperative launch, psi updated in place; the field
// lives in registers (one patch per CTA), neighbouring CTAs exchange edge nodes through `mailbox` (device scratch
// of the size resident_2d_query reports, zero-filled before its first use).  Packets carry sequence numbers
// seq0 + 1 ... seq0 + 4 steps: a mailbox may be reused by later launches with seq0 advanced accordingly.
struct Resident2D {
    int batch, rows, cols;
    int steps;
    double2 *psi;
    const double *pumping;
    const double *coeffs;       // [batch][23] on the device
    const RhsCoeffs *uniform;   // host: non-null when every member shares these coefficients
    void *mailbox;
    unsigned seq0;
    double dt;
};
int resident_2d_query(int order, const Resident2D &r, bool *fits, size_t *mailbox_bytes);
int launch_rk4_resident_2d(int order, const Resident2D &r, const CrossWeights &w, cudaStream_t stream);
// Planar (re-plane / im-plane) working copy of psi used by the whole-grid time loop: lets the fused
// kernel fetch its frames with TMA (cp.async.bulk.tensor, out-of-bounds zero fill = the truncated
// stencil boundary).  Layout: psi planes [batch][2][rows][pitch], c12*P [batch][rows][pitch], pitch even.
struct PlanarMaps {
    alignas(64) unsigned char psi_a[128];   // CUtensorMap of buffer A, buffer B, and of c12*P
    alignas(64) unsigned char psi_b[128];
    alignas(64) unsigned char cp[128];
};
struct Fused2DPlanar {
    int batch, rows, cols, pitch;
    int grow0, grows;        // slab decomposition: global row of local row 0, global row count (else 0, rows)
    int out_row0, out_row1;  // local rows to produce (else 0, rows)
    double *psi_a, *psi_b;   // two planar psi buffers (ping-pong)
    double *cp;              // c12 * P
    const double *coeffs;    // [batch][23] on the device
    const RhsCoeffs *uniform;
    double dt;
};
inline int planar_pitch(int cols) { return (cols + 1) & ~1; }
inline size_t planar_psi_doubles(int batch, int rows, int cols) { return (size_t)2 * batch * rows * planar_pitch(cols); }
inline size_t planar_cp_doubles(int batch, int rows, int cols) { return (size_t)batch * rows * planar_pitch(cols); }
int make_planar_maps(int order, int variant, const Fused2DPlanar &p, PlanarMaps *maps);
// psi (interleaved) -> planes of buffer A, and c12*P; buffer A/B -> psi (interleaved)
int launch_split_planar(const Fused2DPlanar &p, const double2 *psi, const double *pumping, cudaStream_t stream);
int launch_join_planar(const Fused2DPlanar &p, bool from_b, double2 *psi, cudaStream_t stream);
// one RK4 step A -> B (a_to_b) or B -> A.  overlap: the launch may begin before the previous kernel of the stream
// has drained (programmatic dependent launch; the kernel waits for it before touching psi) -- only valid when
// that previous kernel is the preceding step of the same time loop (c12*P is read before the wait).
int launch_rk4_step_fused_2d_planar(int order, int variant, const Fused2DPlanar &p, const PlanarMaps &maps, bool a_to_b,
                                    bool overlap, const CrossWeights &w, cudaStream_t stream);
// Persistent variant: `steps` RK4 steps (A -> B -> A ...) in ONE cooperative launch; usable when all tiles
// are resident at once (persistent_2d_fits).  flags: zeroed device ints, one per tile.
int persistent_2d_fits(int order, int variant, const Fused2DPlanar &p, bool *fits, long long *tiles);
int launch_rk4_persistent_2d_planar(int order, int variant, const Fused2DPlanar &p, const PlanarMaps &maps, int steps,
                                    int *flags, const CrossWeights &w, cudaStream_t stream);
int launch_cross_matvec_2d(int rows, int cols, int order, const CrossWeights &w, const double *x, double *y,
                           double sign, cudaStream_t stream);
// diagnostics.cu: fused scalar diagnostics of device-resident states (8 doubles per member, see nls_b200.h)
size_t diagnostics_scratch_bytes(int batch);
int launch_finish_diagnostics(int batch, int parts, const void *partial, double *out8, cudaStream_t stream);
int launch_diagnostics_2d(int batch, int rows, int cols, int order, double dx, const CrossWeights &w,
                          const double *pumping, const double *coeffs, const double2 *u, void *scratch, double *out8,
                          cudaStream_t stream);
int launch_diagnostics_1d(int batch, int n, int order, double dx, const double *taps, const double *pumping,
                          const double *coeffs, const double2 *u, void *scratch, double *out8, cudaStream_t stream);
// pumping_gen.cu: profiles of an ensemble generated on the device; params_dev: [batch][5] on the device
int launch_pumping_profiles(int dim, int kind, int batch, int n, double dx, const double *params_dev, double *out,
                            cudaStream_t stream);
int launch_divide_check(size_t n, const double *a, const double *b, double *fast, double *exact, cudaStream_t stream);
int launch_reservoir(size_t npts, RhsCoeffs c, const double *pumping, const double *u_sqr, double *r,
                     cudaStream_t stream);

}  // namespace nlsb
