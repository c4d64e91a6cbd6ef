/*
 * nls_b200.h -- C ABI of the B200 time-stepping engine for the daskol/nls hot path.
 *
 * The library (libnls_b200.so, built from nls_b200/csrc by nvcc for sm_100a) is the drop-in
 * boundary: its entry points are what the reference's FFI for this path binds.  In the reference
 * that FFI is the f2py extension `nls.native` generated from `module nls` of nls/nls.f90
 * (setup.py:69-79, nls/makefile:45-50); each "host" entry point below takes the place of the
 * Fortran routine cited next to it, with the same argument order and meaning.
 *
 * Conventions
 *   - Plain C: pointers and sizes only, no exceptions, no torch / C++ types in any signature.
 *   - Real kind: double.  The reference computes in real(sp) (nls.f90:10); this engine computes in
 *     float64 / complex128 from the float64 inputs the Python layer builds (documented divergence,
 *     DESIGN.md "Precision contract").  Complex arrays are interleaved (re, im) doubles.
 *   - 2D arrays are n x n with the FIRST index contiguous, exactly as the Fortran dummy arguments
 *     (numpy a[i, j] == Fortran a(i+1, j+1) == memory offset i + n*j).
 *   - `coeffs` has 23 entries (nls.f90:770-786); only Fortran entries 3,4,5,6,12,13,14 are read.
 *   - Return value: 0 on success; < 0 invalid argument (NLSB_E*); > 0 a cudaError_t.  A message for
 *     the calling thread's last failure is returned by nlsb_last_error().  Unlike the reference,
 *     which silently leaves the operator uninitialised for an unsupported order (nls.f90:293-294,
 *     :396-402), these entry points fail with NLSB_EORDER.
 *   - nlsb_* host entry points copy their inputs to the current CUDA device, run on an internal
 *     stream, copy the result back and synchronise before returning.  Nothing is retained between
 *     calls (the reference rebuilds its operator per call too, nls.f90:811, :917).
 *   - nlsb_dev_* entry points take DEVICE pointers and a cudaStream_t (as void*), enqueue work and
 *     return without synchronising.  They are what a device-resident caller (the Python layer with
 *     torch tensors, the ensemble and slab drivers) uses.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef NLS_B200_H
#define NLS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* every declaration below is an exported symbol of libnls_b200.so (the library itself is built with
 * -fvisibility=hidden) */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define NLSB_OK        0
#define NLSB_EINVAL   -1   /* null pointer / negative count */
#define NLSB_EORDER   -2   /* order not in {3, 5, 7} */
#define NLSB_ESIZE    -3   /* n too small for the stencil (n < order) or too large for the kernel */
#define NLSB_ENOMEM   -4   /* host allocation failed */
#define NLSB_EOPERATOR -5  /* operator table is not of the form the engine supports */

typedef void *nlsb_stream_t;   /* cudaStream_t */

const char *nlsb_last_error(void);
/* 1 if a CUDA device is usable by this process, else 0 (never raises). */
int nlsb_device_available(void);
/* Device scratch of the host-buffer entry points comes from a memory pool PRIVATE to this library and stays
 * cached there between calls, as do the instantiated CUDA graphs of the 2D time loop; this returns both to
 * the driver (call it when no nlsb_* work is in flight). */
int nlsb_trim_memory(void);
/* Number of CUDA kernels this library has launched since it was loaded (graph replays included). */
unsigned long long nlsb_kernel_launches(void);

/* ---- module nls, public routines (nls.f90:13-24) ------------------------------------------- */

/* nls.f90:29-37  subroutine version(major, minor, patch) */
void nlsb_version(int *major, int *minor, int *patch);

/* nls.f90:61-91  make_banded_matrix(n, m, row, mat): mat is (m, n) column-major BLAS band storage */
int nlsb_make_banded_matrix(int n, int m, const double *row, double *mat);
/* nls.f90:93-107 */
int nlsb_clear_first_row_of_derivative(int n, int m, double *L1);
/* nls.f90:109-130 */
int nlsb_divide_derivative_on_radius(int n, int m, double h, double *L1);
/* nls.f90:132-295  make_laplacian[_o3|_o5|_o7](n, [m,] h, op): radial operator, op is (m, n) band */
int nlsb_make_laplacian(int n, int m, double h, double *op);
int nlsb_make_laplacian_o3(int n, double h, double *op);
int nlsb_make_laplacian_o5(int n, double h, double *op);
int nlsb_make_laplacian_o7(int n, double h, double *op);
/* nls.f90:297-403  make_laplacian_2d[_o3|_o5|_o7](n, [m,] h, blocks, orders): blocks is (n, 2m-1).
 * For m = 7 the intended 13-point cross stencil is built (the reference's own storage for that
 * order is mis-dimensioned, nls.f90:360). */
int nlsb_make_laplacian_2d(int n, int m, double h, double *blocks, int *orders);
int nlsb_make_laplacian_2d_o3(int n, double h, double *blocks, int *orders);
int nlsb_make_laplacian_2d_o5(int n, double h, double *blocks, int *orders);
int nlsb_make_laplacian_2d_o7(int n, double h, double *blocks, int *orders);

/* nls.f90:530-541  rgbmv(x, u, sign, op, klu, n):  u := u + sign * A x,  A banded (2klu+1, n) */
int nlsb_rgbmv(const double *x, double *u, double sign, const double *op, int klu, int n);
/* nls.f90:408-527  rbbmv[_o3|_o5|_o7](x, y, sign, blocks, ms, [m,] n):  y := y + sign * A x on n*n.
 * Any blocks memory of the make_laplacian_2d layout (ms = (0, .., k, .., 0)) is accepted, here and in
 * hamiltonian_2d / runge_kutta_2d: constant-weight cross stencils (what make_laplacian_2d builds) take the fast
 * kernels, weights that vary along the line take general kernels that read the blocks memory on the device.
 * Other `ms` (which the reference's hard-wired block offsets cannot address either) fail with NLSB_EOPERATOR. */
int nlsb_rbbmv(const double *x, double *y, double sign, const double *blocks, const int *ms, int m, int n);
int nlsb_rbbmv_o3(const double *x, double *y, double sign, const double *blocks, const int *ms, int n);
int nlsb_rbbmv_o5(const double *x, double *y, double sign, const double *blocks, const int *ms, int n);
int nlsb_rbbmv_o7(const double *x, double *y, double sign, const double *blocks, const int *ms, int n);

/* nls.f90:570-581 / :829-839  revervoir[_2d](pumping, coeffs, u_sqr, r, n) */
int nlsb_revervoir(const double *pumping, const double *coeffs, const double *u_sqr, double *r, int n);
int nlsb_revervoir_2d(const double *pumping, const double *coeffs, const double *u_sqr, double *r, int n);

/* nls.f90:621-650  hamiltonian(pumping, coeffs, u, v, op, klu, n) */
int nlsb_hamiltonian(const double *pumping, const double *coeffs, const double *u, double *v,
                     const double *op, int klu, int n);
/* nls.f90:841-870  hamiltonian_2d(pumping, coeffs, u, v, blocks, orders, order, n) */
int nlsb_hamiltonian_2d(const double *pumping, const double *coeffs, const double *u, double *v,
                        const double *blocks, const int *orders, int order, int n);

/* nls.f90:705-734  runge_kutta(dt, t0, u0, op, n, order, iters, u, pumping, coeffs) */
int nlsb_runge_kutta(double dt, double t0, const double *u0, const double *op, int n, int order, int iters,
                     double *u, const double *pumping, const double *coeffs);
/* nls.f90:873-901  runge_kutta_2d(dt, t0, u0, n, blocks, orders, order, iters, u, pumping, coeffs) */
int nlsb_runge_kutta_2d(double dt, double t0, const double *u0, int n, const double *blocks, const int *orders,
                        int order, int iters, double *u, const double *pumping, const double *coeffs);

/* nls.f90:797-813, :815-827, :903-919  solve_nls[_1d|_2d](dt, dx, n, order, iters, pumping, coeffs, u0, u) */
int nlsb_solve_nls(double dt, double dx, int n, int order, int iters, const double *pumping,
                   const double *coeffs, const double *u0, double *u);
int nlsb_solve_nls_1d(double dt, double dx, int n, int order, int iters, const double *pumping,
                      const double *coeffs, const double *u0, double *u);
int nlsb_solve_nls_2d(double dt, double dx, int n, int order, int iters, const double *pumping,
                      const double *coeffs, const double *u0, double *u);

/* nls.f90:921-948  chemical_potential_1d(dx, n, pumping, coeffs, u0, mu): mu is complex (2 doubles) */
int nlsb_chemical_potential_1d(double dx, int n, const double *pumping, const double *coeffs,
                               const double *u0, double *mu);
/* nls.f90:950-971  chemical_potential_2d(dx, n, pumping, coeffs, u0, mu): mu is real */
int nlsb_chemical_potential_2d(double dx, int n, const double *pumping, const double *coeffs,
                               const double *u0, double *mu);

/* ---- host-side operator tables in the engine's layout -------------------------------------- */

/* Radial operator as a row-major tap table taps[i*m + t], t = s + k, multiplying x[i + s]
 * (same numbers as make_laplacian, transposed out of band storage). */
int nlsb_radial_taps(int n, int m, double h, double *taps);
/* Band (m, n) -> tap table (n, m). */
int nlsb_band_to_taps(int n, int m, const double *op, double *taps);
/* Weights of the 2D cross stencil: wx[t] along the contiguous index (centre included), wy[t]
 * across lines (wy[k] = 0), t = s + k. */
int nlsb_cross_weights(int m, double h, double *wx, double *wy);
/* Extract (wx, wy) from a block-band operator; NLSB_EOPERATOR if it is not line-independent. */
int nlsb_blocks_to_weights(int n, int m, const double *blocks, const int *orders, double *wx, double *wy);

/* ---- device-resident entry points (device pointers, asynchronous on `stream`) --------------- */

/* Batched 1D radial systems sharing one tap table.  taps: [n][order]; pumping: [batch][n];
 * coeffs: [batch][23]; psi: [batch][n] complex, advanced in place by `iters` RK4 steps.
 * Each system lives in one CTA's registers / shared memory for the whole time loop. n <= 8192. */
int nlsb_dev_rk4_1d(int batch, int n, int order, int iters, double dt, const double *taps,
                    const double *pumping, const double *coeffs, double *psi, nlsb_stream_t stream);
/* v = H(u) for a batch of 1D systems (one RHS evaluation, nls.f90:621-650). */
int nlsb_dev_hamiltonian_1d(int batch, int n, int order, const double *taps, const double *pumping,
                            const double *coeffs, const double *u, double *v, nlsb_stream_t stream);
/* u := u + sign * A x, real vectors of length n. */
int nlsb_dev_band_matvec_1d(int n, int order, const double *taps, const double *x, double *u, double sign,
                            nlsb_stream_t stream);

/* Batched 2D grids of rows x cols points (cols contiguous).  wx, wy: HOST arrays of `order` weights.
 * pumping: [batch][rows][cols]; coeffs: DEVICE [batch][23]; psi: [batch][rows][cols] complex, in place.
 * shared_coeffs_host: HOST copy of the 23 coefficients when every member uses the same ones (lets the
 * kernel read them from its constant bank), else NULL.
 * workspace: device scratch of nlsb_dev_rk4_2d_workspace() bytes. */
size_t nlsb_dev_rk4_2d_workspace(int batch, int rows, int cols);
/* Which kernels advance 2D grids: 0 = automatic (one launch per RK step: the strip-marching kernel for
 * launches of >= 2^20 nodes, else the tile kernel with 32x32 tiles filled by TMA from a planar working
 * copy), 1 = one launch per RK stage, 2 / 3 = tile kernel with 32x32 / 32x64 tiles filled by plain
 * loads, 4 / 5 = tile kernel with 32x32 / 32x64 tiles filled by TMA, 6 / 7 = as 4 / 5 but the whole time
 * loop in one persistent launch when every tile is resident at once (experimental, slower), 8 = the
 * strip-marching kernel whatever the size, 9 = register-resident kernel, whole time loop in one launch
 * with per-stage edge exchange through L2 (experimental, slower; falls back to 4 when the grid does not
 * fit).  All give the same result to rounding (2..9 bitwise); the switch exists for tests and profiling. */
int nlsb_set_2d_path(int path);
/* Tuning of the strip-marching kernel (tests, profiling): sync = how the threads of a CTA order their shared-memory
 * traffic (0 one barrier per row, 1 one barrier per two rows; -1 = library default),
 * width = threads per strip (one of the compiled widths; 0 = automatic), iters_per_cta = rows a CTA marches
 * (0 = automatic).  Every setting produces the same bits. */
int nlsb_set_stream_tuning(int sync, int width, int iters_per_cta);
/* How nlsb_solve_nls_2d (nls.f90:903-919, host buffers) schedules a solve: overlapped != 0 when the transfers are
 * overlapped with the first / last steps (grids >= 2048^2, >= 160 steps, automatic or strip-marching kernels).  Start:
 * rows [0, r_top + 4k s_up) are uploaded first and the range [0, r_top + 4k (s_up - j)) takes step j while the other
 * rows are on the bus, then the complementary rows catch up; end: the range [0, r_dn + 4k (s_dn - j)) takes step j of the
 * last s_dn steps first and rows [0, r_dn) leave for the host while the rows below finish.  Host arithmetic only. */
int nlsb_solve_nls_2d_plan(int n, int order, int iters, int *overlapped, int *r_top, int *s_up, int *r_dn, int *s_dn);
int nlsb_dev_rk4_2d(int batch, int rows, int cols, int order, int iters, double dt, const double *wx,
                    const double *wy, const double *pumping, const double *coeffs,
                    const double *shared_coeffs_host, double *psi,
                    void *workspace, size_t workspace_bytes, nlsb_stream_t stream);
/* Which kernel nlsb_dev_rk4_2d would use for this problem under the current nlsb_set_2d_path, and its launch
 * geometry (host arithmetic only, needs no device): *kernel = 0 tile kernel 32x32, 1 tile kernel 32x64,
 * 2 strip-marching kernel (then *threads per CTA, *strips of columns, *chunk_rows per CTA), 3 per-stage kernels,
 * 4 register-resident kernel (if the grid fits, else 0). */
int nlsb_dev_rk4_2d_plan(int batch, int rows, int cols, int order, int *kernel, int *threads, int *strips,
                         int *chunk_rows);
/* One whole RK4 step of ONE slab of a 2D grid that is decomposed along its slow (row) axis.
 * psi_in / psi_out / pumping are local arrays of rows_alloc x cols nodes (distinct in/out buffers)
 * whose row 0 is global row `global_row0` (negative when the slab starts with halo rows above the
 * domain); local rows [out_row0, out_row1) of psi_out are produced.  Every node read within 4*k rows of
 * the output rows must hold the current psi (k = (order-1)/2): the caller exchanges those halo rows with
 * its neighbours between steps.  Rows outside [0, global_rows) are treated as zero.  coeffs_host: 23
 * coefficients on the HOST.  A node's result does not depend on how the grid is cut into slabs. */
int nlsb_dev_rk4_step_2d_slab(int rows_alloc, int cols, int order, double dt, const double *wx, const double *wy,
                              int global_row0, int global_rows, int out_row0, int out_row1, const double *pumping,
                              const double *coeffs_host, const double *psi_in, double *psi_out, nlsb_stream_t stream);
/* The same slab step on the PLANAR layout the TMA-fed kernel reads: planes_in / planes_out hold the
 * real plane then the imaginary plane, each rows_alloc x pitch doubles with pitch = nlsb_planar_pitch(cols)
 * (cols rounded up to even, so rows are 16-byte aligned); cp is c12*P in the same rows_alloc x pitch
 * layout.  Padding columns and rows outside the square must be zero in planes_in (the kernel never
 * writes them).  Tiles are fetched with cp.async.bulk.tensor; elements outside the local arrays arrive
 * as zeros. */
int nlsb_planar_pitch(int cols);
int nlsb_dev_rk4_step_2d_slab_planar(int rows_alloc, int cols, int order, double dt, const double *wx,
                                     const double *wy, int global_row0, int global_rows, int out_row0, int out_row1,
                                     const double *cp, const double *coeffs_host, double *planes_in, double *planes_out,
                                     nlsb_stream_t stream);
int nlsb_dev_hamiltonian_2d(int batch, int rows, int cols, int order, const double *wx, const double *wy,
                            const double *pumping, const double *coeffs, const double *u, double *v,
                            nlsb_stream_t stream);
/* y := y + sign * A x, real fields of rows x cols. */
int nlsb_dev_cross_matvec_2d(int rows, int cols, int order, const double *wx, const double *wy,
                             const double *x, double *y, double sign, nlsb_stream_t stream);
/* r = c12 P / (c13 + c14 u_sqr) on npts points (coeffs on the HOST). */
int nlsb_dev_reservoir(size_t npts, const double *coeffs_host, const double *pumping, const double *u_sqr,
                       double *r, nlsb_stream_t stream);

/* Scalar diagnostics of device-resident states, one pass per member, nothing copied back but 8 doubles per
 * member.  Replaces, for a state that lives on the device, what the reference computes on the host from the
 * returned field: chemical_potential_1d/_2d (nls.f90:921-971; mu = i E / M with the MODEL's stencil order --
 * the reference hard-wires 5), Solution.getDampingIntegral / getDensity / getReservoir (nls/model.py:350-380).
 * out8: DEVICE, [batch][8] = {Re M, Im M, Re E, Im E, damping integral, particle number, max |psi|^2,
 * max reservoir}, M = sum w conj(u) u, E = sum w conj(u) H(u), w = i*dx (1D) or 1 (2D); area element 2 pi r dx
 * with r = linspace(0, n dx, n) (1D) or dx^2 (2D).  scratch: DEVICE, nlsb_dev_diagnostics_scratch(batch) bytes.
 * Fixed reduction tree: results are reproducible run to run. */
size_t nlsb_dev_diagnostics_scratch(int batch);
int nlsb_dev_diagnostics_1d(int batch, int n, int order, double dx, const double *taps, const double *pumping,
                            const double *coeffs, const double *psi, void *scratch, double *out8,
                            nlsb_stream_t stream);
int nlsb_dev_diagnostics_2d(int batch, int rows, int cols, int order, double dx, const double *wx,
                            const double *wy, const double *pumping, const double *coeffs, const double *psi,
                            void *scratch, double *out8, nlsb_stream_t stream);
/* Time loop with the diagnostics FUSED into it: as nlsb_dev_rk4_1d / nlsb_dev_rk4_2d, and
 * out8[member][8] receives the diagnostics above of the state ENTERING the last of the `iters` (>= 1) steps -- that
 * step's first stage holds H(psi) of that state, so the reduction rides in the same launch (1D: CTA-resident kernel;
 * 2D: strip-marching kernel, per-CTA partial sums in diag_scratch + a fixed-order finishing launch).  The value is
 * what nlsb_dev_diagnostics_* returns for the state after iters - 1 steps; the convergence loop of tools/check.py
 * (solve a chunk, copy back, chemical potential) becomes one call per chunk with 8 doubles per member read back.
 * Kernels without the fused reduction (tile kernel on small grids, per-stage paths, 1D n > 2048) run the stand-alone
 * pass before their last step: same result.  diag_scratch: nlsb_dev_rk4_2d_diag_scratch() bytes (2D) or
 * nlsb_dev_diagnostics_scratch() bytes (1D). */
size_t nlsb_dev_rk4_2d_diag_scratch(int batch, int rows, int cols, int order);
int nlsb_dev_rk4_2d_diag(int batch, int rows, int cols, int order, int iters, double dt, double dx, const double *wx,
                         const double *wy, const double *pumping, const double *coeffs,
                         const double *shared_coeffs_host, double *psi, void *workspace, size_t workspace_bytes,
                         void *diag_scratch, double *out8, nlsb_stream_t stream);
int nlsb_dev_rk4_1d_diag(int batch, int n, int order, int iters, double dt, double dx, const double *taps,
                         const double *pumping, const double *coeffs, double *psi, void *diag_scratch, double *out8,
                         nlsb_stream_t stream);


/* Pumping profiles of an ensemble generated on the device, sampled on the reference's grid (nls/model.py:220-232:
 * dim 1: x = linspace(0, n dx, n); dim 2: x = y = linspace(-n dx/2, n dx/2, n), out[b][i][j] = P(x_j, y_i)).
 * kind 0: GaussianPumping[1D|2D], kind 1: GaussianRingPumping[1D|2D] (nls/pumping.py:113-178).
 * params_host: HOST [batch][5] = {power, x0, y0, variation, radius}; out: DEVICE [batch][n] or [batch][n][n].
 * Grid and arithmetic are bit-identical to numpy's, exp() may differ in the last place (<= 4 ulp in the result). */
int nlsb_dev_pumping_profiles(int dim, int kind, int batch, int n, double dx, const double *params_host,
                              double *out, nlsb_stream_t stream);

/* Self-test hook: fast[i] = a[i] / b[i] by the divide of the fused kernels (reciprocal seed + Newton step + residual
 * correction, csrc/device_math.cuh), exact[i] = the correctly rounded IEEE quotient; device pointers.  The reservoir
 * n = c12 P / (c13 + c14 |psi|^2) (nls.f90:580) is the only division on the path; its denominator is >= c13 > 0. */
int nlsb_dev_divide_check(size_t n, const double *a, const double *b, double *fast, double *exact, nlsb_stream_t stream);

/* ---- multi-GPU row slabs: peer-mapped buffers and the device-initiated halo exchange -----------------------
 * The reference has no parallelism; BASELINE.json config 4 asks for the 8192^2 grid cut into row
 * slabs over the GPUs of one box (nls_b200/multigpu.py).  One process per GPU.
 * nlsb_peer_alloc: zero-filled device memory whose allocation can be exported (cudaMalloc + cudaIpcGetMemHandle);
 * nlsb_peer_export writes the 64-byte IPC handle, nlsb_peer_open maps another process's allocation into this one
 * (peer access over NVLink is enabled on first use), nlsb_peer_close unmaps it. */
int nlsb_peer_alloc(size_t bytes, void **ptr);
int nlsb_peer_free(void *ptr);
int nlsb_peer_export(const void *ptr, unsigned char *handle64);
int nlsb_peer_open(const unsigned char *handle64, void **ptr);
int nlsb_peer_close(void *ptr);
int nlsb_peer_enable_access(int peer_device);
/* One halo exchange of a slab, enqueued as ONE kernel on `stream`, no host work (capturable in a CUDA graph):
 * the kernel tells both neighbours that this rank's halo rows may be overwritten, waits for the same word from
 * them, copies `complex_count` complex128 values from src_up / src_down (this rank's boundary rows) to dst_up /
 * dst_down (the neighbours' halo rows, peer-mapped pointers), publishes "data of epoch e" in the neighbours' flag
 * blocks and returns when the neighbours' data of the same epoch has arrived in this rank's halo rows.
 * state: 64 zero-initialised bytes of this rank's memory (epoch counter, ticket, time-out count);
 * flags_mine: this rank's flag block (256 zero-initialised bytes from nlsb_peer_alloc); flags_up / flags_down: the
 * neighbours' flag blocks (peer-mapped) or NULL at the ends of the chain.  A wait gives up after timeout_seconds
 * (<= 0: 2 s) and is then counted in the state block (nlsb_dev_halo_status) instead of hanging the device. */
int nlsb_dev_halo_exchange(const double *src_up, double *dst_up, const double *src_down, double *dst_down,
                           size_t complex_count, void *state, void *flags_mine, void *flags_up, void *flags_down,
                           double timeout_seconds, nlsb_stream_t stream);
/* One RK4 step of a slab AND its halo exchange in ONE launch (strip-marching kernel): as nlsb_dev_rk4_step_2d_slab,
 * and the `halo_rows` new rows starting at local row up_row0 (dn_row0) are also stored, from the kernel's own store
 * stage, into the halo rows of the rank above (below) beginning at up_dst (dn_dst) -- peer-mapped pointers into the
 * neighbour's psi_out buffer.  The launch publishes READY at its start, the CTAs that write into a neighbour wait for
 * that neighbour's READY, the last CTA publishes DATA and waits for the neighbours' DATA (protocol and blocks as for
 * nlsb_dev_halo_exchange, with which it shares them).  Fails with NLSB_EINVAL when the launch would not take the
 * strip-marching kernel (small slabs, odd column counts): use the step + nlsb_dev_halo_exchange then. */
int nlsb_dev_rk4_step_2d_slab_exchange(int rows_alloc, int cols, int order, double dt, const double *wx, const double *wy,
                                       int global_row0, int global_rows, int out_row0, int out_row1,
                                       const double *pumping, const double *coeffs_host, const double *psi_in,
                                       double *psi_out, int halo_rows, int up_row0, double *up_dst, int dn_row0,
                                       double *dn_dst, void *state, void *flags_mine, void *flags_up, void *flags_down,
                                       double timeout_seconds, nlsb_stream_t stream);
/* Synchronous read-back of a state block: exchanges completed and waits that timed out. */
int nlsb_dev_halo_status(const void *state, unsigned long long *epoch, unsigned long long *timeouts);
/* Adds to the count reported by nlsb_kernel_launches(): kernels replayed from a CUDA graph the CALLER captured
 * are launched without passing through the library. */
void nlsb_add_kernel_launches(unsigned long long n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* NLS_B200_H */
