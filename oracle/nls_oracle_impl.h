/*
 * nls_oracle_impl.h -- kind-generic body of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is included twice by nls_oracle.c, once with REAL=float (suffix _sp: the precision the
 * reference ships, nls/nls.f90:10) and once with REAL=double (suffix _dp: the kind-promoted
 * restatement that the 1e-10 parity bar of the engine is measured against).
 *
 * It is a restatement, in C, of the algorithm of the reference's `module nls`
 * (/root/reference/nls/nls.f90).  Every routine cites the lines it follows.  Arithmetic is written
 * in the reference's evaluation order; the file must be compiled with -ffp-contract=off so that the
 * compiler does not fuse a*b+c (gfortran -O3 without -ffast-math on x86-64 baseline does not
 * either).
 *
 * Array conventions: all matrices are Fortran (column-major) arrays, exactly as the reference
 * declares them; the macro F2(a, ld, i, j) addresses element (i, j) with 1-based indices.
 *
 * Third-party arithmetic: the reference calls BLAS-2 `sgbmv` from an unpinned system -lblas
 * (nls.f90:540, nls/makefile:17).  BLAS is not part of /root/reference, so `gbmv` below restates
 * the published netlib reference algorithm (LAPACK 3.x SGBMV, TRANS='N', unit increments): a
 * column sweep of AXPYs into y.
 */

#ifndef REAL
#error "include from nls_oracle.c"
#endif

#define F2(a, ld, i, j) ((a)[((size_t)(i) - 1) + (size_t)(ld) * ((size_t)(j) - 1)])

typedef struct { REAL re, im; } KIND(cplx);

/* nls.f90:61-91.  Band storage of a Toeplitz row: column j holds the stencil `row`, with the
 * entries that would fall outside the n x n matrix zeroed in the first and last k columns. */
static void KIND(banded)(int n, int m, const REAL *row, REAL *mat)
{
    int k = (m - 1) / 2;
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= m; ++i)
            F2(mat, m, i, j) = row[i - 1];
    /* left triangle, nls.f90:76-79: rows 1..k+1-j of column j are outside the matrix */
    for (int j = 1; j <= k && j <= n; ++j)
        for (int i = 1; i <= k + 1 - j; ++i)
            F2(mat, m, i, j) = 0;
    /* right triangle, nls.f90:87-90: rows m-j+1..m of column n-k+j are outside the matrix */
    for (int j = 1; j <= k; ++j) {
        int col = n - k + j;
        if (col < 1) continue;
        for (int i = m - j + 1; i <= m; ++i)
            F2(mat, m, i, col) = 0;
    }
}

NLSO_API void KIND(nlso_make_banded_matrix)(int n, int m, const REAL *row, REAL *mat)
{
    KIND(banded)(n, m, row, mat);
}

/* nls.f90:93-107: zero matrix row 0 of the first-derivative band (band entries (i, k+2-i)). */
NLSO_API void KIND(nlso_clear_first_row_of_derivative)(int n, int m, REAL *L1)
{
    (void)n;
    int mid = (m + 1) / 2;
    for (int i = 1; i <= mid; ++i)
        F2(L1, m, i, mid - i + 1) = 0;
}

/* nls.f90:109-130: divide matrix row rho >= 1 of the band by rho*h (0-based row = i + j - (m+3)/2). */
NLSO_API void KIND(nlso_divide_derivative_on_radius)(int n, int m, REAL h, REAL *L1)
{
    for (int i = 1; i <= m; ++i)
        for (int j = 1; j <= n; ++j) {
            int rho = i + j - (m + 3) / 2;
            if (rho > 0)
                F2(L1, m, i, j) = F2(L1, m, i, j) / ((REAL)rho * h);
        }
}

/* Shared tail of make_laplacian_o{3,5,7}: D1 fix-ups, radius division, op = L1 + L2. */
static void KIND(finish_radial)(int n, int m, REAL h, REAL *L1, REAL *L2, REAL *op)
{
    KIND(nlso_clear_first_row_of_derivative)(n, m, L1);
    KIND(nlso_divide_derivative_on_radius)(n, m, h, L1);
    for (size_t t = 0; t < (size_t)m * (size_t)n; ++t)
        op[t] = L1[t] + L2[t];
}

/* nls.f90:132-161 */
NLSO_API int KIND(nlso_make_laplacian_o3)(int n, REAL h, REAL *op)
{
    enum { m = 3 };
    REAL dx1 = 2 * h, dx2 = 1 * (h * h);
    static const int s1[m] = {1, 0, -1}, s2[m] = {1, -2, 1};
    REAL r1[m], r2[m];
    for (int i = 0; i < m; ++i) { r1[i] = (REAL)s1[i] / dx1; r2[i] = (REAL)s2[i] / dx2; }
    REAL *L1 = (REAL *)malloc(sizeof(REAL) * m * (size_t)n), *L2 = (REAL *)malloc(sizeof(REAL) * m * (size_t)n);
    if (!L1 || !L2) { free(L1); free(L2); return -2; }
    KIND(banded)(n, m, r1, L1);
    KIND(banded)(n, m, r2, L2);
    F2(L2, m, 2, 1) = 2 * F2(L2, m, 2, 1);   /* nls.f90:154 */
    F2(L2, m, 1, 2) = 4 * F2(L2, m, 1, 2);   /* nls.f90:155 */
    KIND(finish_radial)(n, m, h, L1, L2, op);
    free(L1); free(L2);
    return 0;
}

/* nls.f90:163-195.  `legacy24` != 0 switches the second-derivative denominator to 24 h^2, the
 * older revision that the stale golden table of test/test_nls.f95:102-108 pins. */
static int KIND(laplacian_o5)(int n, REAL h, REAL *op, int legacy24)
{
    enum { m = 5 };
    REAL dx1 = 12 * h, dx2 = (legacy24 ? 24 : 12) * (h * h);
    static const int s1[m] = {-1, 8, 0, -8, 1}, s2[m] = {-1, 16, -30, 16, -1};
    REAL r1[m], r2[m];
    for (int i = 0; i < m; ++i) { r1[i] = (REAL)s1[i] / dx1; r2[i] = (REAL)s2[i] / dx2; }
    REAL *L1 = (REAL *)malloc(sizeof(REAL) * m * (size_t)n), *L2 = (REAL *)malloc(sizeof(REAL) * m * (size_t)n);
    if (!L1 || !L2) { free(L1); free(L2); return -2; }
    KIND(banded)(n, m, r1, L1);
    KIND(banded)(n, m, r2, L2);
    F2(L1, m, 3, 2) = F2(L1, m, 3, 2) + (REAL)1.0 / dx1;   /* nls.f90:185 */
    F2(L2, m, 3, 2) = F2(L2, m, 3, 2) - (REAL)1.0 / dx2;   /* nls.f90:186 */
    F2(L2, m, 3, 1) = 2 * F2(L2, m, 3, 1);                 /* nls.f90:187 */
    F2(L2, m, 2, 2) = 4 * F2(L2, m, 2, 2);                 /* nls.f90:188 */
    F2(L2, m, 1, 3) = 4 * F2(L2, m, 1, 3);                 /* nls.f90:189 */
    KIND(finish_radial)(n, m, h, L1, L2, op);
    free(L1); free(L2);
    return 0;
}

NLSO_API int KIND(nlso_make_laplacian_o5)(int n, REAL h, REAL *op) { return KIND(laplacian_o5)(n, h, op, 0); }
NLSO_API int KIND(nlso_make_laplacian_o5_legacy24)(int n, REAL h, REAL *op) { return KIND(laplacian_o5)(n, h, op, 1); }

/* nls.f90:197-259 (the self-assignments at :220, :223-227, :229-232, :241, :244-248, :250-253 are no-ops) */
NLSO_API int KIND(nlso_make_laplacian_o7)(int n, REAL h, REAL *op)
{
    enum { m = 7 };
    REAL dx1 = 60 * h, dx2 = 180 * (h * h);
    static const int s1[m] = {1, -9, 45, 0, -45, 9, -1}, s2[m] = {2, -27, 270, -490, 270, -27, 2};
    REAL r1[m], r2[m];
    for (int i = 0; i < m; ++i) { r1[i] = (REAL)s1[i] / dx1; r2[i] = (REAL)s2[i] / dx2; }
    REAL *L1 = (REAL *)malloc(sizeof(REAL) * m * (size_t)n), *L2 = (REAL *)malloc(sizeof(REAL) * m * (size_t)n);
    if (!L1 || !L2) { free(L1); free(L2); return -2; }
    KIND(banded)(n, m, r1, L1);
    KIND(banded)(n, m, r2, L2);
    F2(L1, m, 4, 2) = F2(L1, m, 4, 2) + (REAL)9.0 / dx1;    /* nls.f90:221 */
    F2(L1, m, 3, 3) = F2(L1, m, 3, 3) - (REAL)1.0 / dx1;    /* nls.f90:222 */
    F2(L1, m, 5, 2) = F2(L1, m, 5, 2) - (REAL)1.0 / dx1;    /* nls.f90:228 */
    F2(L2, m, 4, 1) = 2 * F2(L2, m, 4, 1);                  /* nls.f90:235 */
    F2(L2, m, 3, 2) = 4 * F2(L2, m, 3, 2);                  /* nls.f90:236 */
    F2(L2, m, 2, 3) = 4 * F2(L2, m, 2, 3);                  /* nls.f90:237 */
    F2(L2, m, 1, 4) = 4 * F2(L2, m, 1, 4);                  /* nls.f90:238 */
    F2(L2, m, 4, 2) = F2(L2, m, 4, 2) - (REAL)27.0 / dx2;   /* nls.f90:242 */
    F2(L2, m, 3, 3) = F2(L2, m, 3, 3) + (REAL)2.0 / dx2;    /* nls.f90:243 */
    F2(L2, m, 5, 2) = F2(L2, m, 5, 2) + (REAL)2.0 / dx2;    /* nls.f90:249 */
    KIND(finish_radial)(n, m, h, L1, L2, op);
    free(L1); free(L2);
    return 0;
}

/* nls.f90:278-295.  The reference silently leaves `op` undefined for other orders (:293-294);
 * the oracle reports -1 instead of returning garbage. */
NLSO_API int KIND(nlso_make_laplacian)(int n, int m, REAL h, REAL *op)
{
    if (m == 3) return KIND(nlso_make_laplacian_o3)(n, h, op);
    if (m == 5) return KIND(nlso_make_laplacian_o5)(n, h, op);
    if (m == 7) return KIND(nlso_make_laplacian_o7)(n, h, op);
    return -1;
}

/* 2D block-band operator, nls.f90:297-403.  `blocks` is the reference's (n, 2m-1) array used as a
 * flat buffer: k one-row "diagonal" bands of n entries, the (m, n) middle band, k more diagonals.
 * For m = 7 the reference mis-declares the array as (m, 2m-1) (nls.f90:360) which is undefined
 * behaviour; the oracle builds the intended (n, 13) layout with the weights of nls.f90:368-374. */
NLSO_API int KIND(nlso_make_laplacian_2d)(int n, int m, REAL h, REAL *blocks, int *orders)
{
    REAL dx2;
    int w[7], k = (m - 1) / 2;           /* integer stencil numerators, index 0 = offset -k */
    if (m == 3) {                         /* nls.f90:310-320 */
        dx2 = (REAL)1.0 * (h * h);
        w[0] = 1; w[1] = -4; w[2] = 1;
    } else if (m == 5) {                  /* nls.f90:336-350 */
        dx2 = 12 * (h * h);
        w[0] = -1; w[1] = 16; w[2] = -60; w[3] = 16; w[4] = -1;
    } else if (m == 7) {                  /* nls.f90:366-384 */
        dx2 = (REAL)180.0 * (h * h);
        w[0] = 2; w[1] = -27; w[2] = 270; w[3] = -980; w[4] = 270; w[5] = -27; w[6] = 2;
    } else {
        return -1;                        /* nls.f90:396-402 falls through silently */
    }
    REAL middle[7];
    for (int i = 0; i < m; ++i) middle[i] = (REAL)w[i] / dx2;
    REAL *p = blocks;
    for (int b = 0; b < m; ++b) {
        if (b == k) {
            KIND(banded)(n, m, middle, p);
            orders[b] = k;
            p += (size_t)m * (size_t)n;
        } else {
            REAL one = (REAL)w[b] / dx2;  /* left/right one-element rows, e.g. nls.f90:338-342 */
            KIND(banded)(n, 1, &one, p);
            orders[b] = 0;
            p += n;
        }
    }
    return 0;
}

/* Netlib SGBMV/DGBMV, TRANS='N', m=n, kl=ku=klu, incx=incy=1, beta=1 -- the only way the
 * reference calls it (nls.f90:540):  y := alpha*A*x + y. */
static void KIND(gbmv)(int n, int klu, REAL alpha, const REAL *a, int lda, const REAL *x, REAL *y)
{
    if (n == 0 || alpha == 0) return;
    int kup1 = klu + 1;
    for (int j = 1; j <= n; ++j) {
        REAL temp = alpha * x[j - 1];
        int k = kup1 - j;
        int lo = j - klu > 1 ? j - klu : 1;
        int hi = j + klu < n ? j + klu : n;
        for (int i = lo; i <= hi; ++i)
            y[i - 1] = y[i - 1] + temp * F2(a, lda, k + i, j);
    }
}

/* nls.f90:530-541 */
NLSO_API void KIND(nlso_rgbmv)(const REAL *x, REAL *u, REAL sign, const REAL *op, int klu, int n)
{
    KIND(gbmv)(n, klu, sign, op, 2 * klu + 1, x, u);
}

/* nls.f90:408-527.  One rgbmv per grid line and block; the accumulation order per order m is the
 * reference's: o3 (-1, 0, +1), o5 (+2, +1, 0, -1, -2), o7 (-3 .. +3). */
NLSO_API int KIND(nlso_rbbmv)(const REAL *x, REAL *y, REAL sign, const REAL *blocks, const int *ms, int m, int n)
{
    int k = (m - 1) / 2;
    if (m != 3 && m != 5 && m != 7) return -1;
    /* block b (0-based, offset s = b - k: source line = line + s) starts at: */
    const REAL *blk[7];
    const REAL *p = blocks;
    for (int b = 0; b < m; ++b) { blk[b] = p; p += (b == k) ? (size_t)m * (size_t)n : (size_t)n; }
    int seq[7];
    if (m == 5) { int t[5] = {4, 3, 2, 1, 0}; for (int i = 0; i < 5; ++i) seq[i] = t[i]; }   /* nls.f90:447-465 */
    else        { for (int i = 0; i < m; ++i) seq[i] = i; }                                   /* :421-431, :481-507 */
    for (int q = 0; q < m; ++q) {
        int b = seq[q], s = b - k;
        for (int i = 1; i <= n; ++i) {           /* destination line i, source line i + s */
            int src = i + s;
            if (src < 1 || src > n) continue;
            KIND(nlso_rgbmv)(x + (size_t)(src - 1) * n, y + (size_t)(i - 1) * n, sign, blk[b], ms[b], n);
        }
    }
    return 0;
}

/* nls.f90:570-581 and :829-839 (same expression on n or n*n points). */
NLSO_API void KIND(nlso_revervoir)(const REAL *pumping, const REAL *coeffs, const REAL *u_sqr, REAL *r, size_t npts)
{
    for (size_t t = 0; t < npts; ++t)
        r[t] = coeffs[11] * pumping[t] / (coeffs[12] + coeffs[13] * u_sqr[t]);
}

/* Pointwise part of the right-hand side, nls.f90:637-644 / :857-864; work arrays are heap, not the
 * reference's automatic (stack) arrays. */
static int KIND(rhs_pointwise)(const REAL *pumping, const REAL *coeffs, const KIND(cplx) *u, size_t npts,
                               REAL *ur, REAL *ui, REAL *vr, REAL *vi, REAL *r, REAL *usq)
{
    for (size_t t = 0; t < npts; ++t) {
        ur[t] = u[t].re;
        ui[t] = u[t].im;
        /* real(conjg(u) * u): (re, -im) * (re, im) -> re*re - (-im)*im */
        usq[t] = u[t].re * u[t].re - (-u[t].im) * u[t].im;
    }
    KIND(nlso_revervoir)(pumping, coeffs, usq, r, npts);
    for (size_t t = 0; t < npts; ++t) {
        vr[t] = (coeffs[2] * r[t] - coeffs[3]) * ur[t] + (coeffs[4] * usq[t] + coeffs[5] * r[t]) * ui[t];
        vi[t] = (coeffs[2] * r[t] - coeffs[3]) * ui[t] - (coeffs[4] * usq[t] + coeffs[5] * r[t]) * ur[t];
    }
    return 0;
}

/* nls.f90:621-650 */
NLSO_API int KIND(nlso_hamiltonian)(const REAL *pumping, const REAL *coeffs, const KIND(cplx) *u, KIND(cplx) *v,
                                    const REAL *op, int klu, int n)
{
    size_t np = (size_t)n;
    REAL *w = (REAL *)malloc(sizeof(REAL) * 6 * np);
    if (!w) return -2;
    REAL *ur = w, *ui = w + np, *vr = w + 2 * np, *vi = w + 3 * np, *r = w + 4 * np, *usq = w + 5 * np;
    KIND(rhs_pointwise)(pumping, coeffs, u, np, ur, ui, vr, vi, r, usq);
    KIND(nlso_rgbmv)(ui, vr, (REAL)-1.0, op, klu, n);   /* nls.f90:646 */
    KIND(nlso_rgbmv)(ur, vi, (REAL)+1.0, op, klu, n);   /* nls.f90:647 */
    for (size_t t = 0; t < np; ++t) { v[t].re = vr[t]; v[t].im = vi[t]; }
    free(w);
    return 0;
}

/* nls.f90:841-870 */
NLSO_API int KIND(nlso_hamiltonian_2d)(const REAL *pumping, const REAL *coeffs, const KIND(cplx) *u, KIND(cplx) *v,
                                       const REAL *blocks, const int *orders, int order, int n)
{
    size_t np = (size_t)n * (size_t)n;
    REAL *w = (REAL *)malloc(sizeof(REAL) * 6 * np);
    if (!w) return -2;
    REAL *ur = w, *ui = w + np, *vr = w + 2 * np, *vi = w + 3 * np, *r = w + 4 * np, *usq = w + 5 * np;
    KIND(rhs_pointwise)(pumping, coeffs, u, np, ur, ui, vr, vi, r, usq);
    int rc = KIND(nlso_rbbmv)(ui, vr, (REAL)-1.0, blocks, orders, order, n);   /* nls.f90:866 */
    if (!rc) rc = KIND(nlso_rbbmv)(ur, vi, (REAL)+1.0, blocks, orders, order, n);   /* nls.f90:867 */
    for (size_t t = 0; t < np; ++t) { v[t].re = vr[t]; v[t].im = vi[t]; }
    free(w);
    return rc;
}

/* Classical RK4 loop shared by runge_kutta (nls.f90:705-734) and runge_kutta_2d (:873-901).
 * Stage arguments are u + (k*dt)/2 (k3: /1); update is u + ((((k1 + 2*k2) + 2*k3) + k4)*dt)/6. */
static int KIND(rk4)(int dim, REAL dt, const KIND(cplx) *u0, const REAL *op, const int *orders, int order,
                     int n, int iters, KIND(cplx) *u, const REAL *pumping, const REAL *coeffs)
{
    size_t np = dim == 1 ? (size_t)n : (size_t)n * (size_t)n;
    KIND(cplx) *k = (KIND(cplx) *)malloc(sizeof(KIND(cplx)) * 5 * np);
    if (!k) return -2;
    KIND(cplx) *k1 = k, *k2 = k + np, *k3 = k + 2 * np, *k4 = k + 3 * np, *arg = k + 4 * np;
    int klu = (order - 1) / 2, rc = 0;
    const REAL zero = 0, half_div = 2, one_div = 1, six = 6;
    for (size_t t = 0; t < np; ++t) u[t] = u0[t];
    for (int it = 0; it < iters && !rc; ++it) {
        KIND(cplx) *ks[4] = {k1, k2, k3, k4};
        for (int s = 0; s < 4 && !rc; ++s) {
            if (s == 0) {
                REAL z = zero * dt / half_div;               /* u + 0.*dt/2, nls.f90:726 */
                for (size_t t = 0; t < np; ++t) { arg[t].re = u[t].re + z; arg[t].im = u[t].im; }
            } else {
                const KIND(cplx) *kp = ks[s - 1];
                REAL d = s == 3 ? one_div : half_div;        /* nls.f90:727-729 */
                for (size_t t = 0; t < np; ++t) {
                    arg[t].re = u[t].re + kp[t].re * dt / d;
                    arg[t].im = u[t].im + kp[t].im * dt / d;
                }
            }
            rc = dim == 1 ? KIND(nlso_hamiltonian)(pumping, coeffs, arg, ks[s], op, klu, n)
                          : KIND(nlso_hamiltonian_2d)(pumping, coeffs, arg, ks[s], op, orders, order, n);
        }
        for (size_t t = 0; t < np; ++t) {                    /* nls.f90:731 / :898 */
            REAL sr = ((k1[t].re + 2 * k2[t].re) + 2 * k3[t].re) + k4[t].re;
            REAL si = ((k1[t].im + 2 * k2[t].im) + 2 * k3[t].im) + k4[t].im;
            u[t].re = u[t].re + sr * dt / six;
            u[t].im = u[t].im + si * dt / six;
        }
    }
    free(k);
    return rc;
}

NLSO_API int KIND(nlso_runge_kutta)(REAL dt, REAL t0, const KIND(cplx) *u0, const REAL *op, int n, int order,
                                    int iters, KIND(cplx) *u, const REAL *pumping, const REAL *coeffs)
{
    (void)t0;
    if (order != 3 && order != 5 && order != 7) return -1;
    return KIND(rk4)(1, dt, u0, op, NULL, order, n, iters, u, pumping, coeffs);
}

NLSO_API int KIND(nlso_runge_kutta_2d)(REAL dt, REAL t0, const KIND(cplx) *u0, int n, const REAL *blocks,
                                       const int *orders, int order, int iters, KIND(cplx) *u,
                                       const REAL *pumping, const REAL *coeffs)
{
    (void)t0;
    if (order != 3 && order != 5 && order != 7) return -1;
    return KIND(rk4)(2, dt, u0, blocks, orders, order, n, iters, u, pumping, coeffs);
}

/* nls.f90:797-813 (solve_nls_1d, :815-827, forwards here) */
NLSO_API int KIND(nlso_solve_nls)(REAL dt, REAL dx, int n, int order, int iters, const REAL *pumping,
                                  const REAL *coeffs, const KIND(cplx) *u0, KIND(cplx) *u)
{
    if (order != 3 && order != 5 && order != 7) return -1;
    REAL *op = (REAL *)malloc(sizeof(REAL) * (size_t)order * (size_t)n);
    if (!op) return -2;
    int rc = KIND(nlso_make_laplacian)(n, order, dx, op);
    if (!rc) rc = KIND(nlso_runge_kutta)(dt, 0, u0, op, n, order, iters, u, pumping, coeffs);
    free(op);
    return rc;
}

/* nls.f90:903-919 */
NLSO_API int KIND(nlso_solve_nls_2d)(REAL dt, REAL dx, int n, int order, int iters, const REAL *pumping,
                                     const REAL *coeffs, const KIND(cplx) *u0, KIND(cplx) *u)
{
    if (order != 3 && order != 5 && order != 7) return -1;
    int orders[7];
    REAL *blocks = (REAL *)malloc(sizeof(REAL) * (size_t)n * (size_t)(2 * order - 1));
    if (!blocks) return -2;
    int rc = KIND(nlso_make_laplacian_2d)(n, order, dx, blocks, orders);
    if (!rc) rc = KIND(nlso_runge_kutta_2d)(dt, 0, u0, n, blocks, orders, order, iters, u, pumping, coeffs);
    free(blocks);
    return rc;
}

/* Fortran-rule complex division (Smith's scaling, what gfortran emits without -ffast-math). */
static KIND(cplx) KIND(cdiv)(KIND(cplx) a, KIND(cplx) b)
{
    KIND(cplx) q;
    REAL abr = b.re < 0 ? -b.re : b.re, abi = b.im < 0 ? -b.im : b.im;
    if (abr >= abi) {
        REAL t = b.im / b.re, d = b.re + b.im * t;
        q.re = (a.re + a.im * t) / d;
        q.im = (a.im - a.re * t) / d;
    } else {
        REAL t = b.re / b.im, d = b.im + b.re * t;
        q.re = (a.re * t + a.im) / d;
        q.im = (a.im * t - a.re) / d;
    }
    return q;
}

/* conjugating dot product sum(conjg(a) * (b * w)), sequential order; w == NULL means weight 1. */
static KIND(cplx) KIND(cdot)(const KIND(cplx) *a, const KIND(cplx) *b, const REAL *w, size_t np)
{
    KIND(cplx) s = {0, 0};
    for (size_t t = 0; t < np; ++t) {
        REAL br = b[t].re, bi = b[t].im;
        if (w) { br = br * w[t]; bi = bi * w[t]; }
        REAL ar = a[t].re, ai = -a[t].im;
        s.re = s.re + (ar * br - ai * bi);
        s.im = s.im + (ar * bi + ai * br);
    }
    return s;
}

/* nls.f90:921-948: always the order-5 operator; r-weighted dot products; complex result. */
NLSO_API int KIND(nlso_chemical_potential_1d)(REAL dx, int n, const REAL *pumping, const REAL *coeffs,
                                              const KIND(cplx) *u0, KIND(cplx) *mu)
{
    enum { order = 5, klu = 2 };
    size_t np = (size_t)n;
    REAL *op = (REAL *)malloc(sizeof(REAL) * order * np), *r = (REAL *)malloc(sizeof(REAL) * np);
    KIND(cplx) *u = (KIND(cplx) *)malloc(sizeof(KIND(cplx)) * np);
    int rc = (!op || !r || !u) ? -2 : 0;
    if (!rc) rc = KIND(nlso_make_laplacian)(n, order, dx, op);
    if (!rc) rc = KIND(nlso_hamiltonian)(pumping, coeffs, u0, u, op, klu, n);
    if (!rc) {
        for (int i = 1; i <= n; ++i) r[i - 1] = ((REAL)i - (REAL)1.0) * dx;   /* nls.f90:941-943 */
        KIND(cplx) M = KIND(cdot)(u0, u0, r, np);
        KIND(cplx) d = KIND(cdot)(u0, u, r, np);
        KIND(cplx) E = { (REAL)0.0 * d.re - (REAL)1.0 * d.im, (REAL)0.0 * d.im + (REAL)1.0 * d.re };   /* (0,1)*d */
        *mu = KIND(cdiv)(E, M);
    }
    free(op); free(r); free(u);
    return rc;
}

/* nls.f90:950-971: order-5 operator, unweighted, real part only. */
NLSO_API int KIND(nlso_chemical_potential_2d)(REAL dx, int n, const REAL *pumping, const REAL *coeffs,
                                              const KIND(cplx) *u0, REAL *mu)
{
    enum { order = 5 };
    size_t np = (size_t)n * (size_t)n;
    int orders[order];
    REAL *blocks = (REAL *)malloc(sizeof(REAL) * (size_t)n * (2 * order - 1));
    KIND(cplx) *u = (KIND(cplx) *)malloc(sizeof(KIND(cplx)) * np);
    int rc = (!blocks || !u) ? -2 : 0;
    if (!rc) rc = KIND(nlso_make_laplacian_2d)(n, order, dx, blocks, orders);
    if (!rc) rc = KIND(nlso_hamiltonian_2d)(pumping, coeffs, u0, u, blocks, orders, order, n);
    if (!rc) {
        KIND(cplx) M = KIND(cdot)(u0, u0, NULL, np);
        KIND(cplx) d = KIND(cdot)(u0, u, NULL, np);
        KIND(cplx) E = { (REAL)0.0 * d.re - (REAL)1.0 * d.im, (REAL)0.0 * d.im + (REAL)1.0 * d.re };
        *mu = KIND(cdiv)(E, M).re;
    }
    free(blocks); free(u);
    return rc;
}

#undef F2
