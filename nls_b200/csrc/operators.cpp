// operators.cpp -- host-side construction of the finite-difference operators (float64).
//
// O(n*m) setup work that runs once per solve; it stays on the host so that the tables the kernels
// consume are bit-identical to what the reference's builders produce when promoted to double
// (BASELINE.json: "grid construction should be bit-exact").
//
// The reference assembles the radial operator by writing two Toeplitz bands and then patching
// individual band entries (nls.f90:132-259).  Here the same numbers are produced from the
// structure those patches encode: the field is even in r, so a stencil tap that lands on node -q
// folds onto node +q; at r = 0 the operator is 2*D2 (l'Hopital on D1/r); row rho of D1 is divided
// by rho*h.  Each entry is formed with the same roundings as the reference:
//     base = integer / dx ;  folded = base + integer' / dx ;  D1 entry / (rho*h) ;  op = D1 + D2.

#include "internal.h"

#include <cstring>
#include <vector>

namespace nlsb {

namespace {

struct Stencil {
    int m;
    double d1_den_scale;       // dx1 = d1_den_scale * h
    double d2_den_scale;       // dx2 = d2_den_scale * h^2
    int d1[7];                 // numerators; index 0 multiplies x(i + k), as the reference's `row`
    int d2[7];
};

// nls.f90:145-149, :176-180, :210-214
const Stencil kRadial3 = {3, 2.0, 1.0, {1, 0, -1}, {1, -2, 1}};
const Stencil kRadial5 = {5, 12.0, 12.0, {-1, 8, 0, -8, 1}, {-1, 16, -30, 16, -1}};
const Stencil kRadial7 = {7, 60.0, 180.0, {1, -9, 45, 0, -45, 9, -1}, {2, -27, 270, -490, 270, -27, 2}};

const Stencil *radial_stencil(int m)
{
    return m == 3 ? &kRadial3 : m == 5 ? &kRadial5 : m == 7 ? &kRadial7 : nullptr;
}

// nls.f90:310-318, :336-348, :366-382 (centre weight of the 2D cross is one number, not a sum)
bool cross_numerators(int m, double *den_scale, int *w)
{
    static const int w3[3] = {1, -4, 1}, w5[5] = {-1, 16, -60, 16, -1}, w7[7] = {2, -27, 270, -980, 270, -27, 2};
    const int *src = m == 3 ? w3 : m == 5 ? w5 : m == 7 ? w7 : nullptr;
    if (!src) return false;
    *den_scale = m == 3 ? 1.0 : m == 5 ? 12.0 : 180.0;
    for (int i = 0; i < m; ++i) w[i] = src[i];
    return true;
}

}  // namespace

int radial_taps(int n, int m, double h, double *taps)
{
    const Stencil *st = radial_stencil(m);
    if (!st) return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", m);
    if (n < m) return fail(NLSB_ESIZE, "n = %d is smaller than the stencil width %d", n, m);
    const int k = (m - 1) / 2;
    const double dx1 = st->d1_den_scale * h;
    const double dx2 = st->d2_den_scale * (h * h);
    // tap s (multiplying x(i+s)) has numerator row[k - s] in the reference's row ordering
    auto d1 = [&](int s) { return static_cast<double>(st->d1[k - s]) / dx1; };
    auto d2 = [&](int s) { return static_cast<double>(st->d2[k - s]) / dx2; };

    for (int i = 0; i < n; ++i) {
        double *row = taps + static_cast<size_t>(i) * m;
        for (int s = -k; s <= k; ++s) {
            const int j = i + s;
            double a1 = 0.0, a2 = 0.0;
            if (j >= 0 && j < n) {
                a1 = d1(s);
                a2 = d2(s);
                // even fold: the tap that reaches node -j (j >= 1) adds onto node +j
                const int mirror = -j - i;                 // offset whose target is -j
                if (j >= 1 && mirror >= -k) {
                    a1 = a1 + d1(mirror);
                    a2 = a2 + d2(mirror);
                }
                if (i == 0) {
                    a1 = 0.0;                              // D1 row 0 is cleared (nls.f90:104-106)
                    a2 = (j == 0) ? 2 * d2(0) : 4 * d2(s);  // 2*D2 folded (nls.f90:154-155, :187-189, :235-238)
                } else {
                    a1 = a1 / (static_cast<double>(i) * h); // nls.f90:121-129
                }
            }
            row[s + k] = a1 + a2;                          // op = L1 + L2
        }
    }
    return NLSB_OK;
}

int band_to_taps(int n, int m, const double *op, double *taps)
{
    if (m < 1 || (m & 1) == 0) return fail(NLSB_EINVAL, "band height must be odd (got %d)", m);
    const int k = (m - 1) / 2;
    for (int i = 0; i < n; ++i)
        for (int s = -k; s <= k; ++s) {
            const int j = i + s;
            // A(i, j) = op(k + i - j, j) with 0-based band rows
            taps[static_cast<size_t>(i) * m + (s + k)] =
                (j >= 0 && j < n) ? op[static_cast<size_t>(k - s) + static_cast<size_t>(m) * j] : 0.0;
        }
    return NLSB_OK;
}

int taps_to_band(int n, int m, const double *taps, double *op)
{
    const int k = (m - 1) / 2;
    std::memset(op, 0, sizeof(double) * static_cast<size_t>(m) * n);
    for (int i = 0; i < n; ++i)
        for (int s = -k; s <= k; ++s) {
            const int j = i + s;
            if (j >= 0 && j < n)
                op[static_cast<size_t>(k - s) + static_cast<size_t>(m) * j] = taps[static_cast<size_t>(i) * m + (s + k)];
        }
    return NLSB_OK;
}

int banded_from_row(int n, int m, const double *row, double *mat)
{
    if (m < 1 || (m & 1) == 0) return fail(NLSB_EINVAL, "row length must be odd (got %d)", m);
    if (n < m) return fail(NLSB_ESIZE, "n = %d is smaller than the row length %d", n, m);
    const int k = (m - 1) / 2;
    // band row b of column j is matrix entry (i, j) with i = j + b - k: keep it iff 0 <= i < n
    for (int j = 0; j < n; ++j)
        for (int b = 0; b < m; ++b) {
            const int i = j + b - k;
            mat[static_cast<size_t>(b) + static_cast<size_t>(m) * j] = (i >= 0 && i < n) ? row[b] : 0.0;
        }
    return NLSB_OK;
}

int cross_weights(int m, double h, double *wx, double *wy)
{
    int w[7];
    double scale;
    if (!cross_numerators(m, &scale, w)) return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", m);
    const int k = (m - 1) / 2;
    const double dx2 = scale * (h * h);
    for (int t = 0; t < m; ++t) {
        wx[t] = static_cast<double>(w[t]) / dx2;
        wy[t] = (t == k) ? 0.0 : static_cast<double>(w[t]) / dx2;
    }
    return NLSB_OK;
}

int cross_blocks(int n, int m, double h, double *blocks, int *orders)
{
    double wx[7], wy[7];
    int rc = cross_weights(m, h, wx, wy);
    if (rc) return rc;
    if (n < m) return fail(NLSB_ESIZE, "n = %d is smaller than the stencil width %d", n, m);
    const int k = (m - 1) / 2;
    double *p = blocks;
    for (int b = 0; b < m; ++b) {
        if (b == k) {
            // wx is symmetric, so row ordering (index 0 multiplies x(i+k)) does not matter
            rc = banded_from_row(n, m, wx, p);
            if (rc) return rc;
            orders[b] = k;
            p += static_cast<size_t>(m) * n;
        } else {
            for (int i = 0; i < n; ++i) p[i] = wy[b];
            orders[b] = 0;
            p += n;
        }
    }
    return NLSB_OK;
}

int blocks_to_weights(int n, int m, const double *blocks, const int *orders, double *wx, double *wy)
{
    if (m != 3 && m != 5 && m != 7) return fail(NLSB_EORDER, "order must be 3, 5 or 7 (got %d)", m);
    if (n < m) return fail(NLSB_ESIZE, "n = %d is smaller than the stencil width %d", n, m);
    const int k = (m - 1) / 2;
    const double *p = blocks;
    for (int b = 0; b < m; ++b) {
        if (orders[b] != (b == k ? k : 0))
            return fail(NLSB_EOPERATOR, "orders[%d] = %d does not describe a cross stencil", b, orders[b]);
        if (b == k) {
            // A(i, j) = band(k + i - j, j): in the full interior column j = k, band row r holds the
            // weight of offset s = j - i = k - r
            for (int s = -k; s <= k; ++s)
                wx[s + k] = p[static_cast<size_t>(k - s) + static_cast<size_t>(m) * k];
            // verify that the band is the truncated Toeplitz band of wx
            for (int j = 0; j < n; ++j)
                for (int r = 0; r < m; ++r) {
                    const int i = j + r - k;
                    const double expect = (i >= 0 && i < n) ? wx[(j - i) + k] : 0.0;
                    if (p[static_cast<size_t>(r) + static_cast<size_t>(m) * j] != expect)
                        return fail(NLSB_EOPERATOR, "middle block is not a truncated Toeplitz band (column %d)", j);
                }
            p += static_cast<size_t>(m) * n;
        } else {
            // source line = line + (b - k): weight of offset s = b - k across lines
            wy[b] = p[0];
            for (int i = 1; i < n; ++i)
                if (p[i] != wy[b])
                    return fail(NLSB_EOPERATOR, "off-diagonal block %d varies along the line", b);
            p += n;
        }
    }
    wy[k] = 0.0;
    return NLSB_OK;
}

}  // namespace nlsb
