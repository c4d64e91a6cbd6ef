// internal.h -- declarations shared by the translation units of libnls_b200.so (not installed).
#pragma once

#include "../../include/nls_b200.h"

#include <cstddef>

namespace nlsb {

// Records a printf-style message for nlsb_last_error() and returns `code`.
int fail(int code, const char *fmt, ...);
// Adds to the library-wide count of kernel launches (reported by nlsb_kernel_launches()).
void count_launches(unsigned long long n);

// ---- operators.cpp (host) -----------------------------------------------------------------------
int radial_taps(int n, int m, double h, double *taps);
int band_to_taps(int n, int m, const double *op, double *taps);
int taps_to_band(int n, int m, const double *taps, double *op);
int banded_from_row(int n, int m, const double *row, double *mat);
int cross_weights(int m, double h, double *wx, double *wy);
int cross_blocks(int n, int m, double h, double *blocks, int *orders);
int blocks_to_weights(int n, int m, const double *blocks, const int *orders, double *wx, double *wy);

// The seven coefficients of coeffs[23] that the right-hand side reads (nls.f90:580, :643-644),
// named by their Fortran index.
struct RhsCoeffs {
    double c3, c4, c5, c6, c12, c13, c14;
};

inline RhsCoeffs rhs_coeffs_from(const double *coeffs)
{
    return RhsCoeffs{coeffs[2], coeffs[3], coeffs[4], coeffs[5], coeffs[11], coeffs[12], coeffs[13]};
}

struct CrossWeights {
    double wx[7];
    double wy[7];
};

}  // namespace nlsb
