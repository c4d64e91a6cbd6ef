/*
 * nls_oracle.c -- CPU oracle for the daskol/nls hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * the library built from this file.  The engine (nls_b200/) never links, imports or calls it.
 *
 * What it is: a C restatement of /root/reference/nls/nls.f90 (the whole numerical core of the
 * reference) in two real kinds:
 *     *_sp  float   -- the precision the reference ships (nls.f90:10, sgbmv at :540)
 *     *_dp  double  -- the same algorithm kind-promoted; the engine's 1e-10 bar is measured here
 *
 * Pinning status: the reference itself cannot be compiled in this image (no Fortran
 * compiler), so the oracle is pinned only by the reference's own valid golden tables --
 * make_banded_matrix (test/test_nls.f95:50-59) and rgbmv (test/test_nls.f95:166-184) -- plus the
 * stale make_laplacian_o5 table (:102-108) reproduced under the 24h^2 switch, operator identities,
 * SciPy's s/dgbmv, and the reference's importable nls/pumping.py for the inputs.  No reference test
 * pins hamiltonian*, runge_kutta*, solve_nls* or chemical_potential*: for those the status is
 * "PARITY UNPINNED" (restatement reviewed line by line against nls.f90, cited in the impl header).
 *
 * Build: gcc -O3 -ffp-contract=off -fPIC -shared (see oracle/Makefile) -- mirrors the reference's
 * `gfortran -O3` without fast-math (nls/makefile:16).
 */

#include <stdlib.h>
#include <stddef.h>

#define NLSO_API __attribute__((visibility("default")))

#define REAL float
#define KIND(name) name##_sp
#include "nls_oracle_impl.h"
#undef REAL
#undef KIND

#define REAL double
#define KIND(name) name##_dp
#include "nls_oracle_impl.h"
#undef REAL
#undef KIND

NLSO_API void nlso_version(int *major, int *minor, int *patch)
{
    /* nls.f90:29-37 */
    *major = 0; *minor = 2; *patch = 0;
}
