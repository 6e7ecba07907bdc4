/* pecs_b200_selftest.h -- C ABI of libpecs_b200_selftest.so: CPU checkers of the product's device formulas and setup
 * tables.  TEST INFRASTRUCTURE: a separate library (pecs_b200/csrc/selftest), not part of the product's libpecs_b200.so,
 * never on the per-step path (which has no CPU implementation).  Used by tests/ only. */
#ifndef PECS_B200_SELFTEST_H
#define PECS_B200_SELFTEST_H
#include "pecs_b200_host.h"
#ifdef __cplusplus
extern "C" {
#endif

/* CPU check of the arithmetic the production RHS kernels run (pecs_b200/csrc/rhs_math.hpp is compiled into both the
 * kernels and this function): the carrier right-hand sides of subdomain `which` (0 / 1) from host states -- u1, u2 the
 * two carrier vectors of the subdomain, o1, o2 those of the other subdomain (interface traces; NULL: cell terms only, no
 * face terms), X the Poisson vector; rhs1 / rhs2 in the [Jx|Jy|rho] layout.  Test infrastructure: the product's per-step
 * path never calls it. */
pecs_status pecs_solarcell_selftest_carrier_rhs(pecs_solarcell* p, int32_t which, const double* u1, const double* u2,
                                                const double* o1, const double* o2, const double* X, double* rhs1,
                                                double* rhs2);
/* the same for the potential rows of the Poisson right-hand side (static int N_a table + charge row; rows of the Poisson
 * cells, phi_rows[n_poisson_cells]) and for the RT0 field at the patch vertices of the output path (field[4n][2]) */
pecs_status pecs_solarcell_selftest_poisson_rows(pecs_solarcell* p, const double* const densities[4], double* phi_rows);
pecs_status pecs_solarcell_selftest_field_patches(pecs_solarcell* p, const double* X, double scale, double* field);
/* Additionally runs the HOST numeric factorisation and the host reference of the two solve sweeps on rhs b, so
 * that the CPU test-suite can check plan + factor tables against the matrix (residual) without a GPU. */
pecs_status pecs_solarcell_selftest_direct_solve(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, const double* b,
                                                 double* x);
/* The host half of building system `which` (0..3 carriers, 4 Poisson) as pecs_ctx_create runs it -- Schur reduction,
 * nested dissection, symbolic plan, matrix in elimination order -- reduced to hashes: the factorised matrix, T1, T2, A^-1 (each with its ELL table),
 * the permutation, the boundary lists + child maps, P A P^T and its transpose.  The preparation is threaded
 * (PECS_B200_SETUP_THREADS); the CPU tests call this with different thread counts and require identical hashes. */
pecs_status pecs_solarcell_selftest_prepared_hashes(pecs_solarcell* p, int32_t which, uint64_t hashes[8]);
/* The ELL layout the device streams (HostEll: slot-major, groups of four columns) against the CSR matrix it was built
 * from: table 0..3 = S, T1, A^-1, T2 of carrier `which`; y_ell = the slot-by-slot arithmetic of the device kernel on the
 * host table (rows of S and T1 in a scrambled order, as on the device), y_csr = the CSR product in the same row order;
 * shape = {rows, slots per row, columns per slot}. */
pecs_status pecs_solarcell_selftest_ell_matvec(pecs_solarcell* p, int32_t which, int32_t table, const double* x, double* y_ell,
                                               double* y_csr, int32_t* shape);

#ifdef __cplusplus
}
#endif
#endif /* PECS_B200_SELFTEST_H */
