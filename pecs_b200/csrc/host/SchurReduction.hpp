// SchurReduction.hpp -- cell-local elimination of the LDG current unknowns from a carrier system.
//
// The carrier matrix the reference assembles (reference source/LDG.cpp:85-678) has, in the component-wise block
// numbering [q = (Jx, Jy) | u = rho] (reference source/CarrierPair.cpp:29-33), the form
//
//        [ A_qq   G_qu ] [q]   [r_q]           A_qq = mu^-1 (p, q): a DG mass matrix, block diagonal per CELL
//        [ G_uq   S_uu ] [u] = [r_u]           (8 x 8 per cell; no face term touches the (q, q) block)
//
// UMFPACK factorises all 12 unknowns per cell.  Because A_qq is cell-block-diagonal the currents can be
// eliminated exactly, cell by cell, before any global factorisation (the defining trick of LDG):
//
//        S u = r_u - G_uq A_qq^-1 r_q,     S = S_uu - G_uq A_qq^-1 G_qu       (4 unknowns per cell, SPD)
//        q   = A_qq^-1 r_q - (A_qq^-1 G_qu) u
//
// Only S goes through nested dissection: a third of the unknowns, a 13-cell stencil, and factor tables less than
// half the size -- and the per-step solve streams exactly those tables.  The reduction and the back-substitution
// are two sparse mat-vecs with the fixed matrices T1 = G_uq A_qq^-1 and [A_qq^-1 | -A_qq^-1 G_qu].
// If the (q, q) block of a given matrix is NOT cell-block-diagonal the reduction is refused and the caller
// factorises the full system instead.
#pragma once
#include "Csr.hpp"

namespace pecs {

struct SchurReduction {
  int n_cells = 0;
  CsrMatrix S;        // 4n x 4n   reduced density system
  CsrMatrix T1;       // 4n x 8n   r~ = r_u - T1 r_q
  CsrMatrix Ainv;     // 8n x 8n   cell-block-diagonal A_qq^-1
  CsrMatrix T2;       // 8n x 4n   q = Ainv r_q - T2 u
};

// entries of S not above this fraction of their row's largest entry are rounding residue of exact cancellations
double schur_drop_tolerance();

// A: full carrier matrix (12 n_cells rows).  Returns false (and leaves out untouched) when A_qq couples cells.
// threads: rows are built in that many contiguous ranges (the result does not depend on it).
bool build_schur_reduction(const CsrView& A, int n_cells, SchurReduction& out, int threads = 1);

} // namespace pecs
