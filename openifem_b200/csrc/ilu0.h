// ILU(0) factors for the block preconditioner of the SUPG solvers.
//
// Reference: BlockIncompSchurPreconditioner (source/mpi_supg_solver.cpp:35-192) keeps two Hypre-Euclid ILU(0) factorisations
// (source/preconditioner_pilut.cpp:124-138): P_vv ~ A_vv^-1 (:51, applied at :146, :157, :190) and the preconditioner of the
// inner GMRES on T_pp, an ILU(0) of B2pp = A_pp - A_pv diag(rowsum|A_vv|)^-1 A_vp (:68-133, used at :176-179).
//
// Here: a scalar CSR copy of the matrix, factorised in place (IKJ variant restricted to the pattern, unit lower factor) and
// applied by two level-scheduled triangular sweeps. Rows of one dependency level are independent; the levels are found once per
// sparsity pattern on the host. The factorisation and each sweep are ONE kernel of ONE CTA (1024 threads, __syncthreads between
// levels): the cases that need ILU - viscous, launch-latency-bound problems of a few thousand rows such as the reference's
// SUPG goldens and fsi_leaflet_mpi - have a few hundred levels of a few dozen rows, so a grid-wide barrier per level would cost
// more than the work. The depth of the dependency graph grows with the mesh (n^(1/d) levels), which is why the large,
// mass-dominated cases (config 5: dt = 1e-6) keep the Jacobi factors (scnsim.cu decides by size; Euclid itself degrades to
// rank-local blocks in parallel).
#pragma once
#include <vector>

#include "device.cuh"
#include "linalg.h"

namespace ifem
{
  struct Ilu0
  {
    int n = 0;
    int64_t nnz = 0;
    int n_levels_lower = 0, n_levels_upper = 0;
    DevBuf<int> rowptr, col, diag;                      // scalar CSR pattern, position of the diagonal entry of every row
    DevBuf<double> val;                                 // A, then L (strict lower part, unit diagonal implied) and U in place
    DevBuf<int> order_lower, level_lower, order_upper, level_upper; // rows grouped by level, level offsets
    DevBuf<double> tmp;
    // The sweeps read a second, LEVEL-ORDERED copy of the factors: the r-th row in level order keeps its strictly lower (upper)
    // entries in w_lower (w_upper) consecutive padded slots, so every address of a sweep depends on the level counter only - none
    // on the recurrence - and can be prefetched levels ahead; the recurrence itself runs on a work vector in shared memory.
    int w_lower = 0, w_upper = 0;        // padded entries per row (multiples of 32)
    DevBuf<int> lcol, lsrc, ucol, usrc;  // column of a slot, position of its value in `val` (-1 = padding)
    DevBuf<double> lval, uval, udinv;    // packed by factor(); udinv = 1 / u_ii in level order of the upper sweep

    bool ready() const { return n > 0; }
    // pattern analysis: rowptr [n + 1], sorted columns; every row must hold its diagonal
    void setup(Context &ctx, const std::vector<int64_t> &rowptr_h, const std::vector<int> &col_h);
    // val holds the matrix: factorise in place
    void factor(Context &ctx);
    // x = U^-1 L^-1 b (b and x may alias)
    void solve(Context &ctx, const double *b, double *x);
  };

  // scalar CSR pattern of a square block matrix (R = C = bs): rows bs * node + r, columns bs * col + c
  void scalar_pattern(const Pattern &P, int bs, std::vector<int64_t> &rowptr, std::vector<int> &col);
  // values of the row-plane BCSR matrix -> the scalar CSR copy of scalar_pattern (device to device)
  void bcsr_to_scalar(Context &ctx, const Bcsr &A, const int *rowptr_s, double *val_s);
} // namespace ifem
