// Krylov solvers of the hot path, host-driven over device vectors:
//   * FGMRES  = deal.II SolverFGMRES<VectorType> (restart 30, right preconditioned,
//               modified Gram-Schmidt through add_and_dot, convergence tested on
//               the least-squares residual of the (j+1) x j Hessenberg block, so it
//               lags the newest Arnoldi column by one) - call sites
//               reference mpi_insim.cpp:379-388, mpi_supg_solver.cpp:311-321;
//   * CG      = PETSc KSPCG driven by deal.II SolverControl (absolute residual) -
//               mpi_insim.cpp:73-83, 88-109, mpi_solid_solver.cpp:151-157;
//   * BiCGStab (right preconditioned, x0 = 0) = the inexact stand-in for the MUMPS
//               LU of the velocity block (mpi_insim.cpp:124-127), generalising the
//               in-tree Krylov-for-A~ precedent mpi_insimex.cpp:114-124.
// Operators and preconditioners are callables (const double *src, double *dst)
// that enqueue kernels on ctx.stream.
#pragma once
#include <cmath>
#include <functional>
#include <stdexcept>
#include <vector>

#include "linalg.h"

namespace ifem
{
  using LinOp = std::function<void(const double *, double *)>;

  struct SolveResult
  {
    int iterations = 0;
    double residual = 0.0;
    bool converged = false;
  };

  // workspace that grows on demand and is reused across solves
  struct VecPool
  {
    std::vector<DevBuf<double>> v;
    double *get(size_t i, int64_t n)
    {
      if (v.size() <= i) v.resize(i + 1);
      if ((int64_t)v[i].n < n) v[i].alloc(n);
      return v[i].p;
    }
  };

  SolveResult cg(Context &ctx, const VecSpace &n, const LinOp &A, const double *b, double *x, bool x_is_zero, double tol_abs, int max_it,
                 VecPool &pool);

  // preconditioned CG, x0 = 0 (PETSc KSPCG + PC; solid solver, mpi_solid_solver.cpp:143-161)
  SolveResult pcg(Context &ctx, const VecSpace &n, const LinOp &A, const LinOp &prec, const double *b, double *x, double tol_abs,
                  int max_it, VecPool &pool);

  SolveResult bicgstab(Context &ctx, const VecSpace &n, const LinOp &A, const LinOp &prec, const double *b, double *x, double tol_abs,
                       int max_it, VecPool &pool);

  // x0 = 0 is assumed (x is overwritten). Throws on failure like deal.II's NoConvergence.
  // fused_orthogonalisation: the Arnoldi vector is orthogonalised by orthogonalise_cgs2 (three reductions and one host
  // synchronisation per iteration) instead of deal.II's modified Gram-Schmidt loop (one of each per basis vector) - for solves
  // INSIDE a preconditioner, where only the accuracy of the result matters (the T_pp solve of the SUPG block preconditioner)
  SolveResult fgmres(Context &ctx, const VecSpace &n, const LinOp &A, const LinOp &prec, const double *b, double *x, double tol_abs,
                     int64_t max_it, int basis_size, VecPool &pool, bool fused_orthogonalisation = false);
} // namespace ifem

namespace ifem
{
  // Plain CG in fp64 with x0 = 0 and an absolute tolerance, driven from device-resident scalars like the fp32 inner solvers
  // (inner32.h): the two dot products of an iteration are finished inside the kernels that produce them (summed over the ranks
  // through the peer link when there is one, peer.h), alpha and beta never visit the host, and the host looks at the state
  // only every few iterations. Replaces the host-driven cg() where the same small system is solved many times per time step:
  // "CG for Mp" (PETSc KSPCG at reference source/mpi_insim.cpp:73-83). `A` only enqueues work on ctx.stream.
  class DeviceCG64
  {
  public:
    ~DeviceCG64();
    SolveResult solve(Context &ctx, const VecSpace &n, const LinOp &A, const double *b, double *x, double tol_abs, int max_it);

  private:
    DevBuf<double> r, p, ap, partials, red;
    DevBuf<unsigned int> counter;
    DevBuf<int> state;
    void *h_state = nullptr;
    int last_its = 0; // iterations of the previous solve: that many are enqueued before the first look at the state
  };
} // namespace ifem
