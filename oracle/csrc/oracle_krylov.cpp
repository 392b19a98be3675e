// ORACLE (test infrastructure, NOT product code).
//
// The inner Krylov loops of the CPU restatement on all host cores: the reference runs them inside PETSc over MPI ranks
// (KSPCG at source/mpi_insim.cpp:73-83, 88-109; MatMult at :117, :388), one OpenMP thread team stands in for the ranks here.
// oracle/ins.py keeps the same algorithms as plain numpy (cg_py / bicgstab_py); tests/test_oracle_krylov_cpu.py checks the
// two against each other. bench.py's CPU baseline runs through these so that the vector updates between two products use
// every core like the products themselves.
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{
  constexpr int64_t kParallelMin = 50000; // below this a fork/join costs more than the loop

  struct Csr
  {
    int64_t n;
    const int64_t *rp;
    const int *ci;
    const double *v;
  };

  void spmv(const Csr &A, const double *x, double *y)
  {
    const int64_t n = A.n;
#pragma omp parallel for schedule(static) if (n > kParallelMin)
    for (int64_t r = 0; r < n; ++r)
      {
        double s = 0;
        for (int64_t k = A.rp[r]; k < A.rp[r + 1]; ++k) s += A.v[k] * x[A.ci[k]];
        y[r] = s;
      }
  }

  double dot(int64_t n, const double *a, const double *b)
  {
    double s = 0;
#pragma omp parallel for schedule(static) if (n > kParallelMin) reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  }

  // node-block Jacobi: y_node = Binv_node x_node (bs x bs row-major blocks; bs = 0: identity)
  void block_jacobi(int64_t n, int bs, const double *binv, const double *x, double *y)
  {
    if (bs <= 0)
      {
        std::memcpy(y, x, (size_t)n * sizeof(double));
        return;
      }
    const int64_t nn = n / bs;
#pragma omp parallel for schedule(static) if (n > kParallelMin)
    for (int64_t i = 0; i < nn; ++i)
      for (int r = 0; r < bs; ++r)
        {
          double s = 0;
          for (int c = 0; c < bs; ++c) s += binv[(i * bs + r) * bs + c] * x[i * bs + c];
          y[i * bs + r] = s;
        }
  }
} // namespace

extern "C"
{
  int oracle_set_threads(int n)
  {
    if (n > 0) omp_set_num_threads(n);
    int got = 0;
#pragma omp parallel
    {
#pragma omp single
      got = omp_get_num_threads();
    }
    return got;
  }

  // Plain CG, absolute residual tolerance, x holds the initial guess (PETSc KSPCG + PCNONE driven by deal.II SolverControl;
  // reference call sites mpi_insim.cpp:73-83, 88-109). Returns the iteration count; *res_out = final |r|.
  int oracle_cg(int64_t n, const int64_t *rp, const int *ci, const double *v, const double *b, double *x, int x_is_zero, double tol_abs,
                int64_t max_it, double *res_out)
  {
    const Csr A{n, rp, ci, v};
    std::vector<double> r(n), p(n), Ap(n);
    if (x_is_zero)
      std::memcpy(r.data(), b, (size_t)n * sizeof(double));
    else
      {
        spmv(A, x, Ap.data());
#pragma omp parallel for schedule(static) if (n > kParallelMin)
        for (int64_t i = 0; i < n; ++i) r[i] = b[i] - Ap[i];
      }
    double rr = dot(n, r.data(), r.data());
    double res = std::sqrt(rr);
    int64_t it = 0;
    if (res > tol_abs)
      {
        p = r;
        while (it < max_it)
          {
            spmv(A, p.data(), Ap.data());
            const double alpha = rr / dot(n, p.data(), Ap.data());
            double rr_new = 0;
#pragma omp parallel for schedule(static) if (n > kParallelMin) reduction(+ : rr_new)
            for (int64_t i = 0; i < n; ++i)
              {
                x[i] += alpha * p[i];
                r[i] -= alpha * Ap[i];
                rr_new += r[i] * r[i];
              }
            ++it;
            res = std::sqrt(rr_new);
            if (res <= tol_abs) break;
            const double beta = rr_new / rr;
#pragma omp parallel for schedule(static) if (n > kParallelMin)
            for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
            rr = rr_new;
          }
      }
    *res_out = res;
    return (int)it;
  }

  // Right-preconditioned BiCGStab with the node-block Jacobi preconditioner, x0 = 0: the inexact stand-in for the MUMPS LU
  // of the velocity block (mpi_insim.cpp:124-127; in-tree Krylov-for-A precedent mpi_insimex.cpp:114-124). Statement by
  // statement the same as oracle/ins.py bicgstab_py.
  int oracle_bicgstab(int64_t n, const int64_t *rp, const int *ci, const double *v, int bs, const double *binv, const double *b, double *x,
                      double tol_abs, int64_t max_it, double *res_out)
  {
    const Csr A{n, rp, ci, v};
    std::vector<double> r(b, b + n), r0(b, b + n), vv(n, 0.0), p(n, 0.0), ph(n), s(n), sh(n), t(n);
    std::memset(x, 0, (size_t)n * sizeof(double));
    double res = std::sqrt(dot(n, r.data(), r.data()));
    int64_t it = 0;
    if (res > tol_abs)
      {
        double rho = 1, alpha = 1, omega = 1;
        while (it < max_it)
          {
            const double rho_new = dot(n, r0.data(), r.data());
            const double beta = (rho_new / rho) * (alpha / omega);
#pragma omp parallel for schedule(static) if (n > kParallelMin)
            for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * vv[i]);
            block_jacobi(n, bs, binv, p.data(), ph.data());
            spmv(A, ph.data(), vv.data());
            alpha = rho_new / dot(n, r0.data(), vv.data());
            double ss = 0;
#pragma omp parallel for schedule(static) if (n > kParallelMin) reduction(+ : ss)
            for (int64_t i = 0; i < n; ++i)
              {
                s[i] = r[i] - alpha * vv[i];
                ss += s[i] * s[i];
              }
            ++it;
            res = std::sqrt(ss);
            if (res <= tol_abs)
              {
#pragma omp parallel for schedule(static) if (n > kParallelMin)
                for (int64_t i = 0; i < n; ++i) x[i] += alpha * ph[i];
                break;
              }
            block_jacobi(n, bs, binv, s.data(), sh.data());
            spmv(A, sh.data(), t.data());
            double ts = 0, tt = 0;
#pragma omp parallel for schedule(static) if (n > kParallelMin) reduction(+ : ts, tt)
            for (int64_t i = 0; i < n; ++i)
              {
                ts += t[i] * s[i];
                tt += t[i] * t[i];
              }
            omega = ts / tt;
            double rr = 0;
#pragma omp parallel for schedule(static) if (n > kParallelMin) reduction(+ : rr)
            for (int64_t i = 0; i < n; ++i)
              {
                x[i] += alpha * ph[i] + omega * sh[i];
                r[i] = s[i] - omega * t[i];
                rr += r[i] * r[i];
              }
            res = std::sqrt(rr);
            rho = rho_new;
            if (res <= tol_abs) break;
          }
      }
    *res_out = res;
    return (int)it;
  }

  // Split the [u | p] system (CSR, n rows, first nu rows / columns = velocity block) into its four blocks, as the reference
  // holds them from the start (PETScWrappers::MPI::BlockSparseMatrix, mpi_fluid_solver.cpp:320-323). Two passes: count == 1
  // fills the block row pointers rp_b[4][...] (sizes nu+1, nu+1, np+1, np+1), count == 0 fills columns and values.
  void oracle_csr_split(int64_t n, int64_t nu, const int64_t *rp, const int *ci, const double *v, int count, int64_t *rp_uu,
                        int64_t *rp_up, int64_t *rp_pu, int64_t *rp_pp, int *ci_uu, double *v_uu, int *ci_up, double *v_up, int *ci_pu,
                        double *v_pu, int *ci_pp, double *v_pp)
  {
    if (count)
      {
        rp_uu[0] = rp_up[0] = rp_pu[0] = rp_pp[0] = 0;
#pragma omp parallel for schedule(static) if (n > kParallelMin)
        for (int64_t r = 0; r < n; ++r)
          {
            int64_t a = 0;
            for (int64_t k = rp[r]; k < rp[r + 1]; ++k) a += ci[k] < nu;
            const int64_t b = rp[r + 1] - rp[r] - a;
            if (r < nu)
              {
                rp_uu[r + 1] = a;
                rp_up[r + 1] = b;
              }
            else
              {
                rp_pu[r - nu + 1] = a;
                rp_pp[r - nu + 1] = b;
              }
          }
        for (int64_t r = 0; r < nu; ++r)
          {
            rp_uu[r + 1] += rp_uu[r];
            rp_up[r + 1] += rp_up[r];
          }
        for (int64_t r = 0; r < n - nu; ++r)
          {
            rp_pu[r + 1] += rp_pu[r];
            rp_pp[r + 1] += rp_pp[r];
          }
        return;
      }
#pragma omp parallel for schedule(static) if (n > kParallelMin)
    for (int64_t r = 0; r < n; ++r)
      {
        const bool top = r < nu;
        const int64_t rr = top ? r : r - nu;
        int64_t a = top ? rp_uu[rr] : rp_pu[rr], b = top ? rp_up[rr] : rp_pp[rr];
        int *ca = top ? ci_uu : ci_pu, *cb = top ? ci_up : ci_pp;
        double *va = top ? v_uu : v_pu, *vb = top ? v_up : v_pp;
        for (int64_t k = rp[r]; k < rp[r + 1]; ++k)
          if (ci[k] < nu)
            {
              ca[a] = ci[k];
              va[a++] = v[k];
            }
          else
            {
              cb[b] = (int)(ci[k] - nu);
              vb[b++] = v[k];
            }
      }
  }
}
