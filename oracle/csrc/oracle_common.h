// ORACLE (test infrastructure, NOT product code): shared helpers for the CPU
// restatement of OpenIFEM's hot path. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference arm may load the library built from
// these files.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace oracle
{
  // Small dense tensors in the spirit of dealii::Tensor<rank,dim>; everything is
  // fully dense on purpose so the arithmetic (and its cost) follows the
  // reference's Tensor operations term by term.
  template <int dim>
  struct T1
  {
    double v[dim];
    T1() { for (int i = 0; i < dim; ++i) v[i] = 0; }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
  };

  template <int dim>
  struct T2
  {
    double v[dim][dim];
    T2() { for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) v[i][j] = 0; }
    double *operator[](int i) { return v[i]; }
    const double *operator[](int i) const { return v[i]; }
  };

  template <int dim>
  inline double dot(const T1<dim> &a, const T1<dim> &b)
  {
    double s = 0;
    for (int i = 0; i < dim; ++i) s += a[i] * b[i];
    return s;
  }
  // scalar_product(Tensor<2>, Tensor<2>) = double contraction
  template <int dim>
  inline double scalar_product(const T2<dim> &a, const T2<dim> &b)
  {
    double s = 0;
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) s += a[i][j] * b[i][j];
    return s;
  }
  // Tensor<2> * Tensor<1>
  template <int dim>
  inline T1<dim> mul(const T2<dim> &a, const T1<dim> &b)
  {
    T1<dim> r;
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) r[i] += a[i][j] * b[j];
    return r;
  }
  template <int dim>
  inline T2<dim> mul(const T2<dim> &a, const T2<dim> &b)
  {
    T2<dim> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k) r[i][j] += a[i][k] * b[k][j];
    return r;
  }
  template <int dim>
  inline double trace(const T2<dim> &a)
  {
    double s = 0;
    for (int i = 0; i < dim; ++i) s += a[i][i];
    return s;
  }
  inline double det(const T2<2> &a) { return a[0][0] * a[1][1] - a[0][1] * a[1][0]; }
  inline double det(const T2<3> &a)
  {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
           a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
  }
  inline T2<2> invert(const T2<2> &a)
  {
    T2<2> r;
    const double d = 1.0 / det(a);
    r[0][0] = a[1][1] * d;  r[0][1] = -a[0][1] * d;
    r[1][0] = -a[1][0] * d; r[1][1] = a[0][0] * d;
    return r;
  }
  inline T2<3> invert(const T2<3> &a)
  {
    T2<3> r;
    const double d = 1.0 / det(a);
    r[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * d;
    r[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * d;
    r[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * d;
    r[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * d;
    r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * d;
    r[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * d;
    r[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * d;
    r[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * d;
    r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * d;
    return r;
  }

  // Q1 geometry of one cell at one reference point: Jacobian J_ij = dx_i/dxi_j.
  template <int dim>
  inline T2<dim> jacobian(const double *vertices, const int *cell_v, const double *dNgeo /*[2^d][dim]*/)
  {
    T2<dim> J;
    for (int v = 0; v < (1 << dim); ++v)
      {
        const double *X = vertices + (size_t)cell_v[v] * dim;
        for (int i = 0; i < dim; ++i)
          for (int j = 0; j < dim; ++j) J[i][j] += X[i] * dNgeo[v * dim + j];
      }
    return J;
  }

  // Position of entry (row, c) in a CSR row with sorted column indices
  // (what PETSc MatSetValues does with a binary search).
  inline int64_t csr_find(const int64_t *rowptr, const int *col, int row, int c)
  {
    const int *b = col + rowptr[row], *e = col + rowptr[row + 1];
    const int *it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? (it - col) : -1;
  }

  // Constraint lines with masters (hanging nodes of a locally refined mesh, DoFTools::make_hanging_node_constraints at
  // reference source/mpi_fluid_solver.cpp:182-184), already resolved as AffineConstraints::close() leaves them: a dof g with
  // con[g] == 2 equals sum_k weight[k] * x[master[k]] + inhom[g] over ptr[g] <= k < ptr[g + 1], every master unconstrained.
  // Set through oracle_set_constraint_lines (oracle_ins.cpp) before an assembly; empty = pure Dirichlet lines only.
  struct ConstraintLines
  {
    std::vector<int64_t> ptr;
    std::vector<int> master;
    std::vector<double> weight;
  };
  inline ConstraintLines &constraint_lines()
  {
    static ConstraintLines lines;
    return lines;
  }

  // AffineConstraints::distribute_local_to_global: an unconstrained local row i goes to its global row, a row with a
  // hanging-node line (con == 2) to the rows of its masters with their weights, a Dirichlet row (con == 1) nowhere;
  // columns likewise, and every constrained column j moves K_ij * inhomogeneity_j to the right-hand side of the target
  // rows. A constrained row itself only receives |K_ii| (or the average |diag| of the local matrix if that is zero) on
  // the diagonal and, when use_inhomogeneities_for_rhs, rhs_i += diag * inhomogeneity_i.
  // Restated from deal.II's documented algorithm
  // (affine_constraints.templates.h, make_sorted_row_list / set_matrix_diagonals / resolve_vector_entry);
  // call sites: reference mpi_insim.cpp:348-355, mpi_scnsim.cpp:548-560, mpi_hyper_elasticity.cpp:507-522.
  inline void distribute_local_to_global(int n, const double *K, const double *f, const int *dofs,
                                         const unsigned char *con, const double *inhom,
                                         const int64_t *rowptr, const int *col, double *A, double *rhs,
                                         bool use_inhomogeneities_for_rhs)
  {
    const ConstraintLines &L = constraint_lines();
    auto add = [&](int r, int c, double v) {
      const int64_t p = csr_find(rowptr, col, r, c);
      if (p < 0)
        {
          std::fprintf(stderr, "oracle: entry (%d, %d) is missing from the sparsity pattern\n", r, c);
          std::abort();
        }
#pragma omp atomic
      A[p] += v;
    };
    double average_diagonal = 0;
    for (int i = 0; i < n; ++i) average_diagonal += std::fabs(K[i * n + i]);
    average_diagonal /= n;
    for (int i = 0; i < n; ++i)
      {
        const int gi = dofs[i];
        int n_rows = 1, one_row = gi;
        const int *rows = &one_row;
        double one_w = 1.0;
        const double *row_w = &one_w;
        if (con[gi])
          {
            const double d = std::fabs(K[i * n + i]) != 0 ? std::fabs(K[i * n + i]) : average_diagonal;
            add(gi, gi, d);
            if (rhs && use_inhomogeneities_for_rhs && inhom)
              {
#pragma omp atomic
                rhs[gi] += d * inhom[gi];
              }
            if (con[gi] != 2) continue;
            n_rows = (int)(L.ptr[gi + 1] - L.ptr[gi]);
            rows = L.master.data() + L.ptr[gi];
            row_w = L.weight.data() + L.ptr[gi];
          }
        for (int t = 0; t < n_rows; ++t)
          {
            const int r = rows[t];
            const double wr = row_w[t];
            double fi = f ? wr * f[i] : 0.0;
            for (int j = 0; j < n; ++j)
              {
                const int gj = dofs[j];
                const double kij = K[i * n + j];
                if (con[gj])
                  {
                    if (inhom) fi -= wr * kij * inhom[gj];
                    if (con[gj] == 2 && kij != 0.0)
                      for (int64_t k = L.ptr[gj]; k < L.ptr[gj + 1]; ++k) add(r, L.master[k], wr * L.weight[k] * kij);
                    continue;
                  }
                if (kij == 0.0) continue; // deal.II elides exact zeros
                add(r, gj, wr * kij);
              }
            if (rhs)
              {
#pragma omp atomic
                rhs[r] += fi;
              }
          }
      }
  }
} // namespace oracle
