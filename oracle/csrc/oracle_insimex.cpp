// ORACLE (test infrastructure, NOT product code).
// CPU restatement of Fluid::MPI::InsIMEX<dim>::assemble (reference source/mpi_insimex.cpp:150-355): the
// implicit-explicit twin of InsIM. Viscous, grad-div, pressure-coupling and mass/dt terms are implicit (the matrix does
// not depend on the solution and is only assembled while `assemble_system` is set: time steps 1 and 2); convection is
// explicit and lives in the right-hand side, which is evaluated at present_solution. The unknown is the increment of the
// solution over the time step (solution_time_increment).
#include "oracle_common.h"

using namespace oracle;

namespace
{
  template <int dim>
  void assemble_imex(int nu, int np, int n_cells, const double *vertices, const int *cells, const int *cell_dofs, int nq,
                     const double *qw, const double *Nu, const double *dNu, const double *Np, const double *dNgeo, int nqf,
                     const double *qwf, const double *Nu_face, const double *dNgeo_face, const double *present,
                     const double *fsi_acc, const int *indicator, double viscosity, double gamma, double rho, double dt,
                     const double *grav, int n_bfaces, const int *bfaces, int n_neumann, const int *neumann_ids,
                     const double *neumann_vals, const unsigned char *con, const double *inhom, int assemble_system,
                     const int64_t *rowptr, const int *col, double *A, double *M, double *rhs)
  {
    const int dpc = nu * dim + np, nv = 1 << dim;
    T1<dim> gravity;
    for (int d = 0; d < dim; ++d) gravity[d] = grav[d];
    std::vector<std::vector<std::pair<int, double>>> cell_nfaces(n_neumann ? n_cells : 0);
    for (int f = 0; f < (n_neumann ? n_bfaces : 0); ++f)
      for (int k = 0; k < n_neumann; ++k)
        if (bfaces[3 * f + 2] == neumann_ids[k]) cell_nfaces[bfaces[3 * f]].push_back({bfaces[3 * f + 1], neumann_vals[k]});

#pragma omp parallel
    {
      std::vector<double> local_matrix(dpc * dpc), local_mass(dpc * dpc), local_rhs(dpc), div_phi_u(dpc), phi_p(dpc);
      std::vector<T1<dim>> phi_u(dpc);
      std::vector<T2<dim>> grad_phi_u(dpc);
#pragma omp for schedule(dynamic, 4)
      for (int cell = 0; cell < n_cells; ++cell)
        {
          const int *dofs = cell_dofs + (size_t)cell * dpc, *cv = cells + (size_t)cell * nv;
          std::fill(local_matrix.begin(), local_matrix.end(), 0.0);
          std::fill(local_mass.begin(), local_mass.end(), 0.0);
          std::fill(local_rhs.begin(), local_rhs.end(), 0.0);
          const int ind = indicator ? indicator[cell] : 0;
          for (int q = 0; q < nq; ++q)
            {
              const T2<dim> J = jacobian<dim>(vertices, cv, dNgeo + (size_t)q * nv * dim);
              const T2<dim> Jinv = invert(J);
              const double JxW = det(J) * qw[q];
              for (int k = 0; k < dpc; ++k)
                {
                  phi_u[k] = T1<dim>();
                  grad_phi_u[k] = T2<dim>();
                  div_phi_u[k] = 0;
                  phi_p[k] = 0;
                  if (k < nu * dim)
                    {
                      const int node = k / dim, c = k % dim;
                      phi_u[k][c] = Nu[q * nu + node];
                      const double *dr = dNu + ((size_t)q * nu + node) * dim;
                      for (int i = 0; i < dim; ++i)
                        for (int j = 0; j < dim; ++j) grad_phi_u[k][c][i] += dr[j] * Jinv[j][i];
                      div_phi_u[k] = grad_phi_u[k][c][c];
                    }
                  else
                    phi_p[k] = Np[q * np + (k - nu * dim)];
                }
              // get_function_values / gradients / divergences of present_solution (:224-240)
              T1<dim> v, acc;
              T2<dim> g;
              double p = 0;
              for (int k = 0; k < dpc; ++k)
                {
                  const double up = present[dofs[k]], fa = fsi_acc ? fsi_acc[dofs[k]] : 0.0;
                  for (int i = 0; i < dim; ++i)
                    {
                      v[i] += up * phi_u[k][i];
                      acc[i] += fa * phi_u[k][i];
                      for (int j = 0; j < dim; ++j) g[i][j] += up * grad_phi_u[k][i][j];
                    }
                  p += up * phi_p[k];
                }
              const double div = trace(g);
              for (int i = 0; i < dpc; ++i)
                {
                  if (assemble_system)
                    for (int j = 0; j < dpc; ++j)
                      {
                        // :258-270 (no convection in the matrix)
                        local_matrix[i * dpc + j] += (viscosity * scalar_product(grad_phi_u[j], grad_phi_u[i]) - div_phi_u[i] * phi_p[j] -
                                                      phi_p[i] * div_phi_u[j] + gamma * div_phi_u[j] * div_phi_u[i] * rho +
                                                      dot(phi_u[i], phi_u[j]) / dt * rho) *
                                                     JxW;
                        local_mass[i * dpc + j] += (dot(phi_u[i], phi_u[j]) + phi_p[i] * phi_p[j]) * JxW;
                      }
                  // :272-284
                  local_rhs[i] -= (viscosity * scalar_product(g, grad_phi_u[i]) - div * phi_p[i] - p * div_phi_u[i] +
                                   gamma * div * div_phi_u[i] * rho + dot(mul(g, v), phi_u[i]) * rho - dot(gravity, phi_u[i]) * rho) *
                                  JxW;
                  if (ind == 1) local_rhs[i] += dot(acc, phi_u[i]) * rho * JxW; // :285-291 (CellProperty::fsi_stress is zero in the MPI path)
                }
            }
          if (n_neumann) // :300-330
            for (auto &fp : cell_nfaces[cell])
              {
                const int face = fp.first, axis = face / 2, side = face % 2;
                for (int q = 0; q < nqf; ++q)
                  {
                    const size_t fq = (size_t)face * nqf + q;
                    const T2<dim> J = jacobian<dim>(vertices, cv, dNgeo_face + fq * nv * dim);
                    const T2<dim> Jinv = invert(J);
                    const double dJ = det(J);
                    for (int i = 0; i < nu * dim; ++i)
                      local_rhs[i] += -(Nu_face[fq * nu + i / dim] * dJ * Jinv[axis][i % dim] * (side ? 1.0 : -1.0) * qwf[q] * fp.second);
                  }
              }
          if (assemble_system) // :337-347
            {
              distribute_local_to_global(dpc, local_matrix.data(), local_rhs.data(), dofs, con, inhom, rowptr, col, A, rhs, true);
              distribute_local_to_global(dpc, local_mass.data(), nullptr, dofs, con, inhom, rowptr, col, M, nullptr, false);
            }
          else // :349-352: vector-only distribute - constrained rows receive nothing
            for (int i = 0; i < dpc; ++i)
              if (!con[dofs[i]])
                {
#pragma omp atomic
                  rhs[dofs[i]] += local_rhs[i];
                }
        }
    }
  }
} // namespace

extern "C" int oracle_insimex_assemble(int dim, int nu, int np, int n_cells, const double *vertices, const int *cells,
                                       const int *cell_dofs, int nq, const double *qw, const double *Nu, const double *dNu,
                                       const double *Np, const double *dNgeo, int nqf, const double *qwf, const double *Nu_face,
                                       const double *dNgeo_face, const double *present, const double *fsi_acc, const int *indicator,
                                       double viscosity, double gamma, double rho, double dt, const double *gravity, int n_bfaces,
                                       const int *bfaces, int n_neumann, const int *neumann_ids, const double *neumann_vals,
                                       const unsigned char *con, const double *inhom, int assemble_system, const int64_t *rowptr,
                                       const int *col, double *A, double *M, double *rhs)
{
  if (dim == 2)
    assemble_imex<2>(nu, np, n_cells, vertices, cells, cell_dofs, nq, qw, Nu, dNu, Np, dNgeo, nqf, qwf, Nu_face, dNgeo_face, present,
                     fsi_acc, indicator, viscosity, gamma, rho, dt, gravity, n_bfaces, bfaces, n_neumann, neumann_ids, neumann_vals,
                     con, inhom, assemble_system, rowptr, col, A, M, rhs);
  else if (dim == 3)
    assemble_imex<3>(nu, np, n_cells, vertices, cells, cell_dofs, nq, qw, Nu, dNu, Np, dNgeo, nqf, qwf, Nu_face, dNgeo_face, present,
                     fsi_acc, indicator, viscosity, gamma, rho, dt, gravity, n_bfaces, bfaces, n_neumann, neumann_ids, neumann_vals,
                     con, inhom, assemble_system, rowptr, col, A, M, rhs);
  else
    return 1;
  return 0;
}
